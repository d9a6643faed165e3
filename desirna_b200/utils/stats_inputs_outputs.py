"""The data formats either side of the design loop, as in the reference's utils/stats_inputs_outputs.py: the input-file
reader (read_input :183-214, InputFile :795-864), the -sf parser (:217-237), the step counters (Stats :671-792) and the
result files a run leaves behind (*_traj.csv :308-334, *_multifasta.fas / *_best_fasta.fas :337-381, *_results.csv
:384-477, *_best_str :480-523, *_stats :526-576).  Plots and the output-directory shuffle are not reproduced."""
import csv
import time

from . import sequence_utils as seq_utils


class InputFile:
    def __init__(self, name, sec_struct, seq_restr):
        self.name = name
        self.sec_struct = sec_struct
        self.seq_restr = seq_restr
        self.pairs = []
        self.alt_pairs = None
        self.seed_seq = None
        self.alt_sec_struct = None
        self.alt_sec_structs = None
        self.target_pairs_tupl = {}
        self.graphs = None
        self.excluded_alt_pairs = None
        self.allsnakes = None

    def add_seed_seq(self, seed_seq):
        self.seed_seq = seed_seq

    def add_alt_sec_struct(self, alt_sec_structs):
        self.alt_sec_struct = alt_sec_structs[0].strip()
        self.alt_sec_structs = [x.strip() for x in alt_sec_structs]

    def set_target_pairs_tupl(self):
        self.target_pairs_tupl = {tuple(pair) for pair in self.pairs}


def make_input(name, sec_struct, seq_restr=None):
    """InputFile with the derived fields DesiRNA.py fills before the loop starts (DesiRNA.py:274-284)."""
    if seq_restr is None:
        seq_restr = "".join("&" if ch == "&" else "N" for ch in sec_struct)
    inp = InputFile(name, sec_struct, seq_restr)
    if len(inp.seq_restr) != len(inp.sec_struct):
        raise ValueError("Secondary structure and sequence restraints are of different length. Check input file.")
    if [i for i, ch in enumerate(inp.seq_restr) if ch == "&"] != [i for i, ch in enumerate(inp.sec_struct) if ch == "&"]:
        raise ValueError("Strand breaks ('&') of the sequence restraints and of the secondary structure differ. Check input file.")
    inp.pairs = seq_utils.check_dot_bracket(inp.sec_struct)
    inp.set_target_pairs_tupl()
    return inp


def read_input(infile):
    """'>key' blocks: name, seq_restr, sec_struct, optional seed_seq and alt_sec_struct (:183-214)"""
    with open(infile, encoding="utf-8") as f:
        text = f.read()
    blocks = {}
    for block in text.lstrip(">").rstrip("\n").split("\n>"):
        lines = block.split("\n")
        blocks[lines[0]] = lines[1:]
    inp = make_input(blocks["name"][0], blocks["sec_struct"][0], blocks["seq_restr"][0])
    if "seed_seq" in blocks:
        inp.add_seed_seq(blocks["seed_seq"][0])
    if "alt_sec_struct" in blocks:
        inp.add_alt_sec_struct(blocks["alt_sec_struct"])
    return inp


def parse_scoring_functions(scoring_f_str):
    """'Ed-Epf:0.5,1-MCC:0.5' -> [(function, weight)].  The reference returns from inside its loop (:232-237), i.e. keeps
    the FIRST term only; kept, because its shipped results were produced that way."""
    for item in scoring_f_str.split(","):
        if ":" not in item:
            raise ValueError(f"Invalid scoring function format: {item}. Expected format: 'function:weight'")
        function, weight = item.split(":")
        return [(function, float(weight))]
    return []


def parse_scoring_functions_all(scoring_f_str):
    """every term of the -sf string (what the reference's help text describes)"""
    out = []
    for item in scoring_f_str.split(","):
        if ":" not in item:
            raise ValueError(f"Invalid scoring function format: {item}. Expected format: 'function:weight'")
        function, weight = item.split(":")
        out.append((function, float(weight)))
    return out


class Stats:
    def __init__(self):
        self.global_step = 0
        self.step = 0
        self.acc_mc_step = 0
        self.acc_mc_better_e = 0
        self.rej_mc_step = 0
        self.acc_re_step = 0
        self.acc_re_better_e = 0
        self.rej_re_step = 0

    def update_global_step(self):
        self.global_step += 1

    def update_step(self, steps):
        self.step += steps

    def update_acc_mc_step(self):
        self.acc_mc_step += 1

    def update_acc_mc_better_e(self):
        self.acc_mc_better_e += 1

    def update_rej_mc_step(self):
        self.rej_mc_step += 1

    def reset_mc_stats(self):
        self.acc_mc_step = self.acc_mc_better_e = self.rej_mc_step = 0

    def update_acc_re_step(self):
        self.acc_re_step += 1

    def update_acc_re_better_e(self):
        self.acc_re_better_e += 1

    def update_rej_re_step(self):
        self.rej_re_step += 1


# ------------------------------------------------------------------------------------------ result files
def _write_dicts(path, rows):
    with open(path, "w", newline="", encoding="utf-8") as f:
        w = csv.DictWriter(f, fieldnames=list(rows[0].keys()))
        w.writeheader()
        w.writerows(rows)


def sort_trajectory(simulation_data):
    return sorted(seq_utils.round_floats(simulation_data), key=lambda d: (d["sim_step"], d["replica_num"]))


def generate_trajectory_csv(sorted_data, output_name):
    _write_dicts(output_name + "_traj.csv", sorted_data)


def _fasta(items, sim_options, now):
    return "".join(f">{sim_options.infile}|{now}|{i['replica_num']}|{i['sim_step']}|{i['scoring_function']}\n{i['sequence']}\n" for i in items)


def generate_multifasta(sorted_data, sim_options, now):
    with open(sim_options.outname + "_multifasta.fas", "w", encoding="utf-8") as f:
        f.write(_fasta(sorted_data, sim_options, now))


def generate_best_fasta(simulation_data, sim_options, now):
    best = sorted(seq_utils.round_floats(simulation_data), key=lambda d: d["scoring_function"])[:sim_options.num_results]
    with open(sim_options.outname + "_best_fasta.fas", "w", encoding="utf-8") as f:
        f.write(_fasta(best, sim_options, now))


def sort_and_filter_simulation_data(simulation_data, sim_options, input_file):
    """distinct sequences (the LAST record of a sequence wins), best first: 1-MCC, then the option-dependent keys (:384-419)"""
    rows = list({item["sequence"]: item for item in simulation_data}.values())
    if sim_options.oligo != "off":
        key = lambda d: (d["mcc"], d["scoring_function"], d["edesired_minus_Epf"], d["Epf"])
    elif sim_options.dimer != "off":
        a, b = input_file.sec_struct.split("&")[:2]
        sign = -1 if a != b else 1
        key = lambda d: (d["mcc"], sign * d["oligo_fraction"], d["edesired_minus_Epf"], d["Epf"])
    elif sim_options.subopt != "off":
        key = lambda d: (d["mcc"], d["edesired_minus_Epf"], -d["esubopt_minus_Epf"])
    else:
        key = lambda d: (d["mcc"], d["edesired_minus_Epf"], d["Epf"], d["scoring_function"])
    return sorted(seq_utils.round_floats(rows), key=key)[:sim_options.num_results]


def generate_csv_from_data(sorted_data, output_name):
    _write_dicts(output_name + "_results.csv", sorted_data)


def check_if_design_solved(sorted_results, input_file, sim_options):
    correct_count = sum(r["mcc"] == 0.0 for r in sorted_results[:10])
    correct_bool = correct_count > 0
    oligo_txt = f",oligo fraction: {sorted_results[0]['oligo_fraction']}" if sim_options.oligo != "off" and correct_bool else ""
    top = sorted_results[0]
    return f">{input_file.name},{correct_bool},{correct_count},{top['sequence']},{top['mfe_ss']}{oligo_txt}", correct_bool


def write_best_str_file(correct_result_txt, output_name):
    with open(output_name + "_best_str", "w", newline="", encoding="utf-8") as f:
        f.write(correct_result_txt)


def generate_simulation_stats_text(stats, sorted_results, correct_bool, finish_time, sim_options):
    sum_mc = stats.acc_mc_step + stats.rej_mc_step
    acc_perc = round(stats.acc_mc_step / sum_mc, 3) if sum_mc else 0
    sum_metro = sum_mc - stats.acc_mc_better_e
    acc_metro = stats.acc_mc_step - stats.acc_mc_better_e
    swaps = (stats.acc_re_step + stats.rej_re_step) or 1
    top = sorted_results[0]
    head = "\nDesign solved succesfully!\n\nBest solution:\n" if correct_bool else "\nDesign not solved!\n\nTarget structure:\n"
    best = (f"{head}{top['sequence']}\nMFE Secondary Structure: \n{top['mfe_ss']}\nPartition Function Energy: {round(top['Epf'], 3)}"
            f"                        \n1-MCC: {round(top['mcc'], 3)}\n")
    return (f"\n>{sim_options.outname} \ntime={sim_options.timlim}s\n\nAcc_ratio={acc_perc}, Iterations={stats.step}, "
            f"Accepted={stats.acc_mc_step}/{sum_mc}, Rejected={stats.rej_mc_step}/{sum_mc}\n"
            f"Accepted Metropolis={acc_metro}/{sum_metro}, Rejected Metropolis={sum_metro - acc_metro}/{sum_metro}\n"
            f"Replica exchange attempts: {stats.global_step}\nReplica swaps attempts: {swaps}\nReplica swaps accepted: {stats.acc_re_step}\n"
            f"Replica swaps rejected: {stats.rej_re_step}\nReplica exchange acc_ratio: {round(stats.acc_re_step / swaps, 3)}\n{best}\n\n"
            f"Simulation time: {time.strftime('%H:%M:%S', time.gmtime(finish_time))}\n")


def write_stats_to_file(stats_txt, output_name):
    with open(output_name + "_stats", "w", newline="\n", encoding="utf-8") as f:
        f.write(stats_txt)


def parse_and_output_results(simulation_data, input_file, stats, finish_time, sim_options, now):
    """the files of parse_and_output_results (:593-633) without the plot and the move into the output directory;
    returns (sorted_results, solved)"""
    trajectory = sort_trajectory(simulation_data)
    generate_trajectory_csv(trajectory, sim_options.outname)
    generate_multifasta(trajectory, sim_options, now)
    generate_best_fasta(simulation_data, sim_options, now)
    sorted_results = sort_and_filter_simulation_data(simulation_data, sim_options, input_file)
    generate_csv_from_data(sorted_results, sim_options.outname)
    txt, solved = check_if_design_solved(sorted_results, input_file, sim_options)
    write_best_str_file(txt, sim_options.outname)
    write_stats_to_file(generate_simulation_stats_text(stats, sorted_results, solved, finish_time, sim_options), sim_options.outname)
    return sorted_results, solved
