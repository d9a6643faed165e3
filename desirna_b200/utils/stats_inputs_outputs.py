"""Input side of the reference's utils/stats_inputs_outputs.py, as far as the design loop needs it: the input-file
reader (read_input :183-214, InputFile :795-864), the -sf parser (:217-237) and the step counters (Stats :671-792).
Plots, CSV/FASTA writers and result post-processing are outside the accelerated path."""
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
    inp = InputFile(name, sec_struct, seq_restr if seq_restr is not None else "N" * len(sec_struct))
    if len(inp.seq_restr) != len(inp.sec_struct):
        raise ValueError("Secondary structure and sequence restraints are of different length. Check input file.")
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
