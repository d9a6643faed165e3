"""CPU: host side of the design loop -- restraints, start sequences, temperature shelves, the move generator's mirror
(desirna_b200/utils/sequence_utils.py) and the input reader, checked against the reference's documented behaviour
(utils/sequence_utils.py:454-1136, utils/stats_inputs_outputs.py:183-237) and its shipped example input."""
import random
from collections import Counter
from types import SimpleNamespace

import numpy as np
import pytest

from desirna_b200.utils import sequence_utils as su
from desirna_b200.utils import stats_inputs_outputs as sio

PAIR_OK = {("A", "U"), ("U", "A"), ("G", "C"), ("C", "G"), ("G", "U"), ("U", "G")}


def opts(**kw):
    o = SimpleNamespace(replicas=10, T_min=10.0, T_max=150.0, acgu_percentages="off", nt_percentages={"A": 15, "C": 30, "G": 30, "U": 15},
                        point_mutations="on", tm_max=0.7, tm_min=0.0, oligo_state="none", diff_start_replicas="one")
    o.__dict__.update(kw)
    o.rep_temps_shelfs = su.get_rep_temps(o)
    return o


def test_check_dot_bracket_families_and_errors():
    assert su.check_dot_bracket("((..))") == [[1, 4], [0, 5]]
    assert su.check_dot_bracket("([.)]") == [[0, 3], [1, 4]]
    with pytest.raises(ValueError):
        su.check_dot_bracket("(()")
    with pytest.raises(ValueError):
        su.check_dot_bracket("())")
    with pytest.raises(ValueError):
        su.check_dot_bracket("(x)")


def test_allowed_letters_follow_the_partner_restraint():
    inp = sio.make_input("t", "((...))", "NANNNNS")
    nts = su.get_nt_list(inp)
    # position 0 pairs with 6 (S = C/G): letters that pair with C or G are G, C, U
    assert sorted(nts[0].letters_allowed) == ["C", "G", "U"]
    assert sorted(nts[6].letters_allowed) == ["C", "G"]
    # position 1 is a fixed A: its partner 5 must be U
    assert nts[1].letters_allowed == ["A"] and nts[5].letters_allowed == ["U"]
    assert sorted(nts[3].letters_allowed) == ["A", "C", "G", "U"] and nts[3].pairs_with is None
    assert list(su.allowed_masks(nts)) == [2 | 4 | 8, 1, 15, 15, 15, 8, 2 | 4]
    with pytest.raises(ValueError):
        su.get_nt_list(sio.make_input("bad", "(...)", "ANNNC"))


def test_rep_temps_defaults():
    t = su.get_rep_temps(opts())
    assert t[0] == 10.0 and t[-1] == 150.0 and len(t) == 10
    assert t[1] == round(10 + 140 / 9, 3)
    assert su.get_rep_temps(opts(replicas=1)) == [150.0]
    assert su.targeted_move_probabilities(opts())[0] == 0.7 and su.targeted_move_probabilities(opts())[-1] == 0.0


def test_initial_sequence_rules():
    random.seed(3)
    inp = sio.make_input("t", "((((....))))..((...))", None)
    nts = su.get_nt_list(inp)
    s = su.initial_sequence_generator(nts, inp, opts())
    for a, b in inp.pairs:
        assert {s[a], s[b]} == {"C", "G"}            # strongest pair where the restraints allow it
    assert s[4] == "G" and s[5:8] == "AAA"            # loop boosting, A elsewhere in loops
    assert s[12] == "G" and s[13] == "A"
    inp2 = sio.make_input("t2", "((.((...))))", "NNNNNNNNNNNN")
    s2 = su.initial_sequence_generator(su.get_nt_list(inp2), inp2, opts())
    assert s2[2] == "A"                                # a one-nucleotide bulge stays A
    inp3 = sio.make_input("t3", "((...))", "WNNNNNW")
    s3 = su.initial_sequence_generator(su.get_nt_list(inp3), inp3, opts())
    assert {s3[0], s3[6]} == {"A", "U"}


def test_expand_cases_bounds():
    assert su.expand_cases([0, 9], 9) == [1, 2, 3, 6, 7, 8, 9]
    assert su.expand_cases([], 9) == []


def test_move_generator_respects_restraints_and_targets_wrong_pairs():
    random.seed(11)
    inp = sio.make_input("t", "((((....))))....", "NNNNNNNNNNNNNNNA")
    nts = su.get_nt_list(inp)
    o = opts(replicas=4)
    cur = SimpleNamespace(sequence="GGGGAAAACCCCAAAA", mfe_ss="((((....))))....", temp_shelf=o.rep_temps_shelfs[0])
    hits = Counter()
    for _ in range(4000):
        m = su.propose_mutation(cur, nts, o, inp)
        diff = [i for i in range(16) if m[i] != cur.sequence[i]]
        assert 1 <= len(diff) <= 2 and 15 not in diff
        for a, b in inp.pairs:
            assert (m[a], m[b]) in PAIR_OK
        hits.update(diff)
    # MFE structure == target: no wrong pairs, moves are uniform over the 15 mutable positions (pairs move together)
    assert min(hits[i] for i in range(15)) > 150
    # one wrong pair at (0, 11): on the coldest shelf 70 % of the moves fall within 3 of position 0 or 11, never on 0 itself
    cur.mfe_ss = ".(((....)))....."
    def first_moved(n_draws):
        out = []
        for _ in range(n_draws):
            m = su.propose_mutation(cur, nts, o, inp)
            out.append([i for i in range(16) if m[i] != cur.sequence[i]][0])
        return out

    moved = first_moved(4000)
    assert sum(1 for i in moved if min(abs(i - 0), abs(i - 11)) <= 3) / 4000 > 0.7
    # hottest shelf: tm_min = 0 -> uniform over the mutable positions again
    cur.temp_shelf = o.rep_temps_shelfs[-1]
    moved = first_moved(2000)
    assert sum(1 for i in moved if i in (4, 5, 6, 7)) > 300


def test_read_input_and_sf_parser(tmp_path):
    p = tmp_path / "in.txt"
    p.write_text(">name\nEte_1\n>seq_restr\nNNNNNNNNNNNNNNNN\n>sec_struct\n(((((......)))))\n")
    inp = sio.read_input(str(p))
    assert inp.name == "Ete_1" and inp.sec_struct == "(((((......)))))" and len(inp.pairs) == 5
    assert (0, 15) in inp.target_pairs_tupl
    assert sio.parse_scoring_functions("Ed-Epf:0.5,1-MCC:0.5") == [("Ed-Epf", 0.5)]      # the reference keeps the first term
    assert sio.parse_scoring_functions_all("Ed-Epf:0.5,1-MCC:0.5") == [("Ed-Epf", 0.5), ("1-MCC", 0.5)]
    with pytest.raises(ValueError):
        sio.parse_scoring_functions("Ed-Epf")


def test_bucketing():
    from desirna_b200.design import bucket_jobs
    g = bucket_jobs([12, 36, 400, 104, 105, 41])
    assert g == [[0, 1], [5], [3], [4], [2]]


def test_job_sharding_is_balanced_and_deterministic():
    from desirna_b200.design import shard_jobs
    lengths = [12, 400, 36, 104, 398, 250, 300, 90, 385, 120, 60, 399]
    for world in (1, 2, 4, 8):
        sh = shard_jobs(lengths, world)
        assert sorted(k for s in sh for k in s) == list(range(len(lengths)))
        assert sh == shard_jobs(lengths, world)
        loads = [sum(lengths[k] ** 3 for k in s) for s in sh]
        if world <= 4:
            assert max(loads) <= 1.5 * (sum(loads) / world)


SHARD_WORKER = r'''
import os, sys, pickle
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from desirna_b200 import design
from desirna_b200.utils import stats_inputs_outputs as sio
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
seen = []
def fake_design_batch(inputs, sim_options, **kw):   # stands in for the GPU loop: the sharding and the gather are host code
    seen.extend(i.name for i in inputs)
    res = [{"name": i.name, "sequence": "A" * len(i.sec_struct), "solved": len(i.sec_struct) % 2 == 0, "seed": kw["seed"]} for i in inputs]
    return res, {"folds": 10 * len(inputs), "seconds": 1.0 + dist.get_rank(), "solved": sum(r["solved"] for r in res), "jobs": len(inputs),
                 "global_steps": 3, "buckets": []}
design.design_batch = fake_design_batch
inputs = [sio.make_input("job%d" % k, "(" * 3 + "." * n + ")" * 3) for k, n in enumerate([3, 30, 4, 90, 5, 60, 7, 8])]
results, info = design.design_batch_sharded(inputs, None, seed=5, global_steps=3)
pickle.dump((seen, results, info), open(sys.argv[4] + ".%s" % sys.argv[3], "wb"))
dist.destroy_process_group()
'''


def test_design_sharding_two_ranks_gloo(tmp_path):
    import os
    import pickle
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "worker.py"
    script.write_text(SHARD_WORKER)
    port = str(31500 + os.getpid() % 2000)
    out = str(tmp_path / "res")
    procs = [subprocess.Popen([sys.executable, str(script), root, port, str(r), out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    got = [pickle.load(open(out + ".%d" % r, "rb")) for r in range(2)]
    # disjoint shards covering every job; both ranks return the same, complete, input-ordered result list
    assert sorted(got[0][0] + got[1][0]) == sorted("job%d" % k for k in range(8))
    assert not set(got[0][0]) & set(got[1][0])
    assert got[0][1] == got[1][1]
    assert [r["name"] for r in got[0][1]] == ["job%d" % k for k in range(8)]
    assert {r["seed"] for r in got[0][1]} == {5 * 131, 5 * 131 + 1}      # per-rank seeds
    for r in range(2):
        info = got[r][2]
        assert info["jobs"] == 8 and info["folds"] == 80 and info["seconds"] == 2.0 and len(info["per_rank"]) == 2
        assert info["solved"] == sum(1 for x in got[0][1] if x["solved"])


def test_result_files_have_the_reference_layout(tmp_path):
    """*_traj.csv / *_results.csv / *_best_str / *_stats / fasta files (utils/stats_inputs_outputs.py:308-633)"""
    from types import SimpleNamespace
    inp = sio.make_input("Ete_1", "(((((......)))))")
    o = SimpleNamespace(infile="in.txt", outname=str(tmp_path / "run"), num_results=2, oligo="off", dimer="off", subopt="off", timlim=60)
    rows = []
    for step, (seq, ss, mcc, s) in enumerate([("GCCUGGAUUAACAGGC", "(((((......)))))", 0.0, 1.25), ("GCCUGGAUUAACAGGU", "((((........))))", 0.123, 2.5),
                                            ("GCCUGGAUUAACAGGC", "(((((......)))))", 0.0, 1.2504), ("ACCUGGAUUAACAGGU", "(((((......)))))", 0.0, 0.9)]):
        rows.append({"sequence": seq, "scoring_function": s, "replica_num": 1 + step % 2, "temp_shelf": 10.0, "sim_step": 100 * (step // 2),
                     "edesired_minus_Epf": s, "Epf": -7.1234567, "edesired": -7.1234567 + s, "mcc": mcc, "mfe_ss": ss})
    st = sio.Stats()
    st.acc_mc_step, st.acc_mc_better_e, st.rej_mc_step, st.step, st.global_step, st.acc_re_step, st.rej_re_step = 120, 80, 80, 200, 2, 3, 1
    best, solved = sio.parse_and_output_results(rows, inp, st, 61.0, o, "20260101.000000")
    assert solved and [r["sequence"] for r in best] == ["ACCUGGAUUAACAGGU", "GCCUGGAUUAACAGGC"]   # distinct, 1-MCC then Ed-Epf
    assert best[1]["scoring_function"] == 1.2504     # the last record of a sequence wins; round_floats leaves LISTS unrounded (reference quirk)
    traj = (tmp_path / "run_traj.csv").read_text().splitlines()
    assert traj[0].startswith("sequence,scoring_function,replica_num,temp_shelf,sim_step") and len(traj) == 5
    assert (tmp_path / "run_best_str").read_text() == ">Ete_1,True,2,ACCUGGAUUAACAGGU,(((((......)))))"
    assert (tmp_path / "run_best_fasta.fas").read_text().startswith(">in.txt|20260101.000000|2|100|0.9\nACCUGGAUUAACAGGU\n")
    stats = (tmp_path / "run_stats").read_text()
    assert "Acc_ratio=0.6, Iterations=200, Accepted=120/200, Rejected=80/200" in stats
    assert "Accepted Metropolis=40/120, Rejected Metropolis=80/120" in stats
    assert "Replica swaps accepted: 3" in stats and "Design solved succesfully!" in stats and "Simulation time: 00:01:01" in stats
    assert len((tmp_path / "run_results.csv").read_text().splitlines()) == 3


def test_homodimer_moves_keep_the_strands_in_step():
    """utils/sequence_utils.py:1102-1128: identical target halves -> identical strands after every move"""
    random.seed(5)
    inp = sio.make_input("homo", "((((....))))&((((....))))", "NNNNNNNNNNNN&NNNNNNNNNNNN")
    nts = su.get_nt_list(inp)
    o = opts(replicas=3, oligo_state="homodimer")
    cur = SimpleNamespace(sequence="GGGGAAAACCCC&GGGGAAAACCCC", mfe_ss="((((....))))&((((....))))", temp_shelf=o.rep_temps_shelfs[1])
    changed = 0
    for _ in range(300):
        m = su.propose_mutation(cur, nts, o, inp)
        a, b = m.split("&")
        assert a == b and len(a) == 12
        for i, j in ((0, 11), (1, 10), (2, 9), (3, 8)):
            assert (a[i], a[j]) in PAIR_OK
        changed += m != cur.sequence
    assert changed > 250
    # different halves: an inter-strand pair's two letters are mirrored onto the other strand (same length strands)
    inp2 = sio.make_input("homo2", "((((....&....))))", "NNNNNNNN&NNNNNNNN")
    nts2 = su.get_nt_list(inp2)
    cur2 = SimpleNamespace(sequence="GGGGAAAA&AAAACCCC", mfe_ss="((((....&....))))", temp_shelf=o.rep_temps_shelfs[1])
    for _ in range(200):
        m = su.propose_mutation(cur2, nts2, o, inp2)
        a, b = m.split("&")
        assert len(a) == len(b) == 8 and set(m) <= set("ACGU&")
    # different halves with a helix INSIDE each strand: the reference's index arithmetic wraps around there; here the two new
    # letters are mirrored onto the other strand (documented choice, same on the device): strands of the right length, the
    # mutated pair identical in both strands
    inp3 = sio.make_input("homo3", "((((....))))..((((&))))..((((....))))", "NNNNNNNNNNNNNNNNNN&NNNNNNNNNNNNNNNNNN")
    nts3 = su.get_nt_list(inp3)
    s3 = "GGGGAAAACCCCAAGGGG"
    cur3 = SimpleNamespace(sequence=s3 + "&" + s3, mfe_ss=inp3.sec_struct, temp_shelf=o.rep_temps_shelfs[1])
    intra = 0
    for _ in range(400):
        m = su.propose_mutation(cur3, nts3, o, inp3)
        a, b = m.split("&")
        assert len(a) == len(b) == 18 and set(m) <= set("ACGU&")
        da = [k for k in range(18) if a[k] != s3[k]]
        db = [k for k in range(18) if b[k] != s3[k]]
        if any(k in (0, 1, 2, 3, 8, 9, 10, 11) for k in da) or any(k in (6, 7, 8, 9, 14, 15, 16, 17) for k in db):   # a pair inside one strand
            intra += 1
            assert da == db and all(a[k] == b[k] for k in da), m
    assert intra > 50


class _FakeLoop:
    """stands in for design.DesignLoop (no GPU): a loop whose jobs get solved after a given number of global steps"""
    made = []

    def __init__(self, inputs, sim_options, seed=0, init_seqs=None):
        self.inputs, self.J, self.R = list(inputs), len(inputs), sim_options.replicas
        self.stride = max(len(i.sec_struct) for i in inputs)
        self.active = np.ones(self.J, np.uint8)
        self.steps = 0
        self.pending = 0
        self.calls = []
        self.solve_at = [len(i.sec_struct) // 10 for i in inputs]     # longer targets take more global steps
        _FakeLoop.made.append(self)

    def run(self, n):
        self.calls.append(int(n))
        self.pending += int(n)

    def busy(self):
        return False

    def _flush(self):
        self.steps += self.pending
        self.pending = 0

    def jobs(self):
        self._flush()
        step = np.array([s if self.steps >= s else -1 for s in self.solve_at], np.int32)
        from desirna_b200.design import REC
        rec = np.zeros((self.J, REC))
        rec[:, 8] = (step < 0)
        return {"sequence": ["A" * len(i.sec_struct) for i in self.inputs], "mfe_ss": [i.sec_struct for i in self.inputs], "rec": rec,
                "solved_step": step, "n_solved": np.zeros(self.J, np.uint32)}

    def set_active(self, mask):
        self.active = np.array(mask, np.uint8)

    def close(self):
        pass


def test_design_batch_scheduling_without_a_gpu(monkeypatch):
    """every length bucket is advanced on its own, solved jobs are retired, step limits and stop_when_solved are honoured"""
    from desirna_b200 import design
    _FakeLoop.made = []
    monkeypatch.setattr(design, "DesignLoop", _FakeLoop)
    inputs = [sio.make_input("j%d" % k, "(" * 4 + "." * n + ")" * 4) for k, n in enumerate([12, 22, 52, 92, 192, 292])]
    o = design.DesignOptions(replicas=4, RE_attempt=10)
    results, info = design.design_batch(inputs, o, global_steps=1000, seed=1)
    assert info["solved"] == 6 and all(r["solved"] for r in results)
    assert [r["name"] for r in results] == ["j%d" % k for k in range(6)]
    assert len(_FakeLoop.made) == len(info["buckets"]) >= 4                  # one loop per length bucket
    for loop in _FakeLoop.made:
        assert not loop.active.any()                                       # every job retired once solved
        assert loop.steps >= max(loop.solve_at) and loop.steps <= max(loop.solve_at) + 64
        assert loop.calls[0] == 1 and max(loop.calls) <= 64                  # first batch is one step, later ones sized from the pace
    # a step limit stops unsolved loops; stop_when_solved=False keeps solved jobs running to the limit
    _FakeLoop.made = []
    results, info = design.design_batch(inputs, o, global_steps=5, seed=1, stop_when_solved=False)
    assert all(loop.steps == 5 for loop in _FakeLoop.made) and info["global_steps"] == 5
    assert [r["solved"] for r in results] == [True, True, False, False, False, False]
    # a time limit of zero: only the start sequences are read
    _FakeLoop.made = []
    results, info = design.design_batch(inputs, o, time_limit=0.0, seed=1)
    assert all(loop.steps == 0 for loop in _FakeLoop.made) and info["solved"] == 0


# ------------------------------------------------------------------------------------------------ alternative structures: snake graphs
def test_snake_graphs_and_moves_match_the_reference_draw_for_draw():
    """tests/golden/S1.json was produced by importing the reference's utils/sequence_utils.py (tests/golden/make_snake_golden.py):
    conflict graphs and their colourings (order included: the first colouring is the start state), the pairs moved into the
    ordinary restraints, per-position letters, start sequences and 2 x 150 consecutive moves from the same seeds.
    The reference draws the partner letter of a pair move from an UNSORTED set of strings (utils/sequence_utils.py:1062-1070), so
    its draws depend on the interpreter's string-hash seed: vectors and check both run under PYTHONHASHSEED=0 (a subprocess)."""
    import os
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    env = dict(os.environ, PYTHONHASHSEED="0")
    out = subprocess.run([sys.executable, os.path.join(here, "snake_golden_check.py")], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "moves checked" in out.stdout and int(out.stdout.split("moves checked")[0].split()[-1]) >= 1500
