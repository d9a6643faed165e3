"""CPU: the lock-step replica loop makes the same decisions, from the same random draws, as the reference's
one-process-per-replica loop (utils/replica_exchange_monte_carlo.py:176-271), also when sharded over two ranks."""
import os
import random
import subprocess
import sys
import types

import pytest

import oracle_backend

from conftest import PAR1999, ROOT, load_golden

TARGET = "((((((.((((((((....))))).)).).))))))"


class Stats:
    def __init__(self):
        self.global_step = self.step = 0
        self.acc_mc_step = self.acc_mc_better_e = self.rej_mc_step = 0
        self.acc_re_step = self.acc_re_better_e = self.rej_re_step = 0

    def update_step(self, n): self.step += n
    def reset_mc_stats(self): self.acc_mc_step = self.acc_mc_better_e = self.rej_mc_step = 0
    def update_acc_mc_step(self): self.acc_mc_step += 1
    def update_acc_mc_better_e(self): self.acc_mc_better_e += 1
    def update_rej_mc_step(self): self.rej_mc_step += 1
    def update_acc_re_step(self): self.acc_re_step += 1
    def update_acc_re_better_e(self): self.acc_re_better_e += 1
    def update_rej_re_step(self): self.rej_re_step += 1


def toy_mutate(seq_obj, nt_list, sim_options, input_file):
    """stand-in for DesiRNA's move generator: a point mutation drawn from the global `random` stream, ending, like
    sequence_utils.mutate_sequence (:1132-1136), in es.score_sequence + the two stamps"""
    from desirna_b200.utils import energy_scores as es
    s = seq_obj.sequence
    pos = random.randrange(len(s))
    new = s[:pos] + random.choice("ACGU") + s[pos + 1:]
    out = es.score_sequence(new, input_file, sim_options)
    out.get_replica_num(seq_obj.replica_num)
    out.get_temp_shelf(seq_obj.temp_shelf)
    return out


def setup(R):
    from desirna_b200 import RNA
    from desirna_b200.utils import energy_scores as es
    from oracle_backend import OracleBackend
    oracle_backend.install(OracleBackend(PAR1999))
    opt = types.SimpleNamespace(oligo_state="none", pks="off", scoring_f=[("Ed-Epf", 1.0)], subopt="off", motifs={}, RE_attempt=6, L=504.12, replicas=R)
    inp = types.SimpleNamespace(sec_struct=TARGET, alt_sec_struct=None, alt_sec_structs=None)
    rows = load_golden("G1")[:R]
    objs = []
    for r, row in enumerate(rows):
        o = es.score_sequence(row["sequence"], inp, opt)
        o.get_replica_num(r + 1)
        o.get_temp_shelf(10.0 + 20.0 * r)
        objs.append(o)
    return opt, inp, objs


def reference_order(objs, opt, inp):
    """what the reference's pool does: replica after replica, each on random.seed(index)"""
    from desirna_b200.utils import replica_exchange_monte_carlo as remc
    out, tot = [], Stats()
    for idx, o in enumerate(objs):
        random.seed(idx)
        ws = Stats()
        res, ws = remc.single_replica_design(o, None, ws, opt, inp, mutate=toy_mutate)
        out.append(res)
        for a in ("acc_mc_step", "acc_mc_better_e", "rej_mc_step"):
            setattr(tot, a, getattr(tot, a) + getattr(ws, a))
    return out, tot


def test_lockstep_equals_sequential():
    from desirna_b200 import RNA
    from desirna_b200.utils import replica_exchange_monte_carlo as remc
    old = oracle_backend.current()
    try:
        opt, inp, objs = setup(5)
        want, wstats = reference_order(objs, opt, inp)
        be = oracle_backend.current()
        calls0 = be.calls
        random.seed(2137)
        got, st = remc.mutate_sequence_re(objs, None, Stats(), opt, inp, mutate=toy_mutate)
        assert random.random() == random.Random(2137).random()  # the parent stream is untouched
        assert be.calls - calls0 == opt.RE_attempt              # ONE engine call per sub-step for all replicas
        assert [vars(a) for a in got] == [vars(b) for b in want]
        assert (st.acc_mc_step, st.acc_mc_better_e, st.rej_mc_step) == (wstats.acc_mc_step, wstats.acc_mc_better_e, wstats.rej_mc_step)
        assert st.step == opt.RE_attempt
    finally:
        oracle_backend.install(old)


def test_replica_exchange_swaps_neighbours():
    from desirna_b200.utils import replica_exchange_monte_carlo as remc
    old = oracle_backend.current()
    try:
        opt, inp, objs = setup(5)
        st = Stats(); st.global_step = 1          # odd step: pairs (0,1), (2,3)
        objs[0].scoring_function, objs[1].scoring_function = 5.0, 1.0   # colder replica is worse -> certain swap
        t0, t1 = objs[0].temp_shelf, objs[1].temp_shelf
        random.seed(1)
        out, st = remc.replica_exchange(objs, st, opt)
        assert [o.replica_num for o in out] == [1, 2, 3, 4, 5]
        assert (out[0].temp_shelf, out[1].temp_shelf) == (t1, t0)
        assert st.acc_re_step + st.rej_re_step == 2
    finally:
        oracle_backend.install(old)


WORKER = r'''
import os, sys, random, pickle
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
import test_remc_lockstep as T
from desirna_b200.utils import replica_exchange_monte_carlo as remc
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
opt, inp, objs = T.setup(5)
random.seed(2137)
got, st = remc.mutate_sequence_re(objs, None, T.Stats(), opt, inp, mutate=T.toy_mutate)
st.global_step = 1
got, st = remc.replica_exchange(got, st, opt)
pickle.dump(([vars(o) for o in got], vars(st)), open(sys.argv[4] + ".%s" % sys.argv[3], "wb"))
dist.destroy_process_group()
'''


def test_two_ranks_gloo_match_single_process(tmp_path):
    import pickle
    from desirna_b200 import RNA
    from desirna_b200.utils import replica_exchange_monte_carlo as remc
    old = oracle_backend.current()
    try:
        opt, inp, objs = setup(5)
        random.seed(2137)
        want, st = remc.mutate_sequence_re(objs, None, Stats(), opt, inp, mutate=toy_mutate)
        st.global_step = 1
        want, st = remc.replica_exchange(want, st, opt)
    finally:
        oracle_backend.install(old)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    out = str(tmp_path / "res")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r), out]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=240) == 0
    for r in range(2):
        got, gst = pickle.load(open(out + ".%d" % r, "rb"))
        assert len(got) == len(want)
        for a, b in zip(got, want):
            vb = vars(b)
            for k, v in a.items():
                if isinstance(v, float):
                    assert abs(v - vb[k]) < 1e-12, k
                else:
                    assert v == vb[k], k
        for k in ("acc_mc_step", "acc_mc_better_e", "rej_mc_step", "acc_re_step", "rej_re_step", "step"):
            assert gst[k] == vars(st)[k], k


def test_lockstep_with_the_mirrored_move_generator():
    """the lock-step loop driven by desirna_b200.utils.sequence_utils.mutate_sequence (targeted moves, paired letters)
    takes the same decisions from the same draws as replica-after-replica execution"""
    from desirna_b200 import RNA
    from desirna_b200.utils import replica_exchange_monte_carlo as remc
    from desirna_b200.utils import sequence_utils as su
    from desirna_b200.utils import stats_inputs_outputs as sio
    old = oracle_backend.current()
    try:
        opt, _, objs = setup(4)
        inp = sio.make_input("t", TARGET)
        opt.__dict__.update(point_mutations="on", tm_max=0.7, tm_min=0.0, acgu_percentages="off", nt_percentages={"A": 15, "C": 30, "G": 30, "U": 15},
                            rep_temps_shelfs=[o.temp_shelf for o in objs])
        nts = su.get_nt_list(inp)
        want = []
        for idx, o in enumerate(objs):
            random.seed(idx)
            res, _ = remc.single_replica_design(o, nts, Stats(), opt, inp, mutate=su.mutate_sequence)
            want.append(res)
        random.seed(1)
        got, st = remc.mutate_sequence_re(objs, nts, Stats(), opt, inp)      # default move generator = the mirror
        assert [vars(a) for a in got] == [vars(b) for b in want]
        assert any(a.sequence != b.sequence for a, b in zip(got, objs))
        for o in got:
            for a, b in inp.pairs:
                assert (o.sequence[a], o.sequence[b]) in {("A", "U"), ("U", "A"), ("G", "C"), ("C", "G"), ("G", "U"), ("U", "G")}
    finally:
        oracle_backend.install(old)


def test_replica_streams_draw_like_seed_getstate_setstate():
    """the per-replica generators of the lock-step loop (random.Random seeded with the replica index, bound under the module-level
    names) give the draws of the reference's scheme: random.seed(r) in each worker, its state carried from use to use
    (utils/replica_exchange_monte_carlo.py:227-228, :250); the parent stream is left where it was"""
    from desirna_b200.utils import replica_exchange_monte_carlo as remc
    random.seed(99)
    parent_before = random.getstate()
    # reference scheme: one module-level stream, state saved / restored around every use
    states = {}
    for r in (0, 3, 5):
        random.seed(r)
        states[r] = random.getstate()
    want = []
    for rnd in range(4):
        for r in (0, 3, 5):
            random.setstate(states[r])
            want.append((random.choice("ACGU"), random.random(), random.choices([1, 2, 3], weights=[0.2, 0.3, 0.5])[0], random.randint(0, 99)))
            states[r] = random.getstate()
    random.setstate(parent_before)
    streams = remc._ReplicaStreams([0, 3, 5])
    got = []
    try:
        for rnd in range(4):
            for r in (0, 3, 5):
                streams.use(r)
                got.append((random.choice("ACGU"), random.random(), random.choices([1, 2, 3], weights=[0.2, 0.3, 0.5])[0], random.randint(0, 99)))
            streams.restore()
    finally:
        streams.restore()
    assert got == want
    assert random.getstate() == parent_before
    assert random.random.__self__ is random._inst
