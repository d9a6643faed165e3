"""GPU: the device-resident design loop (bf_design_*, csrc/bf_design.cu) -- its on-device score records against the
host mirror of the reference's scoring (golden-pinned in test_host_scoring.py / test_gpu_parity.py), its move generator
against the host mirror of the reference's mutate_sequence in distribution, and the loop's invariants."""
import random
from collections import Counter
from types import SimpleNamespace

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def f32(x):
    import struct
    return struct.unpack("f", struct.pack("f", x))[0]

PAIR_OK = {("A", "U"), ("U", "A"), ("G", "C"), ("C", "G"), ("G", "U"), ("U", "G")}


def small_inputs(max_len=48, limit=8):
    from desirna_b200.utils import stats_inputs_outputs as sio
    rows = [r for r in load_golden("E1") if len(r["target"]) <= max_len and not set(r["target"]) - set(".()")]
    return [sio.make_input(r["file"], r["target"]) for r in rows[:limit]]


def host_opts(o):
    """sim_options as the host scoring mirror wants them"""
    return o


@pytest.mark.parametrize("overlap", ["0", "1"])
@pytest.mark.parametrize("scoring", [[("Ed-Epf", 1.0)], [("Ed-Epf", 0.5), ("1-MCC", 0.5)],
                                     [("sln_Epf", 1.0), ("Ed-MFE", 0.7), ("1-precision", 0.2), ("1-recall", 0.1)], [("Edef", 5.0), ("Ed-Epf", 0.3)]])
def test_records_match_host_scoring_and_invariants_hold(engine, monkeypatch, scoring, overlap):
    """overlap = 1 (default): in small batches the partition function runs beside the MFE fill, scaled by the PARENT sequence's MFE
    instead of the mutant's own; the ensemble energy then agrees with the host path to rounding (<= 1 float32 ulp after the API's
    float32 rounding) instead of bit for bit."""
    monkeypatch.setenv("BF_DESIGN_OVERLAP", overlap)
    exact = overlap == "0"
    from desirna_b200 import design
    from desirna_b200.utils import energy_scores as es
    inputs = small_inputs()
    o = design.DesignOptions(replicas=6, RE_attempt=15, scoring_f=scoring)
    random.seed(5)
    loop = design.DesignLoop(inputs, o, seed=7)
    loop.run(4)
    rep = loop.replicas()
    jobs = loop.jobs()
    J, R = loop.J, loop.R
    # every replica made RE_attempt decisions per global step
    assert (rep["counts"][:, 0] + rep["counts"][:, 2] == 4 * 15).all()
    assert (rep["counts"][:, 1] <= rep["counts"][:, 0]).all()
    # shelves stay a permutation within a job
    assert (np.sort(rep["shelf"], axis=1) == np.arange(R)).all()
    for j, inp in enumerate(inputs):
        seqs = rep["sequence"][j * R:(j + 1) * R]
        ref = es.score_sequences(seqs, inp, o)
        for r, (s, h) in enumerate(zip(seqs, ref)):
            g = j * R + r
            rec = dict(zip(design.REC_FIELDS, rep["rec"][g]))
            for a, b in inp.pairs:
                assert (s[a], s[b]) in PAIR_OK, (inp.name, s)
            assert rep["mfe_ss"][g] == h.mfe_ss
            assert rec["edesired"] == h.edesired                                   # float32-rounded API values, bit for bit
            assert rec["Epf"] == h.Epf if exact else abs(rec["Epf"] - h.Epf) <= 4e-6
            assert rec["mcc"] == pytest.approx(h.mcc, abs=1e-12) and rec["precision"] == pytest.approx(h.precision, abs=1e-12)
            assert rec["recall"] == pytest.approx(h.recall, abs=1e-12)
            assert rec["scoring_function"] == pytest.approx(h.scoring_function, abs=1e-9 if exact else 1e-5)
            if any(f == "Ed-MFE" for f, _ in scoring):
                assert rec["MFE"] == h.MFE
            if any(f == "Edef" for f, _ in scoring):
                assert rec["ensemble_defect"] == pytest.approx(h.ensemble_defect, abs=1e-9)
            assert (rec["distance"] == 0) == (h.mfe_ss == inp.sec_struct)
        # the per-job best is one of the states seen, and is at least as good as every current state
        cur_key = min((rep["rec"][j * R + r][8], rep["rec"][j * R + r][0]) for r in range(R))
        assert (jobs["rec"][j][8], jobs["rec"][j][0]) <= cur_key
        if jobs["solved_step"][j] >= 0:
            assert jobs["mfe_ss"][j] == inp.sec_struct and jobs["rec"][j][8] == 0
    loop.close()


def test_device_records_against_the_oracle(engine, oracle):
    """test_records_match_host_scoring... compares the device loop with the host mirror of score_sequence, and that mirror folds
    through the same CUDA kernels -- it checks the loop's bookkeeping, not the folds.  Here the records of a few global steps are
    checked against the CPU oracle instead: MFE structure and energy, energy of the target (Ed), ensemble energy (Epf)."""
    from desirna_b200 import design
    inputs = small_inputs(limit=5)
    o = design.DesignOptions(replicas=5, RE_attempt=10, scoring_f=[("Ed-MFE", 0.5), ("Ed-Epf", 0.5)])
    random.seed(11)
    loop = design.DesignLoop(inputs, o, seed=3)
    loop.run(3)
    rep = loop.replicas()
    R = loop.R
    for j, inp in enumerate(inputs):
        for r in range(R):
            g = j * R + r
            s = rep["sequence"][g]
            rec = dict(zip(design.REC_FIELDS, rep["rec"][g]))
            e, ss = oracle.mfe(s)
            assert rep["mfe_ss"][g] == ss and rec["MFE"] == f32(e / 100.0), (inp.name, s)
            assert rec["edesired"] == f32(oracle.eval(s, inp.sec_struct) / 100.0), (inp.name, s)
            assert abs(rec["Epf"] - oracle.pf(s)[4]) <= 1e-5 * max(1.0, abs(rec["Epf"])), (inp.name, s)   # float32-rounded API value
    loop.close()


def test_alternative_structures_and_motifs_on_the_device(engine):
    """the reference's Alternative-structures example (example_files/inputs/Alternative_structures_design_input.txt: one target, two
    alternative structures, scored as mean(eval(alt)) - Epf, utils/energy_scores.py:98-102) next to a plain target, both with two
    IUPAC motifs (-motifs, utils/sequence_utils.py:1231-1256): device records against the host mirror of score_sequence"""
    from desirna_b200 import design
    from desirna_b200.utils import energy_scores as es
    from desirna_b200.utils import stats_inputs_outputs as sio
    alt = sio.make_input("Alt_Struct_Example", "((((((.((((((((....))))).)).).))))))")
    alt.add_alt_sec_struct(["(((((((((((((....)))..)).)).).))))).", "(((((((((((((....)))))...)).).)))))."])
    plain = sio.make_input("plain", "((((....))))....((((...)))).")
    inputs = [alt, plain]
    o = design.DesignOptions(replicas=6, RE_attempt=12, scoring_f=[("Ed-Epf", 0.7), ("1-MCC", 0.3)], motifs={"GNRA": -3.5, "UUCG": 2.0})
    random.seed(21)
    loop = design.DesignLoop(inputs, o, seed=9)
    loop.run(3)
    rep = loop.replicas()
    R = loop.R
    seen_motif = 0
    for j, inp in enumerate(inputs):
        seqs = rep["sequence"][j * R:(j + 1) * R]
        ref = es.score_sequences(seqs, inp, o)
        for r, (s, h) in enumerate(zip(seqs, ref)):
            rec = dict(zip(design.REC_FIELDS, rep["rec"][j * R + r]))
            assert rec["edesired"] == h.edesired and abs(rec["Epf"] - h.Epf) <= 4e-6
            if inp is alt:
                assert rec["edesired2"] == pytest.approx(h.edesired2, abs=1e-12)
            else:
                assert rec["edesired2"] == 0.0
            want_motif = es.score_motifs(s, o)
            assert rec["motif_bonus"] == want_motif, (s, rec["motif_bonus"], want_motif)
            seen_motif += want_motif != 0
            assert rec["scoring_function"] == pytest.approx(h.scoring_function, abs=1e-5)
    assert seen_motif > 0   # (GNRA is common enough in 12 random-ish sequences)
    loop.close()


def test_pseudoknot_overlay_on_the_device(engine):
    """the reference's Pseudoknot example (example_files/inputs/Pseudoknot_design_input.txt): after the MFE fold of every mutant the
    paired positions are forbidden, the sequence is folded again and the new pairs painted with the next bracket family
    (utils/sequence_utils.py:1166-1228).  Device records and overlaid structures against the host mirror (pks = "on")."""
    from desirna_b200 import design
    from desirna_b200.utils import energy_scores as es
    from desirna_b200.utils import stats_inputs_outputs as sio
    inputs = [sio.make_input("Pseudoknot_Example", "((((((....[[[[..))))))......]]]]...."),
              sio.make_input("two_knots", "..((((..[[[..))))..<<<..]]]...>>>.."),
              sio.make_input("plain", "((((....))))....((((...)))).")]
    o = design.DesignOptions(replicas=6, RE_attempt=12, scoring_f=[("Ed-Epf", 0.5), ("1-MCC", 0.5)], pks="on")
    random.seed(33)
    loop = design.DesignLoop(inputs, o, seed=13)
    loop.run(3)
    rep = loop.replicas()
    R = loop.R
    painted = 0
    for j, inp in enumerate(inputs):
        seqs = rep["sequence"][j * R:(j + 1) * R]
        ref = es.score_sequences(seqs, inp, o)
        for r, (s, h) in enumerate(zip(seqs, ref)):
            g = j * R + r
            rec = dict(zip(design.REC_FIELDS, rep["rec"][g]))
            assert rep["mfe_ss"][g] == h.mfe_ss, (inp.name, s, rep["mfe_ss"][g], h.mfe_ss)
            painted += "[" in h.mfe_ss
            assert rec["edesired"] == h.edesired and abs(rec["Epf"] - h.Epf) <= 4e-6
            assert rec["mcc"] == pytest.approx(h.mcc, abs=1e-12) and rec["precision"] == pytest.approx(h.precision, abs=1e-12)
            assert rec["recall"] == pytest.approx(h.recall, abs=1e-12)
            assert rec["scoring_function"] == pytest.approx(h.scoring_function, abs=1e-5)
    assert painted > 0   # some replica did get a second layer of pairs
    loop.close()


def test_the_six_example_scenarios_run_through_design_batch(engine):
    """every input of the reference's example_files/inputs (Standard, Seed sequence, Alternative structures, Pseudoknot, RNA-RNA
    complex, Homodimer) is designed by the device-resident loop; every reported solution is folded again through the host mirror"""
    from desirna_b200 import design
    from desirna_b200.utils import energy_scores as es
    from desirna_b200.utils import stats_inputs_outputs as sio
    std = sio.make_input("Design", "((((((.((((((((....))))).)).).))))))")
    seed = sio.make_input("Seed_Seq_Example", "((((((.((((((((....))))).)).).))))))")
    seed.add_seed_seq("GCCCCGGCCCCCGGCGAAAGCCGGUGGAGGCGGGGC")
    alt = sio.make_input("Alt_Struct_Example", "((((((.((((((((....))))).)).).))))))")
    alt.add_alt_sec_struct(["(((((((((((((....)))..)).)).).))))).", "(((((((((((((....)))))...)).).)))))."])
    pk = sio.make_input("Pseudoknot_Example", "((((((....[[[[..))))))......]]]]....")
    het = sio.make_input("RNA-RNA Complex_Example", "(((.(((((....))..&(((....)))..))))))", "NNNNNNNNNNNNNNNNN&NNNNNNNNNNNNNNNNNN")
    hom = sio.make_input("Homodimer_Example", "((((....((((.....&))))....)))).....", "NNNNNNNNNNNNNNNNN&NNNNNNNNNNNNNNNNN")
    runs = [([std, seed, alt], design.DesignOptions(replicas=10, RE_attempt=20)),
            ([pk], design.DesignOptions(replicas=10, RE_attempt=20, pks="on")),
            ([het], design.DesignOptions(replicas=10, RE_attempt=20, oligo_state="heterodimer")),
            ([hom], design.DesignOptions(replicas=10, RE_attempt=20, oligo_state="homodimer"))]
    random.seed(2)
    solved = 0
    for inputs, o in runs:
        res, info = design.design_batch(inputs, o, global_steps=30, seed=17)
        assert len(res) == len(inputs) and info["jobs"] == len(inputs)
        for inp, r in zip(inputs, res):
            h = es.score_sequences([r["sequence"]], inp, o)[0]
            assert h.mfe_ss == r["mfe_ss"], (inp.name, r["sequence"])
            assert r["scoring_function"] == pytest.approx(h.scoring_function, abs=1e-5)
            if r["solved"]:
                solved += 1
                assert h.mfe_ss == inp.sec_struct or sorted(map(tuple, sio.seq_utils.check_dot_bracket(h.mfe_ss))) == sorted(map(tuple, inp.pairs))
    assert solved >= 4   # six 35-nt targets, 600 Monte-Carlo sub-steps x 10 replicas each


def test_negative_design_term_on_the_device(engine):
    """-nd on (utils/energy_scores.py:104-107): a mutant that folds into its target also pays Epf - E(second-best structure); the
    device loop gets that energy from the 2-best DP (csrc/bf_twobest.cu) for exactly those rows.  Records against the host mirror."""
    from desirna_b200 import design
    from desirna_b200.utils import energy_scores as es
    from desirna_b200.utils import stats_inputs_outputs as sio
    inputs = [sio.make_input("hp", "((((....))))"), sio.make_input("two", "((((....))))..(((....)))"), sio.make_input("std", "((((((.((((((((....))))).)).).))))))")]
    o = design.DesignOptions(replicas=8, RE_attempt=25, scoring_f=[("Ed-Epf", 1.0)], subopt="on")
    random.seed(41)
    loop = design.DesignLoop(inputs, o, seed=19)
    loop.run(6)
    rep = loop.replicas()
    R = loop.R
    hits = 0
    for j, inp in enumerate(inputs):
        seqs = rep["sequence"][j * R:(j + 1) * R]
        ref = es.score_sequences(seqs, inp, o)
        for r, (s, h) in enumerate(zip(seqs, ref)):
            rec = dict(zip(design.REC_FIELDS, rep["rec"][j * R + r]))
            assert rep["mfe_ss"][j * R + r] == h.mfe_ss
            assert rec["subopt_e"] == h.subopt_e, (inp.name, s, rec["subopt_e"], h.subopt_e)
            assert (rec["subopt_e"] != 0) <= (h.mfe_ss == inp.sec_struct)
            hits += h.mfe_ss == inp.sec_struct
            assert rec["scoring_function"] == pytest.approx(h.scoring_function, abs=1e-5)
    assert hits > 0   # some replica reached its target, i.e. the term was exercised
    loop.close()


def test_same_seed_same_trajectory(engine):
    from desirna_b200 import design
    inputs = small_inputs(limit=4)
    o = design.DesignOptions(replicas=4, RE_attempt=10)
    out = []
    for seed in (3, 3, 4):
        random.seed(1)
        loop = design.DesignLoop(inputs, o, seed=seed)
        loop.run(3)
        out.append(loop.replicas()["sequence"])
        loop.close()
    assert out[0] == out[1]
    assert out[0] != out[2]


def test_move_generator_matches_host_mirror_in_distribution(engine):
    """bf_k_design_propose against desirna_b200.utils.sequence_utils.propose_mutation (the reference's move,
    utils/sequence_utils.py:926-1100): same distribution of mutants per temperature shelf."""
    from desirna_b200 import design
    from desirna_b200.utils import sequence_utils as su
    from desirna_b200.utils import stats_inputs_outputs as sio
    inp = sio.make_input("t", "((((....))))...((...))", "NNNNNNNNSNNNNNANNNNNNN")
    o = design.DesignOptions(replicas=3, RE_attempt=1, tm_max=0.8, tm_min=0.1)
    start = "GGGAAAAACCCCAAAGGAAACC"      # folds into something else than the target: targeted moves are active
    loop = design.DesignLoop([inp], o, seed=11, init_seqs=[start] * 3)
    cur_ss = loop.replicas()["mfe_ss"][0]
    assert cur_ss != inp.sec_struct
    N = 6000
    dev = [Counter() for _ in range(3)]
    for _ in range(N):
        for r, m in enumerate(loop.propose_only()):
            dev[r][m] += 1
    loop.close()
    nts = su.get_nt_list(inp)
    random.seed(2)
    for r in range(3):
        cur = SimpleNamespace(sequence=start, mfe_ss=cur_ss, temp_shelf=o.rep_temps_shelfs[r])
        host = Counter(su.propose_mutation(cur, nts, o, inp) for _ in range(N))
        assert set(dev[r]) <= set(host) | {k for k in dev[r] if dev[r][k] < 5}, "device proposes mutants the reference cannot"

        def where(counter):   # marginal over the mutated positions (<= ~25 categories: sampling noise ~0.03 at N = 6000)
            out = Counter()
            for m, c in counter.items():
                out[tuple(i for i in range(len(start)) if m[i] != start[i])] += c
            return out

        hw, dw = where(host), where(dev[r])
        tv_pos = 0.5 * sum(abs(hw[k] - dw[k]) for k in set(hw) | set(dw)) / N
        tv_all = 0.5 * sum(abs(host[k] - dev[r][k]) for k in set(host) | set(dev[r])) / N
        assert tv_pos < 0.06, (r, tv_pos)
        assert tv_all < 0.15, (r, tv_all)


def test_snake_moves_on_the_device_match_the_host_mirror(engine):
    """alternative structures: positions in more than one pair across the target and the alternatives form conflict graphs whose
    nodes change together from one Watson-Crick colouring to another (utils/sequence_utils.py:119-396, 1085-1094); the clash-free
    pairs of the alternatives are pair restraints of the move generator too.  bf_k_design_propose against the host mirror (itself
    pinned draw for draw to the reference, tests/golden/S1.json): same set of mutants, same distribution."""
    from desirna_b200 import design
    from desirna_b200.utils import sequence_utils as su
    from desirna_b200.utils import stats_inputs_outputs as sio
    inp = sio.make_input("switch", "((((((....))))))....((((....))))")
    inp.add_alt_sec_struct(["....((((((....))))))((((....))))", "((((((....))))))....((((....))))"])
    o = design.DesignOptions(replicas=3, RE_attempt=1, tm_max=0.8, tm_min=0.1)
    random.seed(3)
    loop = design.DesignLoop([inp], o, seed=5)
    assert inp.graphs and loop.replicas()["sequence"][0] == loop.replicas()["sequence"][1]
    start = loop.replicas()["sequence"][0]
    cur_ss = loop.replicas()["mfe_ss"][0]
    N = 6000
    dev = [Counter() for _ in range(3)]
    for _ in range(N):
        for r, m in enumerate(loop.propose_only()):
            dev[r][m] += 1
    loop.close()
    nts = su.get_nt_list(inp)
    snake_nodes = {x for g in inp.graphs for x in g["numbers"]}
    random.seed(2)
    for r in range(3):
        cur = SimpleNamespace(sequence=start, mfe_ss=cur_ss, temp_shelf=o.rep_temps_shelfs[r])
        host = Counter(su.propose_mutation(cur, nts, o, inp) for _ in range(N))
        assert set(dev[r]) <= set(host) | {k for k in dev[r] if dev[r][k] < 5}, "device proposes mutants the reference cannot"
        jumps = lambda c: sum(v for m, v in c.items() if any(m[x] != start[x] for x in snake_nodes))
        assert jumps(host) > 300 and abs(jumps(host) - jumps(dev[r])) < 0.05 * N          # whole-graph jumps are as frequent
        tv_all = 0.5 * sum(abs(host[k] - dev[r][k]) for k in set(host) | set(dev[r])) / N
        assert tv_all < 0.15, (r, tv_all)


def test_design_batch_solves_short_eterna_targets(engine, oracle):
    from desirna_b200 import design
    inputs = small_inputs(max_len=40, limit=10)
    o = design.DesignOptions(replicas=10, RE_attempt=100)
    results, info = design.design_batch(inputs, o, global_steps=30, seed=1)
    assert info["solved"] >= len(inputs) - 1, info
    for inp, res in zip(inputs, results):
        if res["solved"]:
            e, ss = oracle.mfe(res["sequence"])
            assert ss == inp.sec_struct, (inp.name, res["sequence"])
            assert res["mfe_ss"] == inp.sec_struct and res["distance"] == 0
            assert oracle.eval(res["sequence"], inp.sec_struct) == e


HETERO = ("(((.(((((....))..&(((....)))..))))))", "NNNNNNNNNNNNNNNNN&NNNNNNNNNNNNNNNNNN")   # example_files/inputs/RNA_RNA_complex_design_input.txt


def test_heterodimer_records_match_host_scoring(engine):
    """two-strand jobs: FAB as Epf, eval with the nick, '&' -> 'Ee' in the similarity scores, -kT ln(dimer fraction) bonus
    (utils/energy_scores.py:79,153-158,421-430; utils/dimer_multichain_energy.py:36-63)"""
    from desirna_b200 import design
    from desirna_b200.utils import energy_scores as es
    from desirna_b200.utils import stats_inputs_outputs as sio
    inp = sio.make_input("complex", *HETERO)
    o = design.DesignOptions(replicas=8, RE_attempt=20, oligo_state="heterodimer", scoring_f=[("Ed-Epf", 1.0), ("1-MCC", 0.3), ("sln_Epf", 0.2)])
    random.seed(4)
    loop = design.DesignLoop([inp], o, seed=9)
    loop.run(5)
    rep = loop.replicas()
    ref = es.score_sequences(rep["sequence"], inp, o)
    assert (np.sort(rep["shelf"], axis=1) == np.arange(8)).all()
    moved = 0
    for g, (s, h) in enumerate(zip(rep["sequence"], ref)):
        rec = dict(zip(design.REC_FIELDS, rep["rec"][g]))
        assert s.index("&") == 17 and rep["mfe_ss"][g] == h.mfe_ss
        for a, b in inp.pairs:
            assert (s[a], s[b]) in PAIR_OK
        assert rec["edesired"] == h.edesired and rec["Epf"] == h.Epf
        assert rec["mcc"] == pytest.approx(h.mcc, abs=1e-12) and rec["recall"] == pytest.approx(h.recall, abs=1e-12)
        assert rec["oligo_fraction"] == pytest.approx(h.oligo_fraction, rel=1e-9)
        assert rec["oligomer_bonus"] == pytest.approx(h.oligomer_bonus, abs=1e-9)
        assert rec["scoring_function"] == pytest.approx(h.scoring_function, abs=1e-9)
        moved += s != rep["sequence"][0]
    assert moved > 0
    jb = loop.jobs()
    if jb["solved_step"][0] >= 0:
        assert jb["mfe_ss"][0] == inp.sec_struct
    loop.close()


def test_heterodimer_moves_match_host_mirror(engine):
    from desirna_b200 import design
    from desirna_b200.utils import sequence_utils as su
    from desirna_b200.utils import stats_inputs_outputs as sio
    inp = sio.make_input("complex", *HETERO)
    o = design.DesignOptions(replicas=2, RE_attempt=1, tm_max=0.9, tm_min=0.5, oligo_state="heterodimer")
    start = "AACUGAGGGGAAACCAA&GUCUAGUGACCACUCGUU"    # a shipped trajectory point: folds close to, not into, the target
    loop = design.DesignLoop([inp], o, seed=3, init_seqs=[start] * 2)
    cur_ss = loop.replicas()["mfe_ss"][0]
    assert "&" in cur_ss and cur_ss != inp.sec_struct
    N = 6000
    dev = [Counter() for _ in range(2)]
    for _ in range(N):
        for r, m in enumerate(loop.propose_only()):
            dev[r][m] += 1
    loop.close()
    nts = su.get_nt_list(inp)
    random.seed(8)
    for r in range(2):
        cur = SimpleNamespace(sequence=start, mfe_ss=cur_ss, temp_shelf=o.rep_temps_shelfs[r])
        host = Counter(su.propose_mutation(cur, nts, o, inp) for _ in range(N))
        assert set(dev[r]) <= set(host) | {k for k in dev[r] if dev[r][k] < 5}

        def where(counter):
            out = Counter()
            for m, c in counter.items():
                out[tuple(i for i in range(len(start)) if m[i] != start[i])] += c
            return out

        hw, dw = where(host), where(dev[r])
        assert abs(hw[()] - dw[()]) / N < 0.02          # the '&' itself is drawn equally often (a move that changes nothing)
        tv = 0.5 * sum(abs(hw[k] - dw[k]) for k in set(hw) | set(dw)) / N
        assert tv < 0.06, (r, tv)


@pytest.mark.parametrize("acgu,tm", [("on", "on"), ("off", "off")])
def test_move_generator_options_match_host_mirror(engine, acgu, tm):
    """-acgu on (weighted letters for paired positions) and -tm off (uniform positions): device draws vs host mirror"""
    from desirna_b200 import design
    from desirna_b200.utils import sequence_utils as su
    from desirna_b200.utils import stats_inputs_outputs as sio
    inp = sio.make_input("t", "((((....))))...((...))", "NNNNNNNNSNNNNNANNNNNNN")
    o = design.DesignOptions(replicas=2, RE_attempt=1, acgu_percentages=acgu, point_mutations=tm)
    start = "GGGAAAAACCCCAAAGGAAACC"
    loop = design.DesignLoop([inp], o, seed=21, init_seqs=[start] * 2)
    cur_ss = loop.replicas()["mfe_ss"][0]
    N = 5000
    dev = Counter()
    for _ in range(N):
        dev[loop.propose_only()[0]] += 1
    loop.close()
    nts = su.get_nt_list(inp)
    random.seed(6)
    cur = SimpleNamespace(sequence=start, mfe_ss=cur_ss, temp_shelf=o.rep_temps_shelfs[0])
    host = Counter(su.propose_mutation(cur, nts, o, inp) for _ in range(N))
    assert set(dev) <= set(host) | {k for k in dev if dev[k] < 5}

    def letters(counter):   # marginal over (first mutated position, its new letter): <= ~70 categories
        out = Counter()
        for m, c in counter.items():
            d = [i for i in range(len(start)) if m[i] != start[i]]
            out[(d[0], m[d[0]]) if d else ()] += c
        return out

    hl, dl = letters(host), letters(dev)
    tv = 0.5 * sum(abs(hl[k] - dl[k]) for k in set(hl) | set(dl)) / N
    assert tv < 0.09, tv
    gc_dev = sum(c for (k, c) in dl.items() if k and k[1] in "CG")
    gc_host = sum(c for (k, c) in hl.items() if k and k[1] in "CG")
    assert abs(gc_dev - gc_host) / N < 0.03   # -acgu on: C/G weighted 30:15 against A/U at paired positions, on both sides


def test_fixed_sequence_and_single_replica(engine):
    """nothing mutable: the loop keeps scoring the same sequence; one replica: no exchange partner"""
    from desirna_b200 import design
    from desirna_b200.utils import stats_inputs_outputs as sio
    fixed = sio.make_input("fixed", "((((....))))", "GGGGAAAACCCC")
    free = sio.make_input("free", "((((....))))")
    o = design.DesignOptions(replicas=1, RE_attempt=10)
    assert o.rep_temps_shelfs == [150.0]
    random.seed(0)
    loop = design.DesignLoop([fixed, free], o, seed=1)
    loop.run(3)
    rep = loop.replicas()
    assert rep["sequence"][0] == "GGGGAAAACCCC" and rep["mfe_ss"][0] == "((((....))))"
    assert rep["rec"][0][8] == 0 and loop.jobs()["solved_step"][0] == 0       # solved by its start sequence, recorded at step 0
    assert (rep["shelf"] == 0).all()
    assert rep["counts"][1][0] + rep["counts"][1][2] == 30
    loop.close()


def test_trajectory_and_result_files(engine, tmp_path):
    """design_batch(trajectory=True) yields what DesiRNA.py keeps as simulation_data; the reference-format writers run on it"""
    from desirna_b200 import design
    from desirna_b200.utils import stats_inputs_outputs as sio
    inputs = small_inputs(max_len=30, limit=3)
    o = design.DesignOptions(replicas=4, RE_attempt=25)
    results, info = design.design_batch(inputs, o, global_steps=3, seed=2, stop_when_solved=False, trajectory=True)
    for k, inp in enumerate(inputs):
        data, st = info["simulation_data"][k], info["stats"][k]
        assert len(data) == 4 * (3 + 1)                      # start records + one per replica per global step
        assert sorted({d["sim_step"] for d in data}) == [0, 25, 50, 75]
        assert {d["replica_num"] for d in data} == {1, 2, 3, 4}
        assert all(d["temp_shelf"] in o.rep_temps_shelfs for d in data)
        assert st.acc_mc_step + st.rej_mc_step == 4 * 75 and st.step == 75 and st.global_step == 3
        assert st.acc_re_step + st.rej_re_step == 2 + 1 + 2      # odd steps: pairs (0,1),(2,3); even step: pair (1,2)
        so = SimpleNamespace(infile=inp.name, outname=str(tmp_path / ("run%d" % k)), num_results=10, oligo="off", dimer="off", subopt="off", timlim=60)
        best, solved = sio.parse_and_output_results(data, inp, st, info["seconds"], so, "now")
        assert solved == any(d["mcc"] == 0.0 for d in data)
        assert best[0]["mcc"] == min(d["mcc"] for d in data)
        assert (tmp_path / ("run%d_traj.csv" % k)).read_text().count("\n") == len(data) + 1


HOMO_DIFF = ("((((....((((.....&))))....)))).....", "NNNNNNNNNNNNNNNNN&NNNNNNNNNNNNNNNNN")    # example_files/inputs/Homodimer_design_input.txt
HOMO_SAME = ("((((....))))..&((((....))))..", "NNNNNNNNNNNNNN&NNNNNNNNNNNNNN")
HOMO_INTRA = ("((((....))))..((((&))))..((((....))))", "NNNNNNNNNNNNNNNNNN&NNNNNNNNNNNNNNNNNN")    # helices inside each strand, halves differ


@pytest.mark.parametrize("case", [HOMO_DIFF, HOMO_SAME, HOMO_INTRA])
def test_homodimer_jobs_on_the_device(engine, case):
    """-d on (utils/sequence_utils.py:1102-1128, utils/energy_scores.py:120-125): mirrored moves in distribution against the host
    mirror, score records (incl. the dimer / monomer fraction term) against the host scoring"""
    from desirna_b200 import design
    from desirna_b200.utils import energy_scores as es
    from desirna_b200.utils import sequence_utils as su
    from desirna_b200.utils import stats_inputs_outputs as sio
    inp = sio.make_input("homodimer", *case)
    half = len(case[0].split("&")[0])
    o = design.DesignOptions(replicas=4, RE_attempt=15, oligo_state="homodimer", tm_max=0.8, tm_min=0.3)
    random.seed(3)
    start = "GGGGAAAACCCCAAGGU"[:half]
    start = start + "A" * (half - len(start))
    loop = design.DesignLoop([inp], o, seed=5, init_seqs=[start + "&" + start] * 4)
    cur_ss = loop.replicas()["mfe_ss"][0]
    N = 4000
    dev = Counter()
    for _ in range(N):
        dev[loop.propose_only()[0]] += 1
    nts = su.get_nt_list(inp)
    cur = SimpleNamespace(sequence=start + "&" + start, mfe_ss=cur_ss, temp_shelf=o.rep_temps_shelfs[0])
    host = Counter(su.propose_mutation(cur, nts, o, inp) for _ in range(N))
    same = case[0].split("&")[0] == case[0].split("&")[1]
    for m in dev:
        a, b = m.split("&")
        # identical target halves: the strands stay identical; different halves: only the two letters of a pair are mirrored,
        # an unpaired position may differ between the strands -- in the reference as here
        assert a == b or not same, m
    assert set(dev) <= set(host) | {k for k in dev if dev[k] < 5}

    def where(counter):
        out = Counter()
        for m, c in counter.items():
            out[tuple(i for i in range(len(m)) if m[i] != cur.sequence[i])] += c
        return out

    hw, dw = where(host), where(dev)
    tv = 0.5 * sum(abs(hw[k] - dw[k]) for k in set(hw) | set(dw)) / N
    assert tv < 0.07, tv
    # a few global steps, then the records against the host scoring of the same sequences
    loop.run(3)
    rep = loop.replicas()
    ref = es.score_sequences(rep["sequence"], inp, o)
    for g, (s, h) in enumerate(zip(rep["sequence"], ref)):
        rec = dict(zip(design.REC_FIELDS, rep["rec"][g]))
        a, b = s.split("&")
        assert (a == b or not same) and rep["mfe_ss"][g] == h.mfe_ss
        assert rec["edesired"] == h.edesired and abs(rec["Epf"] - h.Epf) <= 4e-6
        assert rec["oligo_fraction"] == pytest.approx(h.oligo_fraction, rel=1e-6)
        assert rec["oligomer_bonus"] == pytest.approx(h.oligomer_bonus, abs=1e-6)
        assert rec["scoring_function"] == pytest.approx(h.scoring_function, abs=1e-5)
    loop.close()
