"""GPU parity, second set: the holes the round-1 review listed (VERDICT.md "Next round" #1).

(a) random two-strand folds L/2 & L/2 and `avoid`-mode seq&seq against the oracle, all five partition-function columns;
(b) base-pair probabilities / ensemble defect at L = 200 and 300 (BASELINE config 5);
(c) the sweep's full batches (4096 x L=100, 1024 x L=400) row by row against the oracle;
(d) the pseudoknot overlay of the reference's G4 trajectories reproduced through the CUDA engine (brackets kept);
(e) a synthetic "Turner-2004-shaped" parameter file (tri-/hexaloops, MLintern < 0, mismatch_interior_1n != mismatch_interior)
    loaded by both the engine's and the oracle's loader and folded on random sequences.
"""
import os
import re
import struct
import types

import numpy as np
import pytest

from conftest import PAR1999, load_golden, synthetic_t2004_shaped_par

pytestmark = pytest.mark.gpu


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def rand_seqs(seed, B, L):
    rng = np.random.default_rng(seed)
    return ["".join("ACGU"[x] for x in row) for row in rng.integers(0, 4, (B, L))]


def close(a, b, rel=1e-6):
    return abs(a - b) <= rel * max(1.0, abs(b))


# ------------------------------------------------------------------------------------------------ (a)
@pytest.mark.parametrize("L,B", [(50, 96), (100, 64), (200, 24), (400, 8)])
def test_two_strand_sweep_vs_oracle(engine, oracle, L, B):
    """SURVEY 8(d) cofold sweep: L/2 + L/2 strands.  MFE energy and structure bit-exact, FA / FB / FcAB / FAB / F0AB to 1e-6
    relative, Ed of the MFE structure == MFE."""
    h = L // 2
    seqs = [s[:h] + "&" + s[h:] for s in rand_seqs(20241000 + L, B, L)]
    # a few uneven cuts, a cut after the first and before the last nucleotide
    seqs[0] = seqs[0].replace("&", "")[:1] + "&" + seqs[0].replace("&", "")[1:]
    seqs[1] = seqs[1].replace("&", "")[:-1] + "&" + seqs[1].replace("&", "")[-1:]
    seqs[2] = seqs[2].replace("&", "")[:L // 3] + "&" + seqs[2].replace("&", "")[L // 3:]
    out = engine.score_batch(seqs, want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF)
    tg = []
    for k, s in enumerate(seqs):
        e, ss = oracle.mfe(s)
        pf = oracle.pf(s)
        assert out["mfe_dcal"][k] == e, (L, k)
        assert out["mfe_ss"][k] == ss, (L, k)
        for c in range(5):
            assert close(out["pf"][k, c], pf[c]), (L, k, c, out["pf"][k, c], pf[c])
        tg.append([ss])
    ev = engine.score_batch(seqs, tg, want=engine.WANT_EVAL)["eval_dcal"][:, 0]
    assert (ev == out["mfe_dcal"]).all()


@pytest.mark.parametrize("L", [36, 100, 150])
def test_avoid_mode_self_dimer_vs_oracle(engine, oracle, L):
    """`-o avoid` folds seq&seq at 2N (energy_scores.py:412-419) and takes the dimer fraction from FcAB - FA - FB."""
    from desirna_b200 import RNA
    from desirna_b200.utils import dimer_multichain_energy as dme
    mono = rand_seqs(99 + L, 12, L)
    seqs = [s + "&" + s for s in mono]
    out = engine.score_batch(seqs, want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF)
    for k, s in enumerate(seqs):
        e, ss = oracle.mfe(s)
        pf = oracle.pf(s)
        assert out["mfe_dcal"][k] == e and out["mfe_ss"][k] == ss
        for c in range(5):
            assert close(out["pf"][k, c], pf[c]), (L, k, c)
        # the two monomer energies of a self-dimer are one and the same ensemble
        assert close(out["pf"][k, 0], out["pf"][k, 1], 1e-12)
        fa = oracle.pf(mono[k])[4]
        assert close(out["pf"][k, 0], fa)
    fc = RNA.fold_compound(seqs[0])
    f = dme.oligo_fraction(seqs[0], fc)
    assert 0.0 <= f <= 1.0


# ------------------------------------------------------------------------------------------------ (b)
@pytest.mark.parametrize("L,B", [(200, 12), (300, 8)])
def test_bpp_and_defect_long_vs_oracle(engine, oracle, L, B):
    """BASELINE config 5 runs Edef at L = 300.  Tolerance of north_star: 1e-5 absolute (met with 1e-9)."""
    seqs = rand_seqs(5151 + L, B, L)
    mfe, ss, epf, ed = oracle.fold_batch(seqs, nthreads=8)
    out = engine.score_batch(seqs, [[s] for s in ss], want=engine.WANT_MFE | engine.WANT_PF | engine.WANT_BPP | engine.WANT_DEFECT)
    for k, s in enumerate(seqs):
        pf, bpp = oracle.pf(s, bpp=True)
        assert close(out["pf"][k, 4], pf[4])
        got = out["bpp"][k][:L, :L]
        assert np.abs(got - bpp).max() < 1e-8, (L, k, np.abs(got - bpp).max())
        assert abs(out["defect"][k] - oracle.ensemble_defect(bpp, ss[k])) < 1e-8
    # defect alone (no bpp matrix leaves the device) gives the same numbers
    d2 = engine.score_batch(seqs, [[s] for s in ss], want=engine.WANT_DEFECT)["defect"]
    assert np.abs(d2 - out["defect"]).max() < 1e-12


# ------------------------------------------------------------------------------------------------ (c)
@pytest.mark.parametrize("L,B", [(100, 4096), (400, 1024)])
def test_full_size_batch_row_by_row_vs_oracle(engine, oracle, L, B):
    """BASELINE config 2's batch, every row against the oracle: MFE energy and structure bit-exact, ensemble free energy
    1e-6 relative, Ed of the oracle's structure bit-exact.  (The oracle needs ~1 s at L=100 and ~10-20 s at L=400 on the box.)"""
    seqs = rand_seqs(20240000 + L, B, L)
    mfe, ss, epf, ed = oracle.fold_batch(seqs, targets=None, nthreads=os.cpu_count() or 8)
    out = engine.score_batch(seqs, [[s] for s in ss], want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF | engine.WANT_EVAL)
    assert (out["mfe_dcal"] == mfe).all()
    assert out["mfe_ss"] == ss
    assert (out["eval_dcal"][:, 0] == mfe).all()
    rel = np.abs(out["pf"][:, 4] - epf) / np.maximum(1.0, np.abs(epf))
    assert rel.max() <= 1e-6, rel.max()


# ------------------------------------------------------------------------------------------------ (d)
def _options(**kw):
    o = types.SimpleNamespace(oligo_state="none", pks="off", scoring_f=[("Ed-Epf", 1.0)], subopt="off", motifs={})
    o.__dict__.update(kw)
    return o


def _input_file(target, alts=()):
    return types.SimpleNamespace(sec_struct=target, alt_sec_struct=(list(alts) or None), alt_sec_structs=list(alts) or None)


def test_pseudoknot_overlay_goldens_through_the_cuda_engine(engine):
    """G4 = the reference's pseudoknot example run (pks on): mfe_ss carries the [] / <> / {} overlay of
    sequence_utils.py:1166-1228.  Reproduced with the batched overlay (masked refolds of all sequences per round)
    on the CUDA engine; Ed bit-exact, Epf to float32, every similarity score and the scoring function."""
    from desirna_b200 import RNA
    from desirna_b200.utils import energy_scores as es
    import oracle_backend
    oracle_backend.install(engine)   # the product's engine module (another test may have swapped the seam)
    rows = load_golden("G4")
    assert any("[" in r["mfe_ss"] for r in rows)
    by_target = {}
    for r in rows:
        by_target.setdefault((r["target"], tuple(r["alts"])), []).append(r)
    opt = _options(pks="on")
    n = 0
    for (target, alts), grp in by_target.items():
        scored = es.score_sequences([r["sequence"] for r in grp], _input_file(target, alts), opt)
        for s, r in zip(scored, grp):
            assert s.mfe_ss == r["mfe_ss"], r["sequence"]
            assert s.edesired == r["Ed"]
            assert abs(s.Epf - r["Epf"]) <= 2.5e-6
            assert abs(s.mcc - r["one_minus_mcc"]) <= 1e-12 and abs(s.recall - r["one_minus_recall"]) <= 1e-12
            assert abs(s.precision - r["one_minus_precision"]) <= 1e-12
            assert abs(s.scoring_function - r["scoring_function"]) <= 5e-6
            n += 1
    assert n == len(rows)
    # the one-sequence path paints the same overlay
    r = next(r for r in rows if "[" in r["mfe_ss"])
    one = es.score_sequence(r["sequence"], _input_file(r["target"], r["alts"]), opt)
    assert one.mfe_ss == r["mfe_ss"]


# ------------------------------------------------------------------------------------------------ (e)
@pytest.fixture
def t2004_shaped(engine, tmp_path):
    from oracle.pyoracle import Oracle
    p = str(tmp_path / "synthetic_t2004_shaped.par")
    synthetic_t2004_shaped_par(p)
    engine.params_load(p)
    try:
        yield Oracle(p)
    finally:
        engine.params_builtin(1999)


def test_synthetic_t2004_shaped_parameters(engine, t2004_shaped):
    O = t2004_shaped
    for name, idx in (("MLintern", ()), ("MLclosing", ()), ("ninio_m", ()), ("n_tri", ()), ("n_hexa", ()), ("mm1nI", (3, 2, 4)), ("dangle5", (1, 3))):
        assert engine.params_get(name, *idx) == O.get(name, *idx), name
    assert engine.params_get("MLintern") == -90 and engine.params_get("n_tri") == 4 and engine.params_get("n_hexa") == 5
    rng = np.random.default_rng(7)
    seqs = []
    for L in (40, 80, 120, 200):
        seqs += rand_seqs(2004 + L, 24 if L < 200 else 8, L)
    # sequences that must close the tabulated tri- and hexaloops, and two-strand rows
    seqs += ["GGGGGAAACCCCC", "GGGGC" + "CAACG" + "GCCCC", "GGGG" + "GUUAC" + "CCCC" + "AAAA" + "GGGA" + "CAGUACU"[0:0] + "ACAGUACU" + "UCCC",
             "GGGC" + "ACAGUGAU" + "GCCC", "GGCGAAAAAACGCC", "GGGGC" + "CAACG" + "GCCCC&GGGGCGAAACGCCCC", "GGGAGAAAACUCC&GGAGUUUUUCCC"]
    out = engine.score_batch(seqs, want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF)
    tg = []
    for k, s in enumerate(seqs):
        e, ss = O.mfe(s)
        pf = O.pf(s)
        assert out["mfe_dcal"][k] == e, s
        assert out["mfe_ss"][k] == ss, s
        for c in ((0, 1, 2, 3, 4) if "&" in s else (4,)):
            assert close(out["pf"][k, c], pf[c]), (s, c)
        tg.append(ss)
    ev = engine.score_batch(seqs, [[t] for t in tg], want=engine.WANT_EVAL)["eval_dcal"][:, 0]
    assert (ev == out["mfe_dcal"]).all()
    for k, s in enumerate(seqs):
        assert ev[k] == O.eval(s, tg[k])
    # outside pass under the same tables
    short = [s for s in seqs if "&" not in s and len(s) <= 80][:16]
    o2 = engine.score_batch(short, [[O.mfe(s)[1]] for s in short], want=engine.WANT_BPP | engine.WANT_DEFECT)
    for k, s in enumerate(short):
        pf, bpp = O.pf(s, bpp=True)
        n = len(s)
        assert np.abs(o2["bpp"][k][:n, :n] - bpp).max() < 1e-9


# ------------------------------------------------------------------------------------------------ third-generation fill kernels
def test_fill3_ragged_batch_vs_oracle(engine, oracle, monkeypatch):
    """bf_fill3.cu takes batches of more than one sequence per SM (the sweep): a ragged batch (lengths 1..110 in one stride,
    consecutive sequences of very different length on the same CTA) with hard constraints on a third of the rows."""
    rng = np.random.default_rng(303)
    lens = list(rng.integers(1, 111, 300)) + [1, 2, 3, 4, 5, 6, 110, 110, 7, 110]
    seqs = ["".join("ACGU"[x] for x in rng.integers(0, 4, int(L))) for L in lens]
    stride = max(lens)
    mask = np.zeros((len(seqs), stride), np.uint8)
    for k in range(0, len(seqs), 3):
        mask[k, :lens[k]] = rng.random(lens[k]) < 0.25
    out = engine.score_batch(seqs, nopair=mask, want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF)
    for k, s in enumerate(seqs):
        e, ss = oracle.mfe(s, nopair=mask[k, :len(s)] if k % 3 == 0 else None)
        assert out["mfe_dcal"][k] == e and out["mfe_ss"][k] == ss, (k, len(s))
        f = oracle.pf(s)[4]
        assert close(out["pf"][k, 4], f), (k, len(s))


@pytest.mark.parametrize("L", [60, 100, 128, 150])
def test_fill3_equals_round1_kernels(engine, monkeypatch, L):
    """the two generations of fill kernels on the same batch: identical integers, ensemble energies to 1e-12"""
    seqs = rand_seqs(808 + L, 200, L)
    want = engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF
    monkeypatch.setenv("BF_FILL3", "1")
    new = engine.score_batch(seqs, want=want)
    tc = engine.debug_table(0, len(seqs)).copy()
    tf = engine.debug_table(1, len(seqs)).copy()
    monkeypatch.setenv("BF_FILL3", "0")
    old = engine.score_batch(seqs, want=want)
    assert (new["mfe_dcal"] == old["mfe_dcal"]).all() and new["mfe_ss"] == old["mfe_ss"]
    assert np.allclose(new["pf"][:, 4], old["pf"][:, 4], rtol=1e-12, atol=0)
    # the DP tables the backtrack and the suboptimal walk read: same c and fML, cell by cell
    assert (engine.debug_table(0, len(seqs)) == tc).all()
    assert (engine.debug_table(1, len(seqs)) == tf).all()


def test_fill3_outside_pass_uses_its_tables(engine, oracle, monkeypatch):
    """base-pair probabilities on top of the third-generation inside pass (qb, qm, qm1 per sequence): B > SM count"""
    seqs = rand_seqs(909, 160, 70)
    mfe, ss, epf, ed = oracle.fold_batch(seqs, nthreads=8)
    out = engine.score_batch(seqs, [[s] for s in ss], want=engine.WANT_BPP | engine.WANT_DEFECT)
    for k in range(0, 160, 9):
        pf, bpp = oracle.pf(seqs[k], bpp=True)
        assert np.abs(out["bpp"][k][:70, :70] - bpp).max() < 1e-9
        assert abs(out["defect"][k] - oracle.ensemble_defect(bpp, ss[k])) < 1e-9


@pytest.mark.parametrize("L", [36, 100, 148, 176, 200])
def test_small_batch_fill3_variants(engine, oracle, monkeypatch, L):
    """batches that leave SMs idle take the 16-warp third-generation kernels (one CTA per sequence with the whole shared memory of its
    SM; MFE while the fML table fits on chip, partition function up to ~178 nt); BF_FILL3_SMALL=0 puts the 16-warp round-1 kernels
    back.  Same integers, ensemble energies to 1e-12, and the oracle on every fifth sequence; ragged lengths in one batch."""
    seqs = rand_seqs(4242 + L, 40, L) + rand_seqs(4243 + L, 8, max(5, L // 3))
    want = engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF
    monkeypatch.setenv("BF_CL", "0")
    monkeypatch.setenv("BF_FILL3_SMALL", "1")
    new = engine.score_batch(seqs, want=want)
    monkeypatch.setenv("BF_FILL3_SMALL", "0")
    old = engine.score_batch(seqs, want=want)
    assert (new["mfe_dcal"] == old["mfe_dcal"]).all() and new["mfe_ss"] == old["mfe_ss"]
    assert np.allclose(new["pf"][:, 4], old["pf"][:, 4], rtol=1e-12, atol=0)
    for k in range(0, len(seqs), 5):
        e, ss = oracle.mfe(seqs[k])
        assert new["mfe_dcal"][k] == e and new["mfe_ss"][k] == ss, (L, k)
        assert close(new["pf"][k, 4], oracle.pf(seqs[k])[4]), (L, k)


@pytest.mark.parametrize("L", [70, 120])
def test_small_batch_outside_pass_on_fill3_tables(engine, oracle, L):
    """base-pair probabilities and ensemble defect of a small batch: the outside pass reads the qb / qm / qm1 tables the 16-warp
    third-generation inside kernel leaves per sequence"""
    seqs = rand_seqs(5151 + L, 12, L)
    mfe, ss, epf, ed = oracle.fold_batch(seqs, nthreads=8)
    out = engine.score_batch(seqs, [[s] for s in ss], want=engine.WANT_BPP | engine.WANT_DEFECT)
    for k in range(0, 12, 3):
        pf, bpp = oracle.pf(seqs[k], bpp=True)
        assert np.abs(out["bpp"][k][:L, :L] - bpp).max() < 1e-9
        assert abs(out["defect"][k] - oracle.ensemble_defect(bpp, ss[k])) < 1e-9
