"""GPU parity of the cluster-per-sequence fill kernels (csrc/bf_cluster.cu; SURVEY section 8 "long context": the O(N^2) tables of one
sequence partitioned over the shared memories of a thread-block cluster, every cross-CTA transfer a DSMEM store).

They compute what the single-CTA kernels compute -- fc.mfe() / fc.pf() of the reference (utils/energy_scores.py:150-151) -- so they
are held to the same bar: MFE energies and structures bit for bit (minima are order-independent), ensemble energies to 1e-10
relative against the single-CTA kernels (sums are taken in another order) and to 1e-6 against the oracle.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ragged(seed, B, L):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(B):
        n = L if k < 2 else int(rng.integers(max(5, L // 3), L + 1))
        out.append("".join("ACGU"[x] for x in rng.integers(0, 4, n)))
    out[-1] = "A" * min(L, 12)     # nothing pairs
    out[-2] = ("GGGG" + "A" * 4 + "CCCC") * (L // 12) or "GGGAAACCC"
    return out


@pytest.mark.parametrize("L,C,BW", [(20, 4, 32), (47, 4, 16), (64, 8, 32), (100, 16, 32), (150, 4, 32), (150, 8, 16), (200, 8, 32),
                                    (256, 16, 32), (300, 4, 32), (300, 16, 16), (400, 8, 32), (400, 16, 32)])
def test_cluster_kernels_equal_single_cta_kernels(engine, monkeypatch, L, C, BW):
    """forced on (BF_CL=1) at one cluster size / block width against the default kernels on a ragged batch with hard constraints"""
    seqs = ragged(4100 + L + C, 20, L)
    rng = np.random.default_rng(L)
    nopair = np.zeros((len(seqs), max(len(s) for s in seqs)), np.uint8)
    nopair[::3] = rng.random((len(nopair[::3]), nopair.shape[1])) < 0.06
    want = engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF
    monkeypatch.setenv("BF_CL", "0")
    ref = engine.score_batch(seqs, nopair=nopair, want=want)
    monkeypatch.setenv("BF_CL", "1")
    monkeypatch.setenv("BF_CL_C", str(C))
    monkeypatch.setenv("BF_CL_BW", str(BW))
    got = engine.score_batch(seqs, nopair=nopair, want=want)
    assert np.array_equal(ref["mfe_dcal"], got["mfe_dcal"])
    assert list(ref["mfe_ss"]) == list(got["mfe_ss"])
    rel = np.abs(ref["pf"][:, 4] - got["pf"][:, 4]) / np.maximum(1.0, np.abs(ref["pf"][:, 4]))
    assert rel.max() < 1e-10, rel.max()


@pytest.mark.parametrize("L,B", [(170, 5), (260, 12), (400, 3)])
def test_small_batches_of_long_sequences_vs_oracle(engine, oracle, L, B):
    """the default dispatch (these batches go to the cluster kernels) against the oracle"""
    seqs = ragged(977 + L, B + 2, L)[:B]
    out = engine.score_batch(seqs, want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF)
    for k, s in enumerate(seqs):
        e, ss = oracle.mfe(s)
        assert out["mfe_dcal"][k] == e and out["mfe_ss"][k] == ss, (L, k)
        f = oracle.pf(s)[4]
        assert abs(out["pf"][k, 4] - f) <= 1e-6 * max(1.0, abs(f)), (L, k, out["pf"][k, 4], f)


def test_outside_pass_on_cluster_tables(engine, monkeypatch):
    """ensemble defect from qb / qm / qm1 written by the cluster partition-function kernel == from the single-CTA kernel's"""
    L = 180
    seqs = ragged(31, 6, L)[:4]
    tg = [["." * len(s)] for s in seqs]
    want = engine.WANT_MFE | engine.WANT_PF | engine.WANT_DEFECT
    monkeypatch.setenv("BF_CL", "0")
    ref = engine.score_batch(seqs, tg, want=want)
    monkeypatch.setenv("BF_CL", "1")
    got = engine.score_batch(seqs, tg, want=want)
    assert np.allclose(ref["defect"], got["defect"], rtol=0, atol=1e-9)
    assert np.allclose(ref["pf"][:, 4], got["pf"][:, 4], rtol=1e-10, atol=0)


def test_generic_kernels_on_single_strands_equal_the_fill_path(engine):
    """the generic kernels (csrc/bf_kernels.cu: the two-strand path, and the fallback for lengths the fill path does not cover) on
    plain single strands, forced by BF_FORCE_GENERIC=1 in a fresh interpreter (the switch is read at bf_init): same energies and
    structures as the fill kernels of this process"""
    import json
    import os
    import subprocess
    import sys
    seqs = ragged(5150, 12, 90)
    ref = engine.score_batch(seqs, want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF)
    code = ("import json, sys; sys.path.insert(0, %r); from desirna_b200 import engine; engine.init(0); engine.params_builtin(1999); "
            "o = engine.score_batch(%r, want=7); print(json.dumps({'mfe': o['mfe_dcal'].tolist(), 'ss': o['mfe_ss'], 'pf': o['pf'][:, 4].tolist()}))"
            % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), seqs))
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, BF_FORCE_GENERIC="1"), capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    got = json.loads(out.stdout.strip().split("\n")[-1])
    assert got["mfe"] == ref["mfe_dcal"].tolist() and got["ss"] == list(ref["mfe_ss"])
    assert np.allclose(got["pf"], ref["pf"][:, 4], rtol=1e-10, atol=0)


@pytest.mark.parametrize("C", [4, 8, 16])
def test_cluster_kernels_fold_the_eterna100_solutions_into_their_targets(engine, monkeypatch, C):
    """the reference's own Eterna100-V1 results (eterna_benchmark/Eterna100V1_benchmark_results, tests/golden/E1.jsonl: sequences
    ViennaRNA folds into their targets, 12..400 nt): forced through the cluster kernels at every cluster size, each still folds
    into its target -- a pin to ViennaRNA-produced data, not to this repo's other kernels"""
    from conftest import load_golden
    rows = [r for r in load_golden("E1") if len(r["target"]) >= 60]
    monkeypatch.setenv("BF_CL", "1")
    monkeypatch.setenv("BF_CL_C", str(C))
    for lo, hi in ((60, 130), (131, 260), (261, 400)):
        part = [r for r in rows if lo <= len(r["target"]) <= hi]
        out = engine.score_batch([r["sequence"] for r in part], [[r["target"]] for r in part],
                                 want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF | engine.WANT_EVAL)
        for k, r in enumerate(part):
            assert out["mfe_ss"][k] == r["target"], (r["file"], C)
            assert out["eval_dcal"][k, 0] == out["mfe_dcal"][k] and out["pf"][k, 4] <= out["mfe_dcal"][k] / 100.0 + 1e-9
