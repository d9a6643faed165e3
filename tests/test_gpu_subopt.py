"""GPU parity of the suboptimal-structure walk (bf_subopt / fold_compound.subopt_cb) against brute-force enumeration.

Reference call sites: utils/energy_scores.py:453-488 (get_first_suboptimal_structure_and_energy, uniq_ML = 1),
utils/sequence_utils.py:783.  ViennaRNA values for subopt are pinned by nothing the reference ships, so the
check is exhaustive enumeration under the validated loop model (oracle) for short sequences, and energy consistency
(every listed structure re-evaluates to its listed energy, no duplicates, first = MFE) for longer ones."""
import numpy as np
import pytest

import oracle_backend

pytestmark = pytest.mark.gpu


def rand_seq(rng, n):
    return "".join("ACGU"[x] for x in rng.integers(0, 4, n))


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_band_equals_brute_force(engine, oracle, seed):
    rng = np.random.default_rng(seed)
    checked = 0
    for n in [8, 10, 12, 13, 14, 15, 16, 17, 18]:
        for _ in range(3):
            s = rand_seq(rng, n) if rng.random() < 0.5 else "GGG" + rand_seq(rng, n - 6) + "CCC"
            delta = int(rng.choice([0, 100, 300, 600]))
            found, trunc = engine.subopt(s, delta, max_out=20000)
            mfe = oracle.mfe(s)[0]
            want = oracle.enumerate_band(s, mfe + delta)
            assert not trunc
            assert [e for _, e in found] == [e for e, _ in want], (s, delta)
            assert sorted(ss for ss, _ in found) == sorted(ss for _, ss in want), (s, delta)
            checked += len(want)
    assert checked > 50


def test_long_sequence_consistency(engine, oracle):
    rng = np.random.default_rng(11)
    for n in [60, 120, 300]:
        s = rand_seq(rng, n)
        found, trunc = engine.subopt(s, 100, max_out=5000)
        mfe, ss = oracle.mfe(s)
        assert found and found[0][1] == mfe
        assert any(x == ss for x, e in found if e == mfe)
        assert len({x for x, _ in found}) == len(found)          # each structure once
        for x, e in found[:200]:
            assert oracle.eval(s, x) == e                          # the walk's energy is the structure's loop-sum energy
            assert e <= mfe + 100


def test_truncation_flag(engine):
    s = "GGGGGAAAAACCCCCAAAAAGGGGGAAAAACCCCC"
    full, trunc = engine.subopt(s, 500, max_out=20000)
    assert not trunc and len(full) > 5
    part, trunc = engine.subopt(s, 500, max_out=5)
    assert trunc and len(part) == 5


def test_reference_helper_through_the_shim(engine, oracle):
    """get_first_suboptimal_structure_and_energy(seq, fc, 1) -> second-best structure and its energy (float32 kcal/mol)."""
    import struct
    from desirna_b200 import RNA
    from desirna_b200.utils import energy_scores as es
    oracle_backend.install(engine)
    s = "GGGAAAUCCCGCGAAAGC"
    fc = RNA.fold_compound(s)
    ss2, e2 = es.get_first_suboptimal_structure_and_energy(s, fc, 1)
    mfe = oracle.mfe(s)[0]
    band = oracle.enumerate_band(s, mfe + 5000)
    assert e2 == struct.unpack("f", struct.pack("f", band[1][0] / 100.0))[0]
    assert oracle.eval(s, ss2) == band[1][0]


def test_subopt_cb_is_complete_when_the_band_overflows_the_first_buffer(engine):
    """The host walk stops at max_out structures in SEARCH order; the shim must not hand a truncated band to
    get_first_suboptimal_structure_and_energy (the k-th best of a truncated band is not the k-th best).  A flat landscape:
    grow delta until 4096 structures no longer hold the band, then the shim's callback count must equal the full band."""
    from desirna_b200 import RNA
    oracle_backend.install(engine)
    rng = np.random.default_rng(4321)
    s = "".join("ACGU"[x] for x in rng.integers(0, 4, 160))
    delta = 100
    while True:
        part, trunc = engine.subopt(s, delta, max_out=4096)
        if trunc:
            break
        delta += 100
        assert delta < 3000
    full, trunc_full = engine.subopt(s, delta, max_out=1 << 18)
    assert not trunc_full and len(full) > 4096
    got = []
    RNA.fold_compound(s).subopt_cb(delta, lambda ss, e, data: got.append((ss, e)), None)
    assert got[-1][0] is None
    got = got[:-1]
    assert len(got) == len(full) and len({g[0] for g in got}) == len(got)
    mfe = full[0][1]
    assert min(e for _, e in full) == mfe and max(e for _, e in full) <= mfe + delta
    assert sorted(e for _, e in full)[:10] == [e for _, e in full[:10]]


# ------------------------------------------------------------------------------------------------ second-best structure by a 2-best DP
@pytest.mark.parametrize("seed", [4, 5])
def test_second_best_equals_brute_force(engine, oracle, seed):
    """bf_second_best (csrc/bf_twobest.cu: one DP over (best, second best) pairs on the unambiguous grammar) against exhaustive
    enumeration of every structure of short sequences: e1 = the lowest energy, e2 = the second entry of the sorted list, ties
    counted as two structures -- what get_first_suboptimal_structure_and_energy(seq, fc, 1)[1] reads off subopt_cb
    (utils/energy_scores.py:453-488)."""
    rng = np.random.default_rng(seed)
    seqs = []
    for n in [6, 8, 9, 10, 12, 13, 14, 15, 16, 17, 18, 19]:
        for _ in range(3):
            seqs.append(rand_seq(rng, n) if rng.random() < 0.5 else "GGG" + rand_seq(rng, n - 6) + "CCC")
    seqs += ["AAAAAAAAAA", "GGGAAACCC", "GGGGAAAACCCC"]
    e1, e2 = engine.second_best(seqs)
    some = 0
    for k, s in enumerate(seqs):
        want = oracle.enumerate_band(s, 10 ** 6)          # every structure, sorted by energy
        assert want[0][0] == e1[k] == oracle.mfe(s)[0], s
        if len(want) >= 2:
            assert want[1][0] == e2[k], (s, want[:3], int(e2[k]))
            some += 1
        else:
            assert e2[k] >= 10000000, s
    assert some > 30


def test_second_best_long_sequences_and_constraints(engine):
    """longer sequences: e1 is the MFE of the fill kernels, e2 the second energy of the band walk (bf_subopt); with hard constraints too"""
    rng = np.random.default_rng(17)
    seqs = [rand_seq(rng, n) for n in (40, 75, 120, 200)]
    nopair = np.zeros((len(seqs), 200), np.uint8)
    nopair[1, 10:30] = 1
    nopair[3, ::7] = 1
    e1, e2 = engine.second_best(seqs, nopair)
    for k, s in enumerate(seqs):
        mask = nopair[k, :len(s)] if nopair[k].any() else None
        band, trunc = engine.subopt(s, int(e2[k] - e1[k]), mask, max_out=200000)
        assert not trunc
        en = sorted(e for _, e in band)
        assert en[0] == e1[k] and en[1] == e2[k], (len(s), en[:3], int(e1[k]), int(e2[k]))


def test_negative_design_term_uses_the_two_best_dp(engine, oracle):
    """get_first_suboptimal_energy == get_first_suboptimal_structure_and_energy(...)[1] (the enumeration) on the shim"""
    from desirna_b200 import RNA
    from desirna_b200.utils import energy_scores as es
    rng = np.random.default_rng(23)
    for n in (20, 36, 60):
        s = "GGGG" + rand_seq(rng, n - 8) + "CCCC"
        fast = es.get_first_suboptimal_energy(s, RNA.fold_compound(s))
        slow = es.get_first_suboptimal_structure_and_energy(s, RNA.fold_compound(s), 1)[1]
        assert fast == slow, (s, fast, slow)
    assert es.get_first_suboptimal_energy("A" * 20, RNA.fold_compound("A" * 20)) == 0   # a single structure: the reference returns 0
