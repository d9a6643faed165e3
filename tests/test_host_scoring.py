"""CPU: the host-side mirror of utils/energy_scores.py / sim_score.py / dimer_multichain_energy.py against the
reference's shipped trajectory values (tests/golden), with the CPU oracle standing in for the GPU engine."""
import re
import types

import pytest

import oracle_backend

from conftest import PAR1999, load_golden


@pytest.fixture(scope="module")
def rna_on_oracle(oracle):
    from desirna_b200 import RNA
    from oracle_backend import OracleBackend
    old = oracle_backend.current()
    be = OracleBackend(PAR1999)
    oracle_backend.install(be)
    yield be
    oracle_backend.install(old)


def options(**kw):
    o = types.SimpleNamespace(oligo_state="none", pks="off", scoring_f=[("Ed-Epf", 1.0)], subopt="off", motifs={})
    o.__dict__.update(kw)
    return o


def input_file(target, alts=()):
    return types.SimpleNamespace(sec_struct=target, alt_sec_struct=(list(alts) or None), alt_sec_structs=list(alts) or None)


FIELDS = [("Ed", "edesired", 0.0), ("Epf", "Epf", 2.5e-6), ("Ed_minus_Epf", "edesired_minus_Epf", 2.5e-6),
          ("one_minus_mcc", "mcc", 1e-12), ("one_minus_recall", "recall", 1e-12), ("one_minus_precision", "precision", 1e-12)]


@pytest.mark.parametrize("tag,pks,stride", [("G1", "off", 7), ("G2", "on", 13), ("G3", "on", 1), ("G4", "on", 7)])
def test_score_sequence_single_chain(rna_on_oracle, tag, pks, stride):
    from desirna_b200.utils import energy_scores as es
    rows = load_golden(tag)[::stride]
    opt = options(pks=pks)
    bad_ss = 0
    for r in rows:
        s = es.score_sequence(r["sequence"], input_file(r["target"], r["alts"]), opt)
        if s.mfe_ss != r["mfe_ss"]:
            bad_ss += 1     # the one tie anomaly of the 2023 run lives in G2
            continue
        for gk, attr, tol in FIELDS:
            assert abs(getattr(s, attr) - r[gk]) <= tol, (tag, r["sequence"], gk, getattr(s, attr), r[gk])
        if r["alts"]:
            assert abs(s.edesired2 - r["Ed2"]) < 1e-9
        assert abs(s.scoring_function - r["scoring_function"]) <= 5e-6, (tag, r["sequence"])
    assert bad_ss <= (1 if tag == "G2" else 0)


@pytest.mark.parametrize("tag,state", [("G5", "heterodimer"), ("G6", "homodimer")])
def test_score_sequence_two_chains(rna_on_oracle, tag, state):
    from desirna_b200.utils import energy_scores as es
    rows = load_golden(tag)[::9]
    for r in rows:
        s = es.score_sequence(r["sequence"], input_file(r["target"]), options(oligo_state=state))
        assert s.mfe_ss == r["mfe_ss"]
        for gk, attr, tol in FIELDS:
            assert abs(getattr(s, attr) - r[gk]) <= tol, (tag, r["sequence"], gk)
        # the Aug-2023 runs predate the oligomer bonus (energy_scores.py:109-116): their score is Ed-Epf alone
        assert abs((s.scoring_function - s.oligomer_bonus) - r["scoring_function"]) <= 5e-6
        assert 0.0 < s.oligo_fraction < 1.0


def test_batched_equals_one_by_one(rna_on_oracle):
    from desirna_b200.utils import energy_scores as es
    rows = load_golden("G3")[:12]
    inp = input_file(rows[0]["target"], rows[0]["alts"])
    opt = options(pks="on", scoring_f=[("Ed-Epf", 0.5), ("1-MCC", 0.5), ("sln_Epf", 0.1), ("Ed-MFE", 0.2), ("1-precision", 0.1), ("1-recall", 0.1)],
                  motifs={"GNRA": (re.compile("G[ACGU][AG]A"), -0.5)})
    calls0 = rna_on_oracle.calls
    batch = es.score_sequences([r["sequence"] for r in rows], inp, opt)
    calls_batched = rna_on_oracle.calls - calls0
    single = [es.score_sequence(r["sequence"], inp, opt) for r in rows]
    for a, b in zip(batch, single):
        assert vars(a) == vars(b)
    # one engine call for the folds + evals of the whole batch; the rest are pk-overlay refolds and Ed-MFE
    assert calls_batched < (rna_on_oracle.calls - calls0 - calls_batched)


def test_avoid_mode_and_edef(rna_on_oracle):
    from desirna_b200.utils import energy_scores as es
    r = load_golden("G1")[3]
    s = es.score_sequence(r["sequence"], input_file(r["target"]), options(oligo_state="avoid", scoring_f=[("Edef", 1.0)]))
    assert 0.0 <= s.ensemble_defect <= 1.0
    assert abs(s.scoring_function - (s.ensemble_defect + s.monomer_bonus)) < 1e-12
    assert s.monomer_bonus >= 0.0


def test_sim_score_matches_reference_semantics():
    from desirna_b200.utils.sim_score import SimScore, pairing_positions
    assert pairing_positions("((..))") == {0: 5, 1: 4, 2: -1, 3: -1, 4: 1, 5: 0}
    assert pairing_positions("([.)]") == {0: 3, 1: 4, 2: -1, 3: 0, 4: 1}
    # '&' -> "Ee": one extra always-matching pair (energy_scores.py:79)
    a = SimScore("((..Ee..))", "((..Ee..))")
    a.find_basepairs(); a.cofusion_matrix()
    assert a.conf_mat == (6, 0, 0, 4) and a.mcc() == 1.0
    b = SimScore("((((....))))", "............")
    b.find_basepairs(); b.cofusion_matrix()
    assert b.conf_mat == (0, 0, 8, 4) and b.mcc() == 0.0 and b.recall() == 0.0 and b.precision() == 0.0
    c = SimScore("............", "............")
    c.find_basepairs(); c.cofusion_matrix()
    assert c.mcc() == 1.0


def test_dimer_module_constants():
    from desirna_b200.utils import dimer_multichain_energy as dme
    assert (dme.KB, dme.RHO, dme.CONC) == (0.001987204259, 55.14, 1e-3) and abs(dme.TEMP - 310.15) < 1e-12
    f = 0.25
    assert abs(dme.kTlog_oligo_fraction(f) + dme.KB * dme.TEMP * __import__("math").log(f)) < 1e-15


def test_subopt_term_of_the_score(rna_on_oracle, oracle):
    """`-nd on`: when the MFE structure IS the target (mcc term 0) the energy gap to the second-best structure enters the score
    (energy_scores.py:104-107, :406-410): scoring_function -= (E_2nd - Epf)."""
    import struct
    from desirna_b200.utils import energy_scores as es
    seq = "GGGAAAUCCCGCGAAAGC"
    mfe, ss = oracle.mfe(seq)
    plain = es.score_sequence(seq, input_file(ss), options())
    nd = es.score_sequence(seq, input_file(ss), options(subopt="on"))
    assert plain.mcc == 0
    band = oracle.enumerate_band(seq, mfe + 5000)
    e2 = struct.unpack("f", struct.pack("f", band[1][0] / 100.0))[0]
    assert nd.subopt_e == e2
    assert abs(nd.scoring_function - (plain.scoring_function - (e2 - plain.Epf))) < 1e-9
