"""Checker behind test_snake_graphs_and_moves_match_the_reference_draw_for_draw (run with PYTHONHASHSEED=0, see there)."""
import json
import os
import random
import sys
from types import SimpleNamespace

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from desirna_b200.utils import sequence_utils as su            # noqa: E402
from desirna_b200.utils import stats_inputs_outputs as sio     # noqa: E402


def main():
    opt = SimpleNamespace(acgu_percentages="off", nt_percentages={"A": 15, "C": 30, "G": 30, "U": 15}, point_mutations="off",
                          oligo_state="none", tm_max=0.7, tm_min=0.0, rep_temps_shelfs=[10.0])
    checked = 0
    with open(os.path.join(HERE, "golden", "S1.json")) as f:
        cases = json.load(f)
    for case in cases:
        inp = sio.make_input(case["name"], case["sec_struct"], case["seq_restr"])
        inp.add_alt_sec_struct(case["alt_sec_structs"])
        if case.get("rejected"):
            try:
                su.prepare_alternative_structures(inp)
            except ValueError:
                continue
            raise AssertionError("accepted an input the reference rejects: %s" % case["name"])
        su.prepare_alternative_structures(inp)
        assert inp.graphs == case["graphs"], case["name"]
        assert sorted(map(list, inp.excluded_alt_pairs)) == case["excluded_alt_pairs"], case["name"]
        nts = su.get_nt_list(inp)
        assert sorted(map(list, inp.pairs)) == case["pairs"], case["name"]
        for nt, want in zip(nts, case["nts"]):
            assert sorted(nt.letters_allowed) == want["allowed"] and nt.pairs_with == want["pairs_with"], (case["name"], nt.number)
            assert bool(nt.snake) == want["snake"] and nt.snake_number == want["snake_number"], (case["name"], nt.number)
        for row in case["init"]:
            random.seed(row["seed"])
            assert su.initial_sequence_generator(nts, inp, opt) == row["sequence"], (case["name"], row["seed"])
        for row in case["moves"]:
            random.seed(row["seed"])
            cur = SimpleNamespace(sequence=su.initial_sequence_generator(nts, inp, opt), mfe_ss=None, temp_shelf=10.0)
            assert cur.sequence == row["sequences"][0]
            for k, want in enumerate(row["sequences"][1:]):
                cur = SimpleNamespace(sequence=su.propose_mutation(cur, nts, opt, inp), mfe_ss=None, temp_shelf=10.0)
                assert cur.sequence == want, (case["name"], row["seed"], k, cur.sequence, want)
                checked += 1
    print(checked, "moves checked")


if __name__ == "__main__":
    main()
