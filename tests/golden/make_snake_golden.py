#!/usr/bin/env python3
"""Golden vectors of the reference's alternative-structure ("snake") move generator, produced by IMPORTING the reference.

Run in the build container only (reads /root/reference; not present on the GPU box).  The reference's utils/sequence_utils.py
imports ViennaRNA (`import RNA`) and its scoring module at module level; none of the functions exercised here call them, so a stub
module stands in for RNA and `es.score_sequence` is replaced by a recorder.  What is recorded, per input:
  * the conflict graphs (utils/sequence_utils.py:143-396: get_pairs_for_graphs, generate_graphs, update_graphs) and the pairs
    that get_nt_list moves from the alternative structures into the ordinary pair list (:454-525);
  * per position: sorted letters_allowed, partner, snake membership;
  * initial_sequence_generator (:667-763) for a few seeds;
  * mutate_sequence (:1008-1136) for a few hundred consecutive moves per seed with point mutations off (the position draw
    does not need a folded structure then), i.e. the complete draw sequence of the snake branch.
Writes tests/golden/S1.json.  Run as  PYTHONHASHSEED=0 python tests/golden/make_snake_golden.py : the reference draws the partner
letter of a pair move from an unsorted set of strings, so its draws depend on the interpreter's string-hash seed."""
import json
import os
import random
import sys
import types

assert os.environ.get("PYTHONHASHSEED") == "0", "run with PYTHONHASHSEED=0 (see the docstring)"
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
_rna = types.ModuleType("RNA")                         # never called by the functions below; energy_scores.py builds RNA.md() at import
_rna.md = lambda *a, **k: types.SimpleNamespace()
sys.modules["RNA"] = _rna
for name in ("matplotlib", "matplotlib.pyplot", "pandas"):
    if name not in sys.modules:
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)

from utils import sequence_utils as su                 # noqa: E402  (the reference's module)
from utils import energy_scores as es                  # noqa: E402


class Inp:                                             # the fields of stats_inputs_outputs.InputFile the functions read (:795-870)
    def __init__(self, name, sec_struct, seq_restr, alts, seed_seq=None):
        self.name, self.sec_struct, self.seq_restr = name, sec_struct, seq_restr
        self.pairs, self.alt_pairs, self.seed_seq = [], None, seed_seq
        self.alt_sec_struct, self.alt_sec_structs = alts[0], list(alts)
        self.target_pairs_tupl, self.graphs, self.excluded_alt_pairs, self.allsnakes = {}, None, None, None


class Opt:
    acgu_percentages = "off"
    nt_percentages = {"A": 15, "C": 30, "G": 30, "U": 15}
    point_mutations = "off"
    oligo_state = "none"
    tm_max, tm_min = 0.7, 0.0
    rep_temps_shelfs = [10.0]


class Rec:                                             # stands in for ScoreSeq: what mutate_sequence reads and returns
    def __init__(self, seq):
        self.sequence, self.mfe_ss, self.replica_num, self.temp_shelf = seq, "." * len(seq), 1, 10.0

    def get_replica_num(self, r):
        self.replica_num = r

    def get_temp_shelf(self, t):
        self.temp_shelf = t


es.score_sequence = lambda seq, input_file, sim_options: Rec(seq)

CASES = [
    ("Alt_Struct_Example", "((((((.((((((((....))))).)).).))))))", "N" * 36,
     ["(((((((((((((....)))..)).)).).))))).", "(((((((((((((....)))))...)).).)))))."]),
    ("switch", "((((((....))))))....((((....))))", "N" * 32, ["....((((((....))))))((((....))))", "((((((....))))))....((((....))))"]),
    ("restrained", "(((((...)))))......", "NNNNNGNNNNNNNNNNNNN", ["......(((((...)))))"]),
    ("shifted", "..((((((...))))))....", "N" * 21, ["...((((((...))))))...", ".((((((...)))))).....", "(((...)))............"]),
    ("toggle3", "((((....))))....((((....))))....", "N" * 32, ["....((((....))))....((((....))))", "((((....))))....((((....))))...."]),
    ("shift2", "..((((((...))))))....", "N" * 21, ["....((((((...))))))..", "..((((((...))))))...."]),
    ("contradiction", "(((((...)))))......", "NNNNNANNNNNNANNNNNN", ["......(((((...)))))"]),
    ("hairpin_vs_long", "((((((((....))))))))............", "NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN", ["............((((((((....))))))))", "((((....))))........((((....))))"]),
]

out = []
for name, ss, restr, alts in CASES:
    inp = Inp(name, ss, restr, alts)
    inp.pairs = su.check_dot_bracket(inp.sec_struct)
    inp.target_pairs_tupl = {tuple(p) for p in inp.pairs}
    alt_pairs = su.get_pairs_for_graphs(inp)
    try:
        inp.graphs = su.generate_graphs(alt_pairs)
        su.update_graphs(inp)
    except (SystemExit, UnboundLocalError):            # the reference rejects the input (odd cycle, contradicting restraints) or
                                                       # trips over an input without any conflict between the structures
        out.append({"name": name, "sec_struct": ss, "seq_restr": restr, "alt_sec_structs": alts, "rejected": True})
        continue
    nts = su.get_nt_list(inp)
    su.check_input_logic(nts)
    rec = {"name": name, "sec_struct": ss, "seq_restr": restr, "alt_sec_structs": alts,
           "graphs": inp.graphs, "excluded_alt_pairs": sorted(map(list, inp.excluded_alt_pairs)), "pairs": sorted(map(list, inp.pairs)),
           "nts": [{"allowed": sorted(nt.letters_allowed), "pairs_with": nt.pairs_with, "snake": bool(nt.snake),
                    "snake_number": nt.snake_number} for nt in nts],
           "init": [], "moves": []}
    for seed in (1, 2, 3):
        random.seed(seed)
        rec["init"].append({"seed": seed, "sequence": su.initial_sequence_generator(nts, inp, Opt)})
    for seed in (11, 12):
        random.seed(seed)
        cur = Rec(su.initial_sequence_generator(nts, inp, Opt))
        seqs = [cur.sequence]
        for _ in range(150):
            cur = su.mutate_sequence(cur, nts, Opt, inp)
            seqs.append(cur.sequence)
        rec["moves"].append({"seed": seed, "sequences": seqs})
    out.append(rec)

with open(os.path.join(HERE, "S1.json"), "w") as f:
    json.dump(out, f, indent=0)
print("wrote", len(out), "cases;", sum(len(r.get("graphs", [])) for r in out), "graphs;", sum(1 for r in out if r.get("rejected")), "rejected")
