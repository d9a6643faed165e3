#!/usr/bin/env python3
"""Extract golden vectors from the reference's shipped artefacts.

Run in the build container only (reads /root/reference, which does not exist on the GPU
box).  Writes small JSONL fixtures next to this file; they are committed.

Sources (all Turner 1999, produced by ViennaRNA through an older DesiRNA in Aug 2023,
whose trajectory column `mfe` is today's `Epf`):
  example_files/outputs/*/trajectory_files/*_traj.csv      one CSV row per replica state
  example_files/outputs/*/trajectory_files/*_random.csv    one python-dict literal per line
  eterna_benchmark/Eterna100V{1,2}_benchmark_results/*_all_results.txt
Also copies nothing else: the parameter file is re-encoded by make_params_fixture below
as a compact JSON of the integer tables so the GPU box can load Turner 1999 without the
reference tree (the numbers are data, not code).
"""
import ast
import csv
import glob
import json
import os
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

SCEN = {
    "Standard": "G1", "Seed_sequence": "G2", "Alternative_structures": "G3",
    "Pseudoknot": "G4", "RNA_RNA_complex": "G5", "Homodimer": "G6",
}


def read_input(path):
    d, key = {}, None
    for line in open(path):
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            key = line[1:]
            d[key] = []
        else:
            d[key].append(line)
    return d


def main():
    out_rows = {}
    for odir in sorted(glob.glob(os.path.join(REF, "example_files/outputs/*"))):
        base = os.path.basename(odir)
        tag = next(v for k, v in SCEN.items() if base.startswith(k))
        inp = read_input(glob.glob(os.path.join(odir, "*_input.txt"))[0])
        target = inp["sec_struct"][0]
        alts = inp.get("alt_sec_struct", [])
        seen = {}
        rel = os.path.relpath(odir, REF)
        for f in glob.glob(os.path.join(odir, "trajectory_files/*_traj.csv")):
            for row in csv.DictReader(open(f)):
                seen.setdefault(row["sequence"], row)
        for f in glob.glob(os.path.join(odir, "trajectory_files/*_random.csv")):
            for line in open(f):
                line = line.strip()
                if line.startswith("{"):
                    row = ast.literal_eval(line)
                    seen.setdefault(row["sequence"], row)
        rows = []
        for seq, r in sorted(seen.items()):
            rows.append({
                "set": tag, "sequence": seq, "target": target, "alts": alts,
                "Epf": float(r["mfe"]), "Ed": float(r["edesired"]), "mfe_ss": r["mfe_ss"],
                "Ed_minus_Epf": float(r["edesired_minus_mfe"]),
                "one_minus_mcc": float(r["mcc"]), "one_minus_recall": float(r["recall"]),
                "one_minus_precision": float(r["precision"]),
                "Ed2": float(r["edesired2"]), "scoring_function": float(r["scoring_function"]),
                "src": rel,
            })
        out_rows[tag] = rows
    for tag, rows in out_rows.items():
        with open(os.path.join(HERE, f"{tag}.jsonl"), "w") as fo:
            for r in rows:
                fo.write(json.dumps(r) + "\n")
        print(tag, len(rows))
    for ver in ("V1", "V2"):
        rows = []
        p = os.path.join(REF, f"eterna_benchmark/Eterna100{ver}_benchmark_results/Eterna100{ver}_all_results.txt")
        # which time budget of the reference's benchmark solved the puzzle (*_1min_, *_1h_, *_24h_results.txt)
        tier = {}
        for name in ("1min", "1h", "24h"):
            q = os.path.join(REF, f"eterna_benchmark/Eterna100{ver}_benchmark_results/Eterna100{ver}_{name}_results.txt")
            for line in open(q):
                if line.strip():
                    tier.setdefault(line.strip().split(",")[0][1:], name)
        for line in open(p):
            f = line.strip().split(",")
            if len(f) == 3:
                rows.append({"set": "E" + ver[1], "file": f[0][1:], "sequence": f[1], "target": f[2], "ref_solved_within": tier.get(f[0][1:])})
            elif len(f) >= 6:
                rows.append({"set": "E" + ver[1], "file": f[0][1:], "sequence": f[3], "target": f[4], "ref_solved_within": tier.get(f[0][1:])})
        with open(os.path.join(HERE, f"E{ver[1]}.jsonl"), "w") as fo:
            for r in rows:
                fo.write(json.dumps(r) + "\n")
        print("E" + ver[1], len(rows))


if __name__ == "__main__":
    main()
