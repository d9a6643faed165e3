#!/usr/bin/env python
"""Eterna100 (V1 targets, Turner 1999) designed CONCURRENTLY on one B200 by the device-resident loop
(desirna_b200.design.design_batch): how many of the 100 puzzles are solved within a wall-clock budget.

The reference publishes, for the same 100 targets and parameters, 90 solved within 1 minute EACH (one DesiRNA
process of 10 worker processes per puzzle), 95 within 1 h, 100 within 24 h
(eterna_benchmark/Eterna100V1_benchmark_results; hardware not stated).  The targets and the reference's per-puzzle
tier come from tests/golden/E1.jsonl (extracted by tests/golden/make_golden.py).

Every sequence reported as solved is folded again through the host-buffer C-ABI entry point (bf_score_batch) and must
fold into its target; prints ONE JSON line."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", type=float, default=60.0, help="wall-clock budget in seconds for ALL puzzles together")
    ap.add_argument("--steps", type=int, default=None, help="stop after this many global steps instead")
    ap.add_argument("--replicas", type=int, default=10)
    ap.add_argument("--exchange", type=int, default=100, help="Monte-Carlo sub-steps per global step (-e)")
    ap.add_argument("--sf", default="Ed-Epf:1.0", help="scoring function terms, e.g. Ed-Epf:0.5,1-MCC:0.5")
    ap.add_argument("--max-len", type=int, default=400)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--poll", type=int, default=1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()

    from desirna_b200 import design, engine
    from desirna_b200.utils import stats_inputs_outputs as sio
    rows = [json.loads(l) for l in open(os.path.join(ROOT, "tests", "golden", "E1.jsonl"))]
    rows = [r for r in rows if len(r["target"]) <= a.max_len]
    inputs = [sio.make_input(r["file"], r["target"]) for r in rows]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:   # one process per GPU (torchrun): targets sharded over the ranks, results gathered at the end
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    engine.init()
    engine.params_builtin(1999)
    o = design.DesignOptions(replicas=a.replicas, RE_attempt=a.exchange, scoring_f=sio.parse_scoring_functions_all(a.sf))
    l0 = engine.kernel_launches()
    t0 = time.time()
    results, info = design.design_batch_sharded(inputs, o, time_limit=None if a.steps else a.time, global_steps=a.steps, seed=a.seed,
                                                poll_steps=a.poll, verbose=a.verbose and rank == 0)
    wall = time.time() - t0
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return 0
    solved = [(r, res) for r, res in zip(rows, results) if res["solved"]]
    verified = 0
    if solved:
        by_len = {}
        for r, res in solved:
            by_len.setdefault(len(r["target"]), []).append((r, res))
        for n, grp in by_len.items():
            out = engine.score_batch([res["sequence"] for _, res in grp], want=engine.WANT_MFE | engine.WANT_SS)
            verified += sum(1 for (r, _), ss in zip(grp, out["mfe_ss"]) if ss == r["target"])
    tiers = {}
    for r, res in zip(rows, results):
        t = tiers.setdefault(r["ref_solved_within"], {"puzzles": 0, "solved_here": 0})
        t["puzzles"] += 1
        t["solved_here"] += int(res["solved"])
    times = sorted(res["solved_after_s"] for _, res in solved)
    line = {
        "benchmark": "Eterna100 V1 targets, Turner 1999, all puzzles designed concurrently, targets sharded over %d GPU(s)" % world,
        "n_gpus": world, "per_rank": info.get("per_rank"),
        "puzzles": len(rows), "solved": len(solved), "solved_and_refolded_ok": verified, "wall_s": round(wall, 2),
        "budget_s": None if a.steps else a.time, "global_steps": info["global_steps"], "replicas": a.replicas, "re_attempt": a.exchange,
        "scoring_function": a.sf, "sequences_scored": info["folds"], "scored_per_s": round(info["folds"] / wall, 1),
        "kernel_launches": engine.kernel_launches() - l0, "buckets_stride_jobs": info["buckets"],
        "solved_after_s": {str(k): sum(1 for t in times if t <= k) for k in (1, 2, 5, 10, 20, 30, 60, 120, 300, 600) if k <= wall + 1},
        "by_reference_tier": tiers,
        "reference": "90/100 within 1 min per puzzle, 95 within 1 h, 100 within 24 h (one 10-process CPU run per puzzle; README.md:34)",
        "unsolved": [r["file"] for r, res in zip(rows, results) if not res["solved"]],
    }
    s = json.dumps(line)
    print(s)
    if a.out:
        with open(a.out, "w") as f:
            f.write(s + "\n")
    return 0 if verified == len(solved) else 1


if __name__ == "__main__":
    sys.exit(main())
