"""sub-step time of single-strand design loops: python scripts/design_substep.py"""
import os, sys, time, random, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import design
from desirna_b200.utils import stats_inputs_outputs as sio
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = [json.loads(l) for l in open(os.path.join(ROOT, "tests", "golden", "E1.jsonl"))]
one = min(rows, key=lambda r: (abs(len(r["target"]) - 104), r["file"]))
for name, inp, R in (("36nt", sio.make_input("Standard_design", "((((((.((((((((....))))).)).).))))))"), 10), ("104nt", sio.make_input(one["file"], one["target"]), 64)):
    random.seed(0)
    loop = design.DesignLoop([inp], design.DesignOptions(replicas=R, RE_attempt=100), seed=1)
    loop.run(1); loop.sync()
    t0 = time.perf_counter()
    loop.run(2); loop.sync()
    dt = time.perf_counter() - t0
    loop.close()
    print(f"{name} R={R} env={ {k: v for k, v in os.environ.items() if k.startswith('BF_')} }: {dt / 200 * 1e3:.4f} ms per sub-step", flush=True)
