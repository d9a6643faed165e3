"""a few sub-steps of the device design loop, for an ncu launch list: python scripts/design_launches.py <36|104>"""
import os, sys, json, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import design
from desirna_b200.utils import stats_inputs_outputs as sio
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
which = sys.argv[1] if len(sys.argv) > 1 else "36"
random.seed(0)
if which == "36":
    inp, R = sio.make_input("Standard_design", "((((((.((((((((....))))).)).).))))))"), 10
else:
    rows = [json.loads(l) for l in open(os.path.join(ROOT, "tests", "golden", "E1.jsonl"))]
    one = min(rows, key=lambda r: (abs(len(r["target"]) - 104), r["file"]))
    inp, R = sio.make_input(one["file"], one["target"]), 64
loop = design.DesignLoop([inp], design.DesignOptions(replicas=R, RE_attempt=4), seed=1)
loop.run(2); loop.sync(); loop.close()
