import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine
engine.init(); engine.params_builtin(1999)
seq = sys.argv[1] if len(sys.argv) > 1 else "GGGGAAACCCC"
n = len(seq)
def tri_off(n, d): return (d - 4) * n - (d * (d - 1) // 2 - 6)
tabs = {}
for name, env in (("ref", {"BF_CL": "0"}), ("cl", {"BF_CL": "1", "BF_CL_C": "2"})):
    os.environ.update(env)
    r = engine.score_batch([seq], want=engine.WANT_MFE | engine.WANT_PF)
    tabs[name] = np.array(engine.debug_table(2, 1)).reshape(-1)
    print(name, r["pf"][0, 4])
    for k in env: os.environ.pop(k)
for d in range(4, n):
    a = tabs["ref"][tri_off(n, d):tri_off(n, d) + n - d]; b = tabs["cl"][tri_off(n, d):tri_off(n, d) + n - d]
    print("d", d, "ref", " ".join(f"{x:.3e}" for x in a)); print("     cl ", " ".join(f"{x:.3e}" for x in b))
