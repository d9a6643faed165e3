"""Debugging aid: cycle stamps of one cluster's phases (BF_CL_TRACE=1, bf_cluster.cu).  usage: cl_trace.py [L] [C]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.update({"BF_CL": "1", "BF_CL_C": sys.argv[2] if len(sys.argv) > 2 else "8", "BF_CL_TRACE": "1"})
from desirna_b200 import engine
engine.init(); engine.params_builtin(1999)
L = int(sys.argv[1]) if len(sys.argv) > 1 else 400
rng = np.random.default_rng(1)
seqs = ["".join("ACGU"[x] for x in rng.integers(0, 4, L))]
os.makedirs("gpurun_out", exist_ok=True)
for _ in range(2): engine.score_batch(seqs, want=engine.WANT_MFE)
t = np.fromfile("gpurun_out/cl_trace.bin", dtype=np.int64).reshape(2048, 16, 8)
NT = 6
for d in (20, 60, 100, 200, 300, 380):
    if d >= L: continue
    base = t[d, :, 0].min()
    print(f"d={d}: phase length (start to start) {t[d + 1, :, 0].min() - base}")
    for w in (0, 1, NT - 1, NT, NT + 1, 14, 15):
        print("   warp", w, " ".join(f"{(x - base) if x else -1:6d}" for x in t[d, w]))
ph = np.array([t[d + 1, :, 0].min() - t[d, :, 0].min() for d in range(5, L - 2)])
print("mean phase cycles", ph.mean(), "min", ph.min(), "max", ph.max())
def sec(w, a, b):
    v = [t[d, w, b] - t[d, w, a] for d in range(10, L - 2) if t[d, w, a] and t[d, w, b]]
    return np.mean(v) if v else -1
print("tap warp 0: cells %.0f stage %.0f to-arrive(2->5) %.0f arrive %.0f wait %.0f" % (sec(0, 0, 1), sec(0, 1, 2), sec(0, 2, 5), sec(0, 5, 6), sec(0, 6, 7)))
for w in (NT, NT + 1, 10, 15):
    print("aux warp %d: clear+combine %.0f flush %.0f split %.0f arrive %.0f wait %.0f" % (w, sec(w, 0, 3), sec(w, 3, 4), sec(w, 4, 5), sec(w, 5, 6), sec(w, 6, 7)))
