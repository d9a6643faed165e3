"""timings of the cluster kernels: python scripts/cl_time.py <mfe|both> L B [env k=v ...]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine
engine.init(); engine.params_builtin(1999)
what, L, B = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
for kv in sys.argv[4:]:
    k, v = kv.split("="); os.environ[k] = v
want = engine.WANT_MFE | engine.WANT_SS | (engine.WANT_PF if what != "mfe" else 0)
rng = np.random.default_rng(L + B)
seqs = ["".join("ACGU"[x] for x in rng.integers(0, 4, L)) for _ in range(B)]
for _ in range(2): engine.score_batch(seqs, want=want)
km = np.zeros(3)
for _ in range(3):
    engine.score_batch(seqs, want=want); km += np.array(engine.last_kernel_ms())
print(f"L={L} B={B} {sys.argv[4:]}: mfe {km[0] / 3:.3f} pf {km[1] / 3:.3f}", flush=True)
