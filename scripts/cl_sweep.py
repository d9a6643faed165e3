"""Kernel times of the default fill kernels against the cluster-per-sequence kernels (bf_cluster.cu) for small batches."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine
engine.init(); engine.params_builtin(1999)
want = engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF
def t(seqs, env):
    for k, v in env.items(): os.environ[k] = str(v)
    try:
        for _ in range(2): engine.score_batch(seqs, want=want)
        km = np.zeros(3)
        for _ in range(3):
            engine.score_batch(seqs, want=want); km += np.array(engine.last_kernel_ms())
    finally:
        for k in env: os.environ.pop(k, None)
    return km / 3
for L in (100, 150, 200, 250, 300, 400):
    for B in (1, 4, 8, 16, 32, 64):
        rng = np.random.default_rng(L * 7 + B)
        seqs = ["".join("ACGU"[x] for x in rng.integers(0, 4, L)) for _ in range(B)]
        row = [f"L={L} B={B}:"]
        base = t(seqs, {"BF_CL": 0}); row.append(f"default mfe {base[0]:.2f} pf {base[1]:.2f} |")
        for C in (2, 4, 8, 16):
            try:
                k = t(seqs, {"BF_CL": 1, "BF_CL_C": C}); row.append(f"C{C} {k[0]:.2f}/{k[1]:.2f}")
            except Exception as e:
                row.append(f"C{C} n/a")
        print(" ".join(row), flush=True)
