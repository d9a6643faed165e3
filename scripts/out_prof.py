"""inputs for an ncu capture of the outside kernel: python scripts/out_prof.py [L] [B]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine
L = int(sys.argv[1]) if len(sys.argv) > 1 else 100; B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
engine.init(); engine.params_builtin(1999)
rng = np.random.default_rng(20240000 + L)
seqs = ["".join("ACGU"[x] for x in row) for row in rng.integers(0, 4, (B, L))]
ss = engine.score_batch(seqs, want=engine.WANT_MFE | engine.WANT_SS)["mfe_ss"]
for _ in range(2):
    out = engine.score_batch(seqs, [[s] for s in ss], want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF | engine.WANT_DEFECT | engine.WANT_EVAL)
print(float(out["defect"].mean()))
