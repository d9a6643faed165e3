import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine
engine.init(); engine.params_builtin(1999)
rng = np.random.default_rng(1)
seqs = ["".join("ACGU"[x] for x in rng.integers(0, 4, 50)) + "&" + "".join("ACGU"[x] for x in rng.integers(0, 4, 50)) for _ in range(4096)]
for _ in range(2): engine.score_batch(seqs, want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF)
