import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine
engine.init(); engine.params_builtin(1999)
rng = np.random.default_rng(1)
seqs = ["".join("ACGU"[x] for x in rng.integers(0, 4, 17)) + "&" + "".join("ACGU"[x] for x in rng.integers(0, 4, 18)) for _ in range(64)]
for _ in range(3): engine.score_batch(seqs, want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF)
