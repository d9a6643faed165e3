import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine
engine.init(); engine.params_builtin(1999)
want = engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF | engine.WANT_EVAL
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for L in [int(x) for x in os.environ.get("LAT_LS", "36,70,100,125,148").split(",")]:
    for B in [int(x) for x in os.environ.get("LAT_BS", "10,64").split(",")]:
        rng = np.random.default_rng(L * 1000 + B)
        seqs = ["".join("ACGU"[x] for x in row) for row in rng.integers(0, 4, (B, L))]
        tg = [["." * L]] * B
        try:
            for _ in range(3): engine.score_batch(seqs, tg, want=want)
            km = np.zeros(3); n = 5
            for _ in range(n):
                engine.score_batch(seqs, tg, want=want); km += np.array(engine.last_kernel_ms())
            print(f"{tag} L{L}_B{B} mfe {km[0] / n:.3f} pf {km[1] / n:.3f}", flush=True)
        except Exception as e:
            print(f"{tag} L{L}_B{B} failed: {e}", flush=True)
