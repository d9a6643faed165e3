import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine
engine.init(); engine.params_builtin(1999)
want = engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF | engine.WANT_EVAL
for L in (36, 100, 140):
    for B in (10, 64, 128):
        rng = np.random.default_rng(L * 1000 + B)
        seqs = ["".join("ACGU"[x] for x in row) for row in rng.integers(0, 4, (B, L))]
        tg = [["." * L]] * B
        for _ in range(3): engine.score_batch(seqs, tg, want=want)
        km = np.zeros(3); n = 5; t0 = time.perf_counter()
        for _ in range(n):
            engine.score_batch(seqs, tg, want=want); km += np.array(engine.last_kernel_ms())
        print(f"L{L}_B{B}", "call_ms %.3f" % ((time.perf_counter() - t0) / n * 1e3), "mfe %.3f pf %.3f" % (km[0] / n, km[1] / n))
