"""step time with the outside pass (defect of the MFE structure): python scripts/out_time.py <tag>"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine
engine.init(); engine.params_builtin(1999)
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for L, B in [(int(x), 4096) for x in os.environ.get("OUT_LS", "50,75,100").split(",")]:
    rng = np.random.default_rng(20240000 + L)
    seqs = ["".join("ACGU"[x] for x in row) for row in rng.integers(0, 4, (B, L))]
    ss = engine.score_batch(seqs, want=engine.WANT_MFE | engine.WANT_SS)["mfe_ss"]
    tg = [[s] for s in ss]
    w0 = engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF | engine.WANT_EVAL
    for _ in range(2): out = engine.score_batch(seqs, tg, want=w0 | engine.WANT_DEFECT)
    t0 = time.perf_counter()
    for _ in range(3): out = engine.score_batch(seqs, tg, want=w0 | engine.WANT_DEFECT)
    t1 = (time.perf_counter() - t0) / 3
    for _ in range(2): engine.score_batch(seqs, tg, want=w0)
    t0 = time.perf_counter()
    for _ in range(3): engine.score_batch(seqs, tg, want=w0)
    t2 = (time.perf_counter() - t0) / 3
    print(f"{tag} L={L} B={B}: with defect {t1 * 1e3:.3f} ms, without {t2 * 1e3:.3f} ms, outside pass ~{(t1 - t2) * 1e3:.3f} ms, defect mean {float(out['defect'].mean()):.12f}", flush=True)
