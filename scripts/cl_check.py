"""Cluster-per-sequence fill kernels (bf_cluster.cu) against the default kernels on the same batches, and their timings.
usage: python scripts/cl_check.py [mfe|pf|both]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desirna_b200 import engine
engine.init(); engine.params_builtin(1999)
what = sys.argv[1] if len(sys.argv) > 1 else "mfe"
want = engine.WANT_MFE | engine.WANT_SS | (engine.WANT_PF if what != "mfe" else 0)

def batch(seed, B, L, ragged):
    rng = np.random.default_rng(seed)
    out = []
    for k in range(B):
        n = int(rng.integers(max(5, L // 2), L + 1)) if ragged and k else L
        out.append("".join("ACGU"[x] for x in rng.integers(0, 4, n)))
    return out

def run(seqs, env):
    for k, v in env.items(): os.environ[k] = str(v)
    try:
        r = engine.score_batch(seqs, want=want)
        ms = engine.last_kernel_ms()
    finally:
        for k in env: os.environ.pop(k, None)
    return r, ms

bad = 0
for L in (20, 47, 64, 100, 150, 200, 256, 300, 400):
    seqs = batch(L, 24, L, True)
    ref, _ = run(seqs, {"BF_CL": 0})
    for C in (2, 4, 8, 16):
        for BW in (32, 16):
            try:
                got, _ = run(seqs, {"BF_CL": 1, "BF_CL_C": C, "BF_CL_BW": BW})
            except Exception as e:
                print(f"L={L} C={C} BW={BW}: {str(e)[:100]}"); continue
            ok = np.array_equal(ref["mfe_dcal"], got["mfe_dcal"]) and all(a == b for a, b in zip(ref["mfe_ss"], got["mfe_ss"]))
            msg = ""
            if what != "mfe":
                rel = np.max(np.abs(ref["pf"][:, 4] - got["pf"][:, 4]) / np.maximum(1e-9, np.abs(ref["pf"][:, 4])))
                ok = ok and rel < 1e-10; msg = f" pf rel {rel:.2e}"
            print(f"L={L} C={C} BW={BW}: {'ok' if ok else 'MISMATCH'}{msg}", flush=True)
            bad += 0 if ok else 1
print("mismatches:", bad)
# timings: kernel ms of the fills, default against cluster
for L, B in ((100, 64), (200, 64), (400, 64), (400, 18), (200, 1024), (300, 1024), (400, 1024)):
    seqs = batch(7 * L + B, B, L, False)
    for env in ({"BF_CL": 0}, {"BF_CL": 1, "BF_CL_C": 4}, {"BF_CL": 1, "BF_CL_C": 8}, {"BF_CL": 1, "BF_CL_C": 16}, {"BF_CL": 1, "BF_CL_C": 8, "BF_CL_BW": 16}):
        try:
            for _ in range(2): run(seqs, env)
            km = np.zeros(3); t0 = time.perf_counter()
            for _ in range(3):
                _, ms = run(seqs, env); km += np.array(ms)
            print(f"L={L} B={B} {env}: call {1e3 * (time.perf_counter() - t0) / 3:.2f} ms, mfe {km[0] / 3:.3f} pf {km[1] / 3:.3f}", flush=True)
        except Exception as e:
            print(f"L={L} B={B} {env}: {str(e)[:100]}")
