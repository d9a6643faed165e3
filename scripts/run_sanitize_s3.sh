mkdir -p gpurun_out
# memcheck of what the last session of round 2 added: two-barrier generic kernels (PRE) and their candidate lists at several strides
# (shared-memory and HBM tables), 16-warp third-generation kernels on small batches, the staged bf_score_batch path with the
# partition function beside the MFE fill, the two-strand design loop with the fills side by side, the backtrack scans
cat > /tmp/san3.py <<'P'
import os, sys, random
sys.path.insert(0, os.getcwd())
import numpy as np
from desirna_b200 import design, engine
from desirna_b200.utils import stats_inputs_outputs as sio
engine.init(0); engine.params_builtin(1999)
rng = np.random.default_rng(3)
def rs(n): return "".join("ACGU"[x] for x in rng.integers(0, 4, n))
for a, b in ((1, 1), (3, 4), (17, 18), (18, 18), (30, 25), (50, 50), (70, 90)):
    seqs = [rs(a) + "&" + rs(b) for _ in range(3)]
    out = engine.score_batch(seqs, [["." * (a + b)]] * 3, want=15)
    print("two-strand", a, b, out["mfe_dcal"].tolist(), flush=True)
seqs = [rs(20) + "&" + rs(20) for _ in range(200)]   # large batch: 8-warp four-barrier variants
print("two-strand x200", int(engine.score_batch(seqs, want=7)["mfe_dcal"].sum()), flush=True)
for L in (5, 36, 100, 148, 200, 228):
    seqs = [rs(L) for _ in range(5)] + [rs(max(1, L // 3))]
    out = engine.score_batch(seqs, [["." * L]] * 6, want=15)
    print("small batch", L, out["mfe_dcal"].tolist(), flush=True)
seqs = [rs(70) for _ in range(6)]
t = engine.score_batch(seqs, want=engine.WANT_MFE | engine.WANT_SS)["mfe_ss"]
out = engine.score_batch(seqs, [[x] for x in t], want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF | engine.WANT_DEFECT | engine.WANT_EVAL)
print("defect", out["defect"].tolist(), flush=True)
random.seed(0)
het = sio.make_input("het", "(((.(((((....))..&(((....)))..))))))", "NNNNNNNNNNNNNNNNN&NNNNNNNNNNNNNNNNNN")
loop = design.DesignLoop([het], design.DesignOptions(replicas=8, RE_attempt=6, oligo_state="heterodimer"), seed=4)
loop.run(2); print("het", loop.jobs()["mfe_ss"], flush=True); loop.close()
loop = design.DesignLoop([sio.make_input("s", "((((((.((((((((....))))).)).).))))))")], design.DesignOptions(replicas=10, RE_attempt=6), seed=5)
loop.run(2); print("std", loop.jobs()["mfe_ss"], flush=True); loop.close()
P
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san3.py > gpurun_out/r02_sanitize_memcheck_s3.log 2>&1; echo "memcheck rc=$?"; tail -25 gpurun_out/r02_sanitize_memcheck_s3.log
