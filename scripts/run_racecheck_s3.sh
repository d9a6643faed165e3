mkdir -p gpurun_out
cat > /tmp/race3.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from desirna_b200 import engine
engine.init(0); engine.params_builtin(1999)
rng = np.random.default_rng(3)
def rs(n): return "".join("ACGU"[x] for x in rng.integers(0, 4, n))
for a, b in ((3, 4), (17, 18), (30, 25), (50, 50)):
    seqs = [rs(a) + "&" + rs(b) for _ in range(2)]
    print("two-strand", a, b, engine.score_batch(seqs, want=7)["mfe_dcal"].tolist(), flush=True)
for L in (36, 100):
    seqs = [rs(L) for _ in range(2)]
    print("small batch", L, engine.score_batch(seqs, want=7)["mfe_dcal"].tolist(), flush=True)
P
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/race3.py > gpurun_out/r02_sanitize_racecheck_s3.log 2>&1; echo "racecheck rc=$?"; tail -12 gpurun_out/r02_sanitize_racecheck_s3.log
