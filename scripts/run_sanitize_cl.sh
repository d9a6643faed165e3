mkdir -p gpurun_out
# memcheck of the cluster-per-sequence kernels, the wide exterior kernels and the fill3 kernels at the stride that used to overflow (60)
cat > /tmp/sancl.py <<'P'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from desirna_b200 import engine
engine.init(0); engine.params_builtin(1999)
rng = np.random.default_rng(3)
want = engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF
for L, C in ((70, 4), (150, 8), (210, 16)):
    seqs = ["".join("ACGU"[x] for x in rng.integers(0, 4, int(n))) for n in (L, L - 7, L // 2)]
    os.environ.update({"BF_CL": "1", "BF_CL_C": str(C)})
    a = engine.score_batch(seqs, want=want)
    os.environ["BF_CL"] = "0"
    b = engine.score_batch(seqs, want=want)
    print(L, C, a["mfe_dcal"].tolist(), bool((a["mfe_dcal"] == b["mfe_dcal"]).all()), float(np.abs(a["pf"][:, 4] - b["pf"][:, 4]).max()))
for k in ("BF_CL", "BF_CL_C"): os.environ.pop(k, None)
seqs = ["".join("ACGU"[x] for x in row) for row in rng.integers(0, 4, (200, 60))]
out = engine.score_batch(seqs, want=want); print("fill3 stride 60:", int(out["mfe_dcal"].sum()))
P
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/sancl.py > gpurun_out/r02_sanitize_memcheck_cluster.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r02_sanitize_memcheck_cluster.log
