bash scripts/gpu_round.sh s4c 400
python - <<'P'
import json
d=json.load(open('gpurun_out/s4c_bench.json'))
print(json.dumps(d['design_loop']))
print(d['by_length'], d['with_ensemble_defect'])
P
python scripts/latency.py > gpurun_out/s4c_latency.json 2>/dev/null
cat > /tmp/san.py <<'P'
import os, sys, random
sys.path.insert(0, os.getcwd())
from desirna_b200 import design, engine
from desirna_b200.utils import stats_inputs_outputs as sio
engine.init(0); engine.params_builtin(1999)
inputs = [sio.make_input("a", "(((((......)))))"), sio.make_input("b", "((((...))))..((((....))))....."), sio.make_input("c", "." * 9 + "((((((....))))))" + "." * 7)]
o = design.DesignOptions(replicas=4, RE_attempt=5, scoring_f=[("Ed-Epf", 0.5), ("1-MCC", 0.5)])
random.seed(0)
loop = design.DesignLoop(inputs, o, seed=1)
loop.run(2); print(loop.jobs()["solved_step"]); loop.propose_only(); loop.set_active([1, 0, 1]); loop.run(1); print(loop.replicas()["shelf"].tolist()); loop.close()
import numpy as np
rng = np.random.default_rng(1)
for L in (60, 130):
    seqs = ["".join("ACGU"[x] for x in row) for row in rng.integers(0, 4, (5, L))]
    out = engine.score_batch(seqs, [["." * L]] * 5, want=engine.WANT_MFE | engine.WANT_SS | engine.WANT_PF | engine.WANT_EVAL)
    print(L, out["mfe_dcal"].tolist())
P
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -1 gpurun_out/sanitize_memcheck.log
