mkdir -p gpurun_out
# memcheck + racecheck of the design-loop kernels and the 16-warp fill variants on small cases
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
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitize_racecheck.log
# launch list of design sub-steps
cat > /tmp/dl.py <<'P'
import os, sys, json, random
sys.path.insert(0, os.getcwd())
from desirna_b200 import design, engine
from desirna_b200.utils import stats_inputs_outputs as sio
engine.init(0); engine.params_builtin(1999)
rows = [json.loads(l) for l in open("tests/golden/E1.jsonl")]
one = min(rows, key=lambda r: (abs(len(r["target"]) - 104), r["file"]))
o = design.DesignOptions(replicas=64, RE_attempt=4)
random.seed(0)
loop = design.DesignLoop([sio.make_input(one["file"], one["target"])], o, seed=1)
loop.run(3); loop.sync(); loop.close()
P
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/design_launches.csv python /tmp/dl.py > gpurun_out/design_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/design_launches.csv')) if len(r)>5]
h=rows[0]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
from collections import defaultdict
t=defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    n=r[kn].split('(')[0].split('::')[-1][:40]; t[n][0]+=1; t[n][1]+=float(r[mv].replace(',',''))
for n,(c,s) in sorted(t.items(), key=lambda x:-x[1][1]): print(f"{n:42s} launches {c:4d}  total {s/1e3:9.1f} us  avg {s/c/1e3:8.1f} us")
P
