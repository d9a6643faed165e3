timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python scripts/latency.py 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read())['latency']; print({k:(v['call_ms'],v['pf_ms']) for k,v in d.items() if 'B64' in k or 'B10' in k})"
python - <<'P'
import os, sys, json, time, random
sys.path.insert(0, os.getcwd())
from desirna_b200 import design, engine
from desirna_b200.utils import stats_inputs_outputs as sio
engine.init(0); engine.params_builtin(1999)
rows = [json.loads(l) for l in open("tests/golden/E1.jsonl")]
out = {}
for want_len, R in ((36, 10), (104, 64), (200, 64), (400, 64)):
    one = min(rows, key=lambda r: (abs(len(r["target"]) - want_len), r["file"]))
    random.seed(0)
    loop = design.DesignLoop([sio.make_input(one["file"], one["target"])], design.DesignOptions(replicas=R, RE_attempt=20), seed=1)
    loop.run(1); loop.sync()
    t0 = time.perf_counter(); loop.run(2); loop.sync(); dt = time.perf_counter() - t0
    loop.close()
    out["L%d_R%d" % (len(one["target"]), R)] = round(dt / 40 * 1e3, 3)
print("ms/sub-step", out)
P
