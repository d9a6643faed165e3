bash scripts/gpu_round.sh s4b 400
python - <<'P'
import json
d=json.load(open('gpurun_out/s4b_bench.json'))
print(json.dumps(d['design_loop']))
print(d['by_length'])
P
