bash scripts/gpu_round.sh s4c 400
python - <<'P'
import json
d=json.load(open('gpurun_out/s4c_bench.json'))
print(json.dumps(d['design_loop']))
print(d['by_length'], d['kernel_ms_by_length'], d['with_ensemble_defect'])
P
python scripts/latency.py > gpurun_out/s4c_latency.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/s4c_latency.json'))['latency']; print({k:v['call_ms'] for k,v in d.items()})"
