mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/quick_bench.json 2>gpurun_out/quick_bench.err; python -c "
import json; d=json.load(open('gpurun_out/quick_bench.json')); print(d['by_length']); print(d['kernel_ms_by_length']); print(json.dumps(d['design_loop']))"
python scripts/latency.py > gpurun_out/quick_latency.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/quick_latency.json'))['latency']; print({k:v['call_ms'] for k,v in d.items()})"
