mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/wide_pytest.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/wide_pytest.log
BF_WIDE=0 python scripts/latency.py > gpurun_out/latency_narrow.json 2>gpurun_out/latency_narrow.err; cat gpurun_out/latency_narrow.json
BF_WIDE=1 python scripts/latency.py > gpurun_out/latency_wide.json 2>gpurun_out/latency_wide.err; cat gpurun_out/latency_wide.json
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/wide_bench.json 2>gpurun_out/wide_bench.err; python -c "
import json; d=json.load(open('gpurun_out/wide_bench.json')); print(d['by_length'], d['kernel_ms_by_length'])"
