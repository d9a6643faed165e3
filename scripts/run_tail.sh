mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3y_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s3y_pytest_gpu.log
(python scripts/design_substep.py; python scripts/two_substep.py; LAT_LS=36,100,148 LAT_BS=64 python scripts/lat3.py default) 2>&1 | tee gpurun_out/s3y_small.log
python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['kernel_ms_by_length'])"
