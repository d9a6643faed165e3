mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "remc or lockstep or design" > gpurun_out/s3q_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s3q_pytest_gpu.log
timeout 300 python bench.py --workload remc --steps 3 --warmup 1 > gpurun_out/r02g_remc_1gpu.json 2> gpurun_out/r02g_remc_1gpu.err; echo "remc rc=$?"
python -c "
import json
for ln in open('gpurun_out/r02g_remc_1gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['value'], d['by_target'])
"
