mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3n_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s3n_pytest_gpu.log
(timeout 300 python scripts/two_time.py 2>&1 | grep "SMEM': '1'"; python scripts/two_substep.py) 2>&1 | tee gpurun_out/s3n_two.log
