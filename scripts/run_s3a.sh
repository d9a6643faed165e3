mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s3a_pytest_gpu.log
timeout 300 python scripts/two_time.py 2>&1 | grep "SMEM': '1'" | tee gpurun_out/s3a_two_time.log
