mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s3b_pytest_gpu.log
(timeout 300 python scripts/two_time.py 2>&1 | grep "SMEM': '1'"; BF_GEN_PRE=0 timeout 300 python scripts/two_time.py 2>&1 | grep "SMEM': '1'" | sed 's/^/PRE=0 /') | tee gpurun_out/s3b_two_time.log
(python scripts/two_substep.py; BF_DESIGN_OVERLAP=0 python scripts/two_substep.py; BF_GEN_PRE=0 BF_DESIGN_OVERLAP=0 python scripts/two_substep.py) 2>&1 | tee gpurun_out/s3b_two_substep.log
