mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3j_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s3j_pytest_gpu.log
(python scripts/design_substep.py; python scripts/two_substep.py; python scripts/lat3.py default; LAT_LS=170,200,230,260 LAT_BS=16,64 python scripts/lat3.py default) 2>&1 | tee gpurun_out/s3j_small.log
