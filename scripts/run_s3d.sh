mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s3d_pytest_gpu.log
timeout 300 python scripts/two_time.py 2>&1 | grep "SMEM': '1'" | tee gpurun_out/s3d_two_time.log
python scripts/two_substep.py 2>&1 | tee gpurun_out/s3d_two_substep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bf_k_mfe -c 1 -s 2 -o gpurun_out/s3d_two_prof -f python scripts/two_prof.py > gpurun_out/s3d_two_prof.log 2>&1; echo rc=$?
