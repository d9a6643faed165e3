mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_design.py -x -q > gpurun_out/design_pytest.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/design_pytest.log
