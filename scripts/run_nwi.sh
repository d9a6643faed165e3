mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "parity or cluster" > gpurun_out/s3w_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s3w_pytest.log
LAT_LS=100,111,125,148,170,180 LAT_BS=64 python scripts/lat3.py default 2>&1 | tee gpurun_out/s3w_nwi.log
