mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s3k_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/s3k_pytest_gpu.log
(python scripts/lat2.py; BF_STAGE=0 BF_SCORE_OVERLAP=0 python scripts/lat2.py | sed 's/^/before: /'; python scripts/two_time.py | grep "SMEM': '1'") 2>&1 | tee gpurun_out/s3k_call.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
