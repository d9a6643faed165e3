mkdir -p gpurun_out
timeout 200 python scripts/eterna100.py --time 60 --replicas 10 --out gpurun_out/r02f_eterna100_r10_60s.json > gpurun_out/r02f_eterna_r10_60s.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r02f_eterna_r10_60s.log | cut -c1-1200
