mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python scripts/eterna100.py --time 60 --replicas 10 --out gpurun_out/eterna_r10_60s.json > gpurun_out/eterna_r10_60s.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/eterna_r10_60s.log | cut -c1-1400
timeout 300 python scripts/eterna100.py --time 60 --replicas 20 --out gpurun_out/eterna_r20_60s.json > gpurun_out/eterna_r20_60s.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/eterna_r20_60s.log | cut -c1-1400
