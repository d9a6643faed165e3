mkdir -p gpurun_out
timeout 900 python scripts/eterna100.py --time 600 --replicas 24 --poll 2 --out gpurun_out/eterna_r24_600s.json > gpurun_out/eterna_r24_600s.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/eterna_r24_600s.log | cut -c1-1600
