mkdir -p gpurun_out
timeout 300 python scripts/eterna100.py --time 60 --verbose --out gpurun_out/eterna_r10_60s.json > gpurun_out/eterna_r10_60s.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/eterna_r10_60s.log | cut -c1-1500
timeout 300 python scripts/eterna100.py --time 60 --replicas 14 --sf Ed-Epf:0.5,1-MCC:0.5 --out gpurun_out/eterna_r14_mix_60s.json > gpurun_out/eterna_r14_mix_60s.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/eterna_r14_mix_60s.log | cut -c1-1500
