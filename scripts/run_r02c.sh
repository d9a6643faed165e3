mkdir -p gpurun_out
timeout 300 python scripts/eterna100.py --time 60 --replicas 10 --out gpurun_out/r02c_eterna_r10_60s.json > gpurun_out/r02c_eterna_r10_60s.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r02c_eterna_r10_60s.log | cut -c1-900
BF_CL=0 BF_EXT_WIDE=0 timeout 300 python scripts/eterna100.py --time 60 --replicas 10 --out gpurun_out/r02c_eterna_r10_60s_nocl.json > gpurun_out/r02c_eterna_r10_60s_nocl.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/r02c_eterna_r10_60s_nocl.log | cut -c1-900
timeout 300 python bench.py --workload remc --steps 3 --warmup 1 > gpurun_out/r02c_remc.json 2> gpurun_out/r02c_remc.err; tail -c 900 gpurun_out/r02c_remc.json
python - <<'P'
import bench, json
print(json.dumps(bench.bench_design_loop()))
P
