mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/quick_bench.json 2>gpurun_out/quick_bench.err; python -c "
import json; d=json.load(open('gpurun_out/quick_bench.json')); print(d['by_length']); print(d['kernel_ms_by_length']); print(d['design_loop']['one_target_64_replicas'], d['design_loop']['eterna100_x_10_replicas']['ms_per_substep'])"
run() { L=$1; shift; r=$(env "$@" python bench.py --steps 2 --warmup 2 --no-sweep --no-cpu --L $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms']['bf_k_mfe'])"); echo "L=$L $* -> mfe $r" | tee -a gpurun_out/sweep_atom.log; }
: > gpurun_out/sweep_atom.log
for L in 130 150 170 250 300; do run $L BF_X=0; done
run 150 BF_MFE_PL=0
run 170 BF_MFE_PL=2
run 250 BF_MFE_ATOM=1
run 200 BF_MFE_ATOM=1
python scripts/latency.py > gpurun_out/quick_latency.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/quick_latency.json'))['latency']; print({k:(v['call_ms'],v['mfe_ms']) for k,v in d.items() if 'B64' in k})"
