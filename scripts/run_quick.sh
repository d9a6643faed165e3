mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run() { L=$1; shift; r=$(env "$@" python bench.py --steps 2 --warmup 2 --no-sweep --no-cpu --L $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms']['bf_k_pf'])"); echo "L=$L $* -> pf $r" | tee -a gpurun_out/sweep_pf.log; }
: > gpurun_out/sweep_pf.log
run 50 BF_X=0; run 50 BF_PF_HALF4=0
run 100 BF_X=0; run 100 BF_PF_HALF4=0
run 120 BF_X=0; run 120 BF_PF_HALF4=0
for L in 150 170 200 240; do run $L BF_X=0; done
