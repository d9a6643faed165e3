# placement sweep at mid lengths: kernel ms (mfe, pf) per 4096 folds
mkdir -p gpurun_out
: > gpurun_out/sweep_pl.log
run() { # L env...
  L=$1; shift
  r=$(env "$@" python bench.py --steps 2 --warmup 2 --no-sweep --no-cpu --L $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'])")
  echo "L=$L $* -> $r" | tee -a gpurun_out/sweep_pl.log
}
for L in 150 200 300; do
  run $L BF_X=0
  run $L BF_PF_PL=4
  run $L BF_MFE_PL=2
  run $L BF_PF_NW=4 BF_MFE_NW=4
done
run 300 BF_BLK_MIN_PF=400
run 300 BF_BLK_MIN=250
run 200 BF_BLK_MIN_PF=180
