run() { L=$1; shift; r=$(env "$@" python bench.py --steps 2 --warmup 2 --no-sweep --no-cpu --L $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms']['bf_k_mfe'], d['checks'])"); echo "L=$L $* -> mfe $r"; }
for L in 300 350 400; do for o in 0 1; do run $L BF_MFE_MINB5=$o; done; done
