run() { L=$1; shift; r=$(env "$@" python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu --L $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], round(d['value']))"); echo "L=$L $* -> $r"; }
for L in 36 50 70; do run $L BF_X=0; run $L BF_MFE_NW=4 BF_PF_NW=4; run $L BF_MFE_NW=2 BF_PF_NW=2; done
