mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { L=$1; shift; r=$(env "$@" python bench.py --steps 2 --warmup 2 --no-sweep --no-cpu --L $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms'], d['value'])"); echo "L=$L $* -> $r"; }
for L in 110 120 135 150 185; do run $L BF_X=0; done
