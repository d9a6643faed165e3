timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
run() { L=$1; shift; r=$(env "$@" python bench.py --steps 2 --warmup 2 --no-sweep --no-cpu --L $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['roofline']['kernel_ms']['bf_k_pf'], d['checks']['ed_equals_mfe_and_epf_le_mfe'])"); echo "L=$L $* -> pf $r"; }
for L in 250 300 350 400; do for o in 1 0; do run $L BF_PF_BLK3=$o; done; done
