#!/bin/bash
# kernel times of the fill3 variants on the bench workload: scripts/sweep_fill3.sh L B
L=${1:-100}; B=${2:-4096}
for cfg in "8 4" "8 3" "8 5" "12 6" "12 5" "12 4" "16 8" "16 6" "6 3" "6 2"; do
  set -- $cfg
  for fms in 1 0; do
    echo -n "NW=$1 NWI=$2 FMS=$fms: "
    BF_FILL3_NW=$1 BF_FILL3_NWI=$2 BF_FILL3_FMS=$fms timeout 120 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --L $L --B $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['kernel_ms'])"
  done
done
