#!/bin/bash
L=${1:-100}; B=${2:-4096}
for dbg in 0 16 1 2 8 11; do
    echo -n "DBG=$dbg: "
    BF_FILL3_DBG=$dbg timeout 120 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --L $L --B $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['kernel_ms']['bf_k_mfe'])"
done
for cfg in "8 4" "8 5" "8 6"; do
  set -- $cfg
  for fms in 1 0; do
    echo -n "NW=$1 NWI=$2 FMS=$fms: "
    BF_FILL3_NW=$1 BF_FILL3_NWI=$2 BF_FILL3_FMS=$fms timeout 120 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --L $L --B $B 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['kernel_ms']['bf_k_mfe'])"
  done
done
