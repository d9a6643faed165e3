#!/bin/bash
L=${1:-100}; B=${2:-4096}
for nw in 8 16; do
for dbg in 0 1 2 8 4 3 11 15; do
  echo -n "NW=$nw DBG=$dbg: "
  BF_FILL3_PF_NW=$nw BF_FILL3_MFE=0 BF_FILL3_DBG=$dbg python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --L $L --B $B 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['roofline']['kernel_ms']['bf_k_pf'])" 2>&1 | tail -1
done; done
