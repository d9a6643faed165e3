#!/bin/bash
# v3 against the round-1 kernels by length
for L in 30 50 75 100 120 150 200 300 400; do
  B=4096; [ $L -ge 300 ] && B=1024
  for q in "BF_FILL3=0" "BF_FILL3=1" "BF_FILL3=1 BF_FILL3_PF_NW=8"; do
    echo -n "L=$L B=$B $q : "
    env $q python bench.py --steps 3 --warmup 2 --no-sweep --no-cpu --L $L --B $B 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), d['roofline']['kernel_ms'])" 2>&1 | tail -1
  done
done
