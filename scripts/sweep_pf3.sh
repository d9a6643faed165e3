#!/bin/bash
L=${1:-100}; B=${2:-4096}
for q in "BF_FILL3_PF_NW=16" "BF_FILL3_PF_NW=116 BF_FILL3_PF_NWI=12" "BF_FILL3_PF_NW=116 BF_FILL3_PF_NWI=10"; do
  echo -n "cfg: $q : "
  env $q python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu --L $L --B $B 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['kernel_ms'])" 2>&1 | tail -1
done
