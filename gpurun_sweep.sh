#!/bin/bash
# tuning sweep (experiments): kernel times for fill-path variants
mkdir -p gpurun_out
run() { echo "== L=$L $*"; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu --L $L 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); k=d['roofline']['kernel_ms']; print('   value',round(d['value']),'mfe %.3f pf %.3f'%(k['bf_k_mfe'],k['bf_k_pf']),'ok',all(d['checks']['ed_equals_mfe_and_epf_le_mfe'].values()))
    elif 'rror' in ln: print(ln.strip())
"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
export L=100
for nw in 2 4 8; do for pl in 0 8 4 12 6; do run BF_PF_NW=$nw BF_PF_PL=$pl BF_MFE_NW=$nw BF_MFE_PL=$((pl/4*2 + (pl&8))); done; done
export L=400
for nw in 4 8; do for pl in 0 4; do run BF_PF_NW=$nw BF_PF_PL=$pl BF_MFE_NW=$nw BF_MFE_PL=$((pl/2)); done; done
