#!/bin/bash
# kernel-time sweep of the tile-wavefront fill kernels (experiments; kernel ms from CUDA events inside bench.py)
run() { echo "== L=$L $*"; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu --L $L 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); k=d['roofline']['kernel_ms']; print('   value',round(d['value']),'mfe %.3f pf %.3f'%(k['bf_k_mfe'],k['bf_k_pf']),'ok',all(d['checks']['ed_equals_mfe_and_epf_le_mfe'].values()))
    elif 'rror' in ln: print(ln.strip())
"; }
for L in $LENGTHS; do
export L
run BF_FILL=diag
for nw in 8 12 16; do for nws in 1 2; do run BF_FILL=tile BF_TILE_NW=$nw BF_TILE_NWS=$nws; done; done
done
