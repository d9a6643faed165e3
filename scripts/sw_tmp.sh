#!/bin/bash
run() { echo "== L=$L $*"; env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu --L $L 2>&1 | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); k=d['roofline']['kernel_ms']; print('   value',round(d['value']),'mfe %.3f pf %.3f'%(k['bf_k_mfe'],k['bf_k_pf']),'ok',all(d['checks']['ed_equals_mfe_and_epf_le_mfe'].values()))
    elif 'rror' in ln: print(ln.strip())
"; }
export L=200
run BF_FILL=tile BF_TILE_NW=8 BF_TILE_NWS=1
run BF_FILL=tile BF_TILE_NW=8 BF_TILE_NWS=2
run BF_FILL=tile BF_TILE_NW=16 BF_TILE_NWS=2
run BF_FILL=tile BF_TILE_NW=8 BF_TILE_NWS=1 BF_TILE_SMEM_MAX=230000
export L=400
run BF_FILL=tile BF_TILE_NW=8 BF_TILE_NWS=1
run BF_FILL=tile BF_TILE_NW=8 BF_TILE_NWS=2
run BF_FILL=tile BF_TILE_NW=16 BF_TILE_NWS=2
run BF_FILL=tile BF_TILE_NW=12 BF_TILE_NWS=2
export L=50
run BF_FILL=tile BF_TILE_NW=8 BF_TILE_NWS=1
