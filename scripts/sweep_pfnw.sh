for L in 30 50 64 75 90; do for nw in 8 116; do
  BF_FILL3_PF_NW=$nw timeout 200 python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu --L $L 2>/dev/null | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); k=d['roofline']['kernel_ms']; print('L=$L PF_NW=$nw value',round(d['value']),'mfe %.3f pf %.3f'%(k['bf_k_mfe'],k['bf_k_pf']))
"; done; done
