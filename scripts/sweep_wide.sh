for L in 200 300 400; do for w in 1 2; do
  BF_WIDE=$w BF_FILL3=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-sweep --no-cpu --L $L --B 1024 2>/dev/null | python -c "
import sys,json
for ln in sys.stdin:
    if ln.startswith('{'):
        d=json.loads(ln); k=d['roofline']['kernel_ms']; print('L=$L BF_WIDE=$w value',round(d['value']),'mfe %.3f pf %.3f'%(k['bf_k_mfe'],k['bf_k_pf']), d['checks']['ed_equals_mfe_and_epf_le_mfe'])
"; done; done
