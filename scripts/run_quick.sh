mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for w in 1 0; do BF_PF_WIDE_RING=$w python scripts/latency.py > gpurun_out/quick_latency_$w.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/quick_latency_$w.json'))['latency']; print('WIDE_RING=$w', {k:(v['call_ms'],v['pf_ms']) for k,v in d.items() if k.startswith('L400') or k.startswith('L200')})"; done
