mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_scale2_bench.json 2> gpurun_out/r02_scale2_bench.err; echo "bench2 rc=$?"
python -c "
import json
for ln in open('gpurun_out/r02_scale2_bench.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(d['n_gpus'], d['value'], d['e2e']['value'], d['by_length'], d['cofold'], d['with_ensemble_defect']['value'])
"
tail -3 gpurun_out/r02_scale2_bench.err
