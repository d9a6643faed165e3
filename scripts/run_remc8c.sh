mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29718 bench.py --workload remc --gpus 8 --steps 3 --warmup 1 > gpurun_out/r02i_remc_8gpu.json 2> gpurun_out/r02i_remc_8gpu.err
echo "remc n=8 rc=$?"; python -c "
import json
for ln in open('gpurun_out/r02i_remc_8gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print(8, d['value'], d['by_target'])
"
