mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2; do
  NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --workload remc --gpus $n --steps 3 --warmup 1 > gpurun_out/r02b_remc_${n}gpu.json 2> gpurun_out/r02b_remc_${n}gpu.err
  echo "remc n=$n rc=$?"; python -c "
import json,sys
for ln in open('gpurun_out/r02b_remc_${n}gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print($n, d['value'], d['by_target'])
"
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 scripts/eterna100.py --time 60 --replicas 10 --out gpurun_out/r02b_eterna100_8gpu_r10_60s.json > gpurun_out/r02b_eterna_8gpu.log 2>&1; echo "eterna8 rc=$?"; tail -1 gpurun_out/r02b_eterna_8gpu.log | cut -c1-900
