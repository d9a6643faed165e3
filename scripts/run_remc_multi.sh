mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 1 2 4; do
  if [ $n -eq 1 ]; then timeout 300 python bench.py --workload remc --steps 3 --warmup 1 > gpurun_out/r02_remc_${n}gpu.json 2> gpurun_out/r02_remc_${n}gpu.err
  else NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --workload remc --gpus $n --steps 3 --warmup 1 > gpurun_out/r02_remc_${n}gpu.json 2> gpurun_out/r02_remc_${n}gpu.err; fi
  echo "remc n=$n rc=$?"; python -c "
import json,sys
for ln in open('gpurun_out/r02_remc_${n}gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print($n, d['by_target'])
"
done
