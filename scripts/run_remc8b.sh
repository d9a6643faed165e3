mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4 2 1; do
  if [ $n -eq 1 ]; then timeout 240 python bench.py --workload remc --steps 3 --warmup 1 > gpurun_out/r02h_remc_${n}gpu.json 2> gpurun_out/r02h_remc_${n}gpu.err
  else NCCL_DEBUG=WARN timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --workload remc --gpus $n --steps 3 --warmup 1 > gpurun_out/r02h_remc_${n}gpu.json 2> gpurun_out/r02h_remc_${n}gpu.err; fi
  echo "remc n=$n rc=$?"; python -c "
import json,sys
for ln in open('gpurun_out/r02h_remc_${n}gpu.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print($n, d['value'], d['by_target'])
"
done
