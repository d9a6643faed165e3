mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 scripts/eterna100.py --time 60 --replicas 20 --out gpurun_out/eterna_4gpu_r20_60s.json > gpurun_out/eterna_4gpu.log 2>&1; echo "eterna4 rc=$?"; tail -1 gpurun_out/eterna_4gpu.log | cut -c1-1800
