mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/scale2_bench.json 2> gpurun_out/scale2_bench.err; echo "bench2 rc=$?"; head -c 700 gpurun_out/scale2_bench.json; echo
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/eterna100.py --time 60 --replicas 20 --out gpurun_out/eterna_2gpu_r20_60s.json > gpurun_out/eterna_2gpu.log 2>&1; echo "eterna2 rc=$?"; tail -1 gpurun_out/eterna_2gpu.log | cut -c1-1500
