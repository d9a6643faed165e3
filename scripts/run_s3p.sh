bash scripts/gpu_round.sh r02f 400
timeout 300 python bench.py --workload remc --steps 3 --warmup 1 > gpurun_out/r02f_remc_1gpu.json 2> gpurun_out/r02f_remc_1gpu.err; echo "remc rc=$?"
