#!/bin/bash
# One GPU visit: parity tests, bench line (+ reference arm), ncu launch list, ncu --set full of the two fill kernels.
# usage: scripts/gpu_round.sh <tag> [L2]     (outputs under gpurun_out/<tag>_*)
tag=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.csv 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${tag}_pytest_gpu.log
tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu > gpurun_out/${tag}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bf_k_(mfe|pf)_fill' -s 2 -c 2 -f -o gpurun_out/${tag}_prof \
   python bench.py --steps 1 --warmup 3 --no-sweep --no-cpu > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
if [ -n "$2" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bf_k_(mfe|pf)_fill' -s 2 -c 2 -f -o gpurun_out/${tag}_prof$2 \
   python bench.py --steps 1 --warmup 3 --no-sweep --no-cpu --L $2 --B 592 > gpurun_out/${tag}_ncu_full$2.log 2>&1; echo "ncu full L=$2 rc=$?"
fi
head -c 3500 gpurun_out/${tag}_bench.json; echo; head -c 1500 gpurun_out/${tag}_bench_reference.json
# cluster-per-sequence kernels: ncu --set full on a small batch of 400-nt sequences (8 sequences: one round of clusters)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'bf_k_(mfe|pf)_cl|bf_k_(f5|q5)_wide' -s 8 -c 4 -f -o gpurun_out/${tag}_prof_cl \
   python scripts/cl_time.py both 400 8 > gpurun_out/${tag}_ncu_cl.log 2>&1; echo "ncu cluster rc=$?"
