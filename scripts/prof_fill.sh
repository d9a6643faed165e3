#!/bin/bash
# ncu --set full of the fill kernels on the bench workload.  usage: scripts/prof_fill.sh <tag> [L] [B] [kernel regex]
tag=${1:-prof}; L=${2:-100}; B=${3:-4096}; K=${4:-'bf_k_(mfe|pf)_fill'}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 2 -c 2 -f -o gpurun_out/${tag} \
   python bench.py --steps 1 --warmup 3 --no-sweep --no-cpu --L $L --B $B > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/ncu_summary.py gpurun_out/${tag}.ncu-rep > gpurun_out/${tag}_summary.txt 2>&1
cat gpurun_out/${tag}_summary.txt
