mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bf_k_pf_out -c 1 -s 1 -o gpurun_out/s3r_out_prof -f python scripts/out_prof.py > gpurun_out/s3r_out_prof.log 2>&1; echo rc=$?
