mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bf_k_mfe|bf_k_pf" -c 2 -s 2 -o gpurun_out/s3e_two_big -f python scripts/two_prof_big.py > gpurun_out/s3e_two_big.log 2>&1; echo rc=$?
