mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bf_k_mfe -c 1 -s 2 -o gpurun_out/s3c_two_prof -f python scripts/two_prof.py > gpurun_out/s3c_two_prof.log 2>&1; echo rc=$?
