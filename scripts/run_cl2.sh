for cfg in "BF_CL=0" "BF_CL=1 BF_CL_C=8" "BF_CL=1 BF_CL_C=16"; do python scripts/cl_time.py mfe 400 18 $cfg; done
for cfg in "BF_CL=0" "BF_CL=1 BF_CL_C=8"; do python scripts/cl_time.py mfe 400 1024 $cfg; python scripts/cl_time.py mfe 200 64 $cfg; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'bf_k_mfe_cl' -c 1 -f -o gpurun_out/cl_mfe python scripts/cl_time.py mfe 400 18 BF_CL=1 BF_CL_C=8 > gpurun_out/cl_mfe_ncu.log 2>&1
