mkdir -p gpurun_out
export OUT_LS=40,50,60,65,70,75,80,85,90,95,100,104
(BF_OUT_HSM=1 python scripts/out_time_dev.py hsm; BF_OUT_HSM=0 python scripts/out_time_dev.py l2; OUT_B=64 OUT_LS=36,70,100 BF_OUT_HSM=1 python scripts/out_time_dev.py hsm; OUT_B=64 OUT_LS=36,70,100 BF_OUT_HSM=0 python scripts/out_time_dev.py l2) 2>&1 | tee gpurun_out/s3u_out_time.log
