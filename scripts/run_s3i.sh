mkdir -p gpurun_out
export LAT_LS=170,200,250,300 LAT_BS=16,64,128
(python scripts/lat3.py default
BF_CL=0 python scripts/lat3.py wide16
BF_CL=0 BF_WIDE=0 BF_FILL3_NW=16 BF_FILL3_NWI=12 BF_FILL3_MFE_KB=220 BF_FILL3_PF_NW=16 BF_FILL3_PF_MAXN=400 python scripts/lat3.py f3_16_12
BF_CL=0 BF_WIDE=0 BF_FILL3_NW=16 BF_FILL3_NWI=12 BF_FILL3_MFE_KB=220 BF_FILL3_PF_NW=116 BF_FILL3_PF_MAXN=400 python scripts/lat3.py f3_16_12_pf116
) 2>&1 | tee gpurun_out/s3i_lat3.log
