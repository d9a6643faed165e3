mkdir -p gpurun_out
(python scripts/lat3.py wide16
BF_CL=0 BF_WIDE=0 python scripts/lat3.py f3_8
BF_CL=0 BF_WIDE=0 BF_FILL3_NW=16 BF_FILL3_NWI=12 BF_FILL3_MFE_KB=220 BF_FILL3_PF_NW=16 python scripts/lat3.py f3_16_12
BF_CL=0 BF_WIDE=0 BF_FILL3_NW=12 BF_FILL3_NWI=9 BF_FILL3_MFE_KB=220 BF_FILL3_PF_NW=116 python scripts/lat3.py f3_12_9_pf116
BF_CL=0 BF_WIDE=0 BF_FILL3_NW=8 BF_FILL3_MFE_KB=220 BF_FILL3_PF_NW=112 BF_FILL3_PF_NWI=9 python scripts/lat3.py f3_8big_pf112
) 2>&1 | tee gpurun_out/s3h_lat3.log
