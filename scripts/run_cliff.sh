mkdir -p gpurun_out
export LAT_LS=172,176,180,190,200 LAT_BS=64
(python scripts/lat3.py default
BF_FILL3_PF_MAXN_SMALL=230 python scripts/lat3.py f3_116_10
BF_FILL3_PF_MAXN_SMALL=230 BF_FILL3_PF_NW=16 BF_FILL3_PF_NWI=12 python scripts/lat3.py f3_16_12
BF_FILL3_PF_MAXN_SMALL=230 BF_FILL3_PF_NW=16 BF_FILL3_PF_NWI=10 python scripts/lat3.py f3_16_10
BF_FILL3_PF_MAXN_SMALL=230 BF_FILL3_PF_NW=112 BF_FILL3_PF_NWI=8 BF_CFG_PRINT=1 python scripts/lat3.py f3_112_8 2>&1 | sort -u | tail -12
) 2>&1 | tee gpurun_out/s3z_cliff.log
