mkdir -p gpurun_out
(
BF_CFG_PRINT=1 python scripts/ovl.py base 2>&1 | sort | uniq | tail -4
BF_SCORE_OVERLAP=2 python scripts/ovl.py ovl_nocap
BF_SCORE_OVERLAP=2 BF_FILL3_MFE_CTAS=2 BF_FILL3_PF_NW=8 BF_FILL3_PF_CTAS=1 BF_CFG_PRINT=1 python scripts/ovl.py ovl_m2_p8x1 2>&1 | sort | uniq | tail -4
BF_SCORE_OVERLAP=2 BF_FILL3_MFE_CTAS=1 BF_FILL3_PF_NW=8 BF_FILL3_PF_CTAS=2 python scripts/ovl.py ovl_m1_p8x2
BF_SCORE_OVERLAP=2 BF_FILL3_MFE_CTAS=1 BF_FILL3_PF_NW=8 BF_FILL3_PF_CTAS=1 python scripts/ovl.py ovl_m1_p8x1
BF_SCORE_OVERLAP=2 BF_FILL3_MFE_CTAS=2 BF_FILL3_PF_NW=8 BF_FILL3_PF_CTAS=2 python scripts/ovl.py ovl_m2_p8x2
BF_FILL3_PF_NW=8 python scripts/ovl.py seq_p8
) 2>&1 | tee gpurun_out/s3o_ovl.log
