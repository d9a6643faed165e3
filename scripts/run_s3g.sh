mkdir -p gpurun_out
(python scripts/design_substep.py; BF_WIDE=0 python scripts/design_substep.py; python scripts/lat2.py; BF_WIDE=0 python scripts/lat2.py | sed 's/^/WIDE=0 /') 2>&1 | tee gpurun_out/s3g_wide.log
