mkdir -p gpurun_out
for w in 36 104; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s3x_design_launches_$w.csv python scripts/design_launches.py $w > /dev/null 2>&1; echo rc=$?
done
