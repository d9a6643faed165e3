timeout 600 python -m pytest tests -m gpu -x -q -k "two_strand or dimer or golden or avoid or ragged" 2>&1 | tail -3
python - <<'P'
import bench, json, os
r = bench.bench_design_loop()
print(json.dumps({k: r[k] for k in r if k != "eterna100_x_10_replicas"}))
P
