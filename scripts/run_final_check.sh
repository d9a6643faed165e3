mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
python bench.py --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['cpu_baseline']['value'], d['by_length'], d['with_ensemble_defect']['value'], {k: v['value'] for k, v in d['cofold'].items()}, {k: (v.get('ms_per_substep') if isinstance(v, dict) else None) for k, v in d['design_loop'].items()})"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>/dev/null; echo "reference rc=$?"; head -c 300 gpurun_out/final_bench_reference.json; echo
