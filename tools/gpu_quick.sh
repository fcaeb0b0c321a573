mkdir -p gpurun_out
timeout 600 python bench.py --steps 1000 --warmup 20 --no-pipeline --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['ms_per_launch_group'], d['e2e']['value'])
for c in d.get('all_configs', []): print('   ', c['workload'], c['value'], c['ms_per_step'], c.get('k1_tflops'))
PY
timeout 300 python bench.py --batch 16384 --steps 50 --no-cpu-baseline --no-pipeline --no-all-configs > gpurun_out/bench_quick_b.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick_b.json').read().strip().splitlines()[-1])
print('B=16384', d['value'], d['ms_per_step'], d['roofline']['achieved'])
PY
timeout 900 python -m pytest tests -m gpu -x -q -k "parity or graph" 2>&1 | tail -3
UB200_IMG=1 timeout 900 python -m pytest tests -m gpu -x -q -k "mlp_forward_backward or c2like or c3like" 2>&1 | tail -2
