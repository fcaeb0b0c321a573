mkdir -p gpurun_out
UB200_IMG=1 timeout 1500 python -m pytest tests -m gpu -x -q -k "parity or graph" 2>&1 | tail -8 > gpurun_out/pytest_img.log
tail -3 gpurun_out/pytest_img.log
for m in 1 0; do
UB200_IMG=$m timeout 600 python bench.py --steps 500 --warmup 20 --no-pipeline > gpurun_out/bench_img_$m.json 2> gpurun_out/bench_img_$m.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_img_$m.json').read().strip().splitlines()[-1])
print('IMG=$m', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['ms_per_launch_group'], d['e2e']['value'])
for c in d.get('all_configs', []): print('   ', c['workload'], c['value'], c['ms_per_step'], c.get('k1_tflops'))
PY
done
UB200_IMG=1 timeout 300 python tools/trace_step.py > gpurun_out/trace_img_c2.txt 2>&1; tail -13 gpurun_out/trace_img_c2.txt
