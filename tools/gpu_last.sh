mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_r2.log; tail -3 gpurun_out/pytest_r2.log
timeout 600 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
timeout 200 python tools/trace_step.py > gpurun_out/trace_r2_c2.txt 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])
PY
