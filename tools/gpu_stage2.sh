mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_s2.log
tail -5 gpurun_out/pytest_s2.log
timeout 600 python bench.py --steps 1000 --warmup 20 --no-all-configs > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['pipeline'])
PY
UB200_STAGE_SLOTS=1 timeout 600 python bench.py --steps 1000 --warmup 20 --no-all-configs > gpurun_out/bench_s2_slots1.json 2> gpurun_out/bench_s2_slots1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_s2_slots1.json').read().strip().splitlines()[-1])
print('slots=1', d['value'], d['ms_per_step'], d['e2e'], d['pipeline'])
PY
