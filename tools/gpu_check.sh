# full GPU suite + the default bench line (one B200)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_check.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_check.log
tail -8 gpurun_out/pytest_check.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_check.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['kernels_per_step'], json.dumps(d.get('pipeline')))"
UB200_EARLY_LOSS=0 timeout 300 python bench.py --no-cpu-baseline --no-pipeline > gpurun_out/bench_check_noearly.json 2>> gpurun_out/bench_check.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_check_noearly.json').read().strip().splitlines()[-1])
print('no early:', d['value'], d['ms_per_step'], d['e2e'], d['kernels_per_step'])"
wc -l gpurun_out/bench_check.json
