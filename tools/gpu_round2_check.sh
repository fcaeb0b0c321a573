# full GPU suite + the default bench line + step timeline (one B200)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_r2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_r2.log
tail -15 gpurun_out/pytest_r2.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline --no-pipeline > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['kernels_per_step'], d['roofline']['achieved'])"
timeout 200 python tools/trace_step.py > gpurun_out/trace_r2_c2.txt 2>&1; tail -25 gpurun_out/trace_r2_c2.txt | cut -c1-200
