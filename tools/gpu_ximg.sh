mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "parity or graph" 2>&1 | tail -6 > gpurun_out/pytest_ximg.log
tail -3 gpurun_out/pytest_ximg.log
for x in 1 0; do
UB200_XIMG=$x timeout 600 python bench.py --steps 1000 --warmup 20 --no-pipeline --no-all-configs --no-cpu-baseline > gpurun_out/bench_ximg_$x.json 2> gpurun_out/bench_ximg_$x.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_ximg_$x.json').read().strip().splitlines()[-1])
print('XIMG=$x', d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['ms_per_launch_group'], d['e2e']['value'])
PY
done
timeout 300 python tools/trace_step.py > gpurun_out/trace_ximg_c2.txt 2>&1; tail -14 gpurun_out/trace_ximg_c2.txt
