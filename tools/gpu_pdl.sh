mkdir -p gpurun_out
for pdl in 0 1; do
UB200_PDL=$pdl timeout 600 python bench.py --steps 1000 --warmup 20 --no-pipeline --no-all-configs --no-cpu-baseline > gpurun_out/bench_pdl_$pdl.json 2> gpurun_out/bench_pdl_$pdl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pdl_$pdl.json').read().strip().splitlines()[-1])
print('PDL=$pdl', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch_group'], d['e2e']['value'])
PY
done
