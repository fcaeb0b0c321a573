mkdir -p gpurun_out
for rep in 1 2; do for pdl in 0 1; do
UB200_PDL=$pdl timeout 600 python bench.py --steps 1000 --warmup 20 --no-pipeline --no-cpu-baseline > gpurun_out/bench_pdl_$pdl.json 2> gpurun_out/bench_pdl_$pdl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pdl_$pdl.json').read().strip().splitlines()[-1])
print('PDL=$pdl', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch_group'], d['e2e']['value'], ' '.join('%s=%.4f'%(c['workload'][:10], c['ms_per_step']) for c in d['all_configs']))
PY
done; done
UB200_PDL=1 timeout 300 python bench.py --batch 16384 --steps 50 --no-cpu-baseline --no-pipeline --no-all-configs > gpurun_out/bench_pdl_b.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_pdl_b.json').read().strip().splitlines()[-1])
print('PDL=1 B=16384', d['value'], d['ms_per_step'], d['roofline']['achieved'])
PY
