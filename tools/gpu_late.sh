mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python tools/bench_kernels.py > gpurun_out/kernels_late.txt 2>&1; grep "dla\|K3\|IPW" gpurun_out/kernels_late.txt | cut -c1-140
for pdl in 0 1; do
UB200_PDL=$pdl timeout 600 python bench.py --steps 1000 --warmup 20 --no-pipeline --no-cpu-baseline > gpurun_out/bench_pdl_$pdl.json 2> gpurun_out/bench_pdl_$pdl.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_pdl_$pdl.json').read().strip().splitlines()[-1])
print('PDL=$pdl', d['value'], d['ms_per_step'], d['roofline']['ms_per_launch_group'], d['e2e']['value'], ' '.join('%s=%.4f'%(c['workload'][:10], c['ms_per_step']) for c in d['all_configs']))
PY
done
timeout 200 python tools/trace_step.py > gpurun_out/trace_late_c2.txt 2>&1; tail -13 gpurun_out/trace_late_c2.txt | cut -c1-150
timeout 200 python tools/trace_step.py c4_lambdarank_mslr30k > gpurun_out/trace_late_c4.txt 2>&1; tail -16 gpurun_out/trace_late_c4.txt | cut -c1-150
