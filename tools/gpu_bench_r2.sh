mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'repeats', d.get('repeats'), 'e2e', d['e2e'])
print('roofline', {k: d['roofline'][k] for k in ('achieved','frac','frac_of_split_ceiling','tf32_dense_peak_measured','f16_dense_peak_measured','ms_per_launch_group','fwd_ms')})
print('pipeline', d.get('pipeline'))
print('cpu', d.get('cpu_baseline'))
for c in d.get('all_configs', []):
    print(c['workload'], c['value'], c['ms_per_step'], c['k1_ms'], c['k1_tflops'])
PY
tail -5 gpurun_out/bench_r2.err
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/ncu_step_c2.csv python tools/ncu_step.py c2_ipw_mslr10k 4 > gpurun_out/ncu_step_c2.log 2>&1; echo "ncu rc=$?"
python tools/ncu_traffic.py gpurun_out/ncu_step_c2.csv c2_ipw_mslr10k 256 4 | tail -20
cp profiles/r02_traffic.json gpurun_out/
