# round-2 evidence run on one B200: tests, bench (both arms), launch list, ncu --set full of the K1 kernels, step timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_r2.log
tail -3 gpurun_out/pytest_r2.log | cut -c1-200
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err
timeout 900 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
timeout 600 python bench.py --batch 16384 --steps 50 --no-cpu-baseline --no-pipeline --no-all-configs > gpurun_out/bench_r2_b16384.json 2> gpurun_out/bench_r2_b16384.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_r2.json', 'gpurun_out/bench_r2_b16384.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'K1', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch_group'])
    for c in d.get('all_configs', []):
        print('   ', c['workload'], c['value'], c['ms_per_step'], c['k1_ms'], c['k1_tflops'])
PY
# launch list of the bench command (serialised, cold caches: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-pipeline --no-all-configs > gpurun_out/launches_r2.log 2>&1; echo "ncu launches rc=$?"
# full captures of the three K1 kernels: bench size and B = 16384
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fwd16|bwd16|wgrad16" -s 6 -c 3 -o gpurun_out/prof_r2_k1 -f python tools/ncu_step.py c2_ipw_mslr10k 4 > gpurun_out/ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fwd16|bwd16|wgrad16" -s 6 -c 3 -o gpurun_out/prof_r2_k1_b16384 -f python tools/ncu_step.py c2_ipw_mslr10k 4 16384 > gpurun_out/ncu_k1_b16384.log 2>&1; echo "ncu k1 b16384 rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"dla_reg|rank_metrics|softmax_ce_reg" -c 3 -o gpurun_out/prof_r2_k2 -f python tools/ncu_step.py c3_dla_yahoo 2 > gpurun_out/ncu_k2.log 2>&1; echo "ncu k2 rc=$?"
timeout 200 python tools/trace_step.py > gpurun_out/trace_r2_c2.txt 2>&1; tail -14 gpurun_out/trace_r2_c2.txt | cut -c1-160
for w in c3_dla_yahoo c4_lambdarank_mslr30k; do timeout 200 python tools/trace_step.py $w > gpurun_out/trace_r2_$w.txt 2>&1; done
ls -la gpurun_out/*.ncu-rep | tail -5
