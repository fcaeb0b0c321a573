mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dp.py tests/test_gpu_dp_main.py -q -x > gpurun_out/pytest_dp2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dp2.log
tail -15 gpurun_out/pytest_dp2.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 200 --warmup 5 --no-pipeline --no-all-configs > gpurun_out/bench_dp2.json 2> gpurun_out/bench_dp2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_dp2.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'dp_check', d.get('dp_check'))
for c in d.get('all_configs', []):
    print(c['workload'], c['value'], c['ms_per_step'])
PY
tail -5 gpurun_out/bench_dp2.err | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/trace_step_dp.py > gpurun_out/trace_dp2.txt 2>&1; tail -14 gpurun_out/trace_dp2.txt | cut -c1-150
grep "DP " gpurun_out/pytest_dp2.log | head
