# the driver's own commands: default bench at N = 2 (torchrun) and N = 1, every informational leg on
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/bench_default_n2.json 2> gpurun_out/bench_default_n2.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default_n2.json').read().strip().splitlines()[-1])
print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d.get('pipeline'), d['dp_check']['replicas_bitwise_equal'], d['dp_check']['vs_single_gpu'])
for c in d.get('all_configs', []): print('   ', c.get('workload'), c.get('value'), c.get('ms_per_step'), c.get('error'))
PY
tail -3 gpurun_out/bench_default_n2.err | cut -c1-200
( time timeout 900 python bench.py > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default_n1.json').read().strip().splitlines()[-1])
print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['steps'], d['warmup'])
PY
