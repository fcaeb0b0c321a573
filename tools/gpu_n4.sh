mkdir -p gpurun_out
nvidia-smi nvlink -gt d > gpurun_out/nvlink_before.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 4 --steps 500 --warmup 10 --no-pipeline --no-all-configs > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench rc=$?"
nvidia-smi nvlink -gt d > gpurun_out/nvlink_after.txt 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n4.json').read().strip().splitlines()[-1])
print('N=4 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'dp_check', d.get('dp_check'), 'launches', d.get('gpu_launches'))
PY
tail -3 gpurun_out/bench_n4.err | cut -c1-300
nproc
