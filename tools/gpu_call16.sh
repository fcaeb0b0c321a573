mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { echo "$1 N=$2" >> gpurun_out/bench16_dp.txt; env $1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $2 --steps 400 --warmup 20 >> gpurun_out/bench16_dp.txt 2>> gpurun_out/bench16_dp.err; echo "rc=$?" >> gpurun_out/bench16_dp.txt; }
run "UB200_DP_PEER=1" 8 29551
run "UB200_DP_PEER=0" 8 29552
run "UB200_DP_PEER=1" 4 29553
run "UB200_DP_PEER=1" 2 29554
python - <<'PY'
import json
for l in open('gpurun_out/bench16_dp.txt'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['kernels_per_step'])
    elif l.startswith('UB200') or l.startswith('rc='): print(l.strip())
PY
grep -iE "error|trap|unavailable|Traceback" gpurun_out/bench16_dp.err | head -8
