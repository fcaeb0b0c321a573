# weak-scaling bench on one node (run through `gpurun --gpus 8`): N = 8 / 4 / 2 with the fused peer-memory exchange, and
# N = 8 with ncclAllReduce for comparison -> gpurun_out/bench_scaling.txt
mkdir -p gpurun_out
run() { echo "$1 N=$2" >> gpurun_out/bench_scaling.txt; env $1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus $2 --steps 400 --warmup 20 >> gpurun_out/bench_scaling.txt 2>> gpurun_out/bench_scaling.err; echo "rc=$?" >> gpurun_out/bench_scaling.txt; }
run "UB200_DP_PEER=1" 8 29551
run "UB200_DP_PEER=0" 8 29552
run "UB200_DP_PEER=1" 4 29553
run "UB200_DP_PEER=1" 2 29554
