mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_dp.py -x -q -m gpu > gpurun_out/pytest_dp11.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dp11.log
tail -12 gpurun_out/pytest_dp11.log
run() { echo "$1" >> gpurun_out/bench11_dp2.txt; env $1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 400 --warmup 20 $3 >> gpurun_out/bench11_dp2.txt 2>> gpurun_out/bench11_dp2.err; }
run "UB200_DP_PEER=1" 29541 ""
run "UB200_DP_PEER=0" 29542 ""
run "UB200_DP_PEER=1" 29543 "--workload c3_dla_yahoo"
grep -E "^UB200|^\{" gpurun_out/bench11_dp2.txt | cut -c1-330
grep -iE "error|trap|unavailable" gpurun_out/bench11_dp2.err | head -5
