mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_dp.py -x -q -m gpu > gpurun_out/pytest_dp9.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dp9.log
tail -12 gpurun_out/pytest_dp9.log
for peer in 1 0; do
echo "UB200_DP_PEER=$peer" >> gpurun_out/bench9_dp2.txt
UB200_DP_PEER=$peer timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 400 --warmup 20 >> gpurun_out/bench9_dp2.txt 2>> gpurun_out/bench9_dp2.err
done
grep -E "^UB200|^\{" gpurun_out/bench9_dp2.txt | cut -c1-420
tail -5 gpurun_out/bench9_dp2.err
