# 2-GPU data-parallel parity test + smoke (run through `gpurun --gpus 2`)
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_dp.py -x -q -m gpu > gpurun_out/pytest_dp.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_dp.log
tail -6 gpurun_out/pytest_dp.log
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log
tail -4 gpurun_out/smoke.log
