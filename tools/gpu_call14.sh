mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "softmax_ce or golden or train_steps or validation" > gpurun_out/pytest_gpu14.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu14.log
tail -4 gpurun_out/pytest_gpu14.log
timeout 200 python tools/bench_kernels.py > gpurun_out/kernels14.txt 2>&1
grep K2 gpurun_out/kernels14.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"softmax_ce_reg" -s 24 -c 2 -o gpurun_out/prof_r1e_k2 python tools/bench_kernels.py > gpurun_out/ncu_k2e.log 2>&1
timeout 100 python bench.py --steps 400 --warmup 20 --no-cpu-baseline > gpurun_out/bench14.json 2> gpurun_out/bench14.err
cut -c1-200 gpurun_out/bench14.json
