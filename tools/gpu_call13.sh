mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "softmax_ce or dla_loss or golden or train_steps" > gpurun_out/pytest_gpu13.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu13.log
tail -4 gpurun_out/pytest_gpu13.log
timeout 200 python tools/bench_kernels.py > gpurun_out/kernels13.txt 2>&1
grep K2 gpurun_out/kernels13.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"softmax_ce_reg" -s 24 -c 2 -o gpurun_out/prof_r1d_k2 python tools/bench_kernels.py > gpurun_out/ncu_k2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"^(?!.*at::).*" -s 56 -c 56 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_l13.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"^(?!.*at::).*" -s 56 -c 14 -o gpurun_out/prof_r1d_full python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_f13.log 2>&1
python tools/trace_step.py > gpurun_out/trace13_c2.txt 2>&1
