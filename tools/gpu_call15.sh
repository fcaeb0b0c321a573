mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu15.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu15.log
tail -8 gpurun_out/pytest_gpu15.log
timeout 200 python tools/bench_kernels.py > gpurun_out/kernels15.txt 2>&1
grep K2 gpurun_out/kernels15.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"softmax_ce_reg" -s 24 -c 1 -o gpurun_out/prof_r1f_k2 python tools/bench_kernels.py > gpurun_out/ncu_k2f.log 2>&1
timeout 300 python bench.py --steps 1000 --warmup 20 > gpurun_out/bench15_c2.json 2> gpurun_out/bench15_c2.err
cut -c1-300 gpurun_out/bench15_c2.json
