mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu12.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu12.log
tail -8 gpurun_out/pytest_gpu12.log
timeout 200 python tools/bench_kernels.py > gpurun_out/kernels12.txt 2>&1
grep K2 gpurun_out/kernels12.txt
timeout 200 python bench.py --steps 400 --warmup 20 --no-cpu-baseline > gpurun_out/bench12.json 2> gpurun_out/bench12.err
cut -c1-700 gpurun_out/bench12.json
