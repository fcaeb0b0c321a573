mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu10.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu10.log
tail -8 gpurun_out/pytest_gpu10.log
timeout 200 python tools/bench_kernels.py > gpurun_out/kernels10.txt 2>&1
cat gpurun_out/kernels10.txt
timeout 200 python bench.py --steps 400 --warmup 20 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err
cut -c1-700 gpurun_out/bench10.json
