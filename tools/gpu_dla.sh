mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "dla or golden or validation or dropin" 2>&1 | tail -5 > gpurun_out/pytest_dla.log
tail -3 gpurun_out/pytest_dla.log
timeout 600 python tools/bench_kernels.py > gpurun_out/kernels_dla.txt 2>&1
grep -i "dla" gpurun_out/kernels_dla.txt
