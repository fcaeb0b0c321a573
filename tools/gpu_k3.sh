mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pairwise or prsrank or lambdarank or golden" 2>&1 | tail -5 > gpurun_out/pytest_k3.log
tail -3 gpurun_out/pytest_k3.log
timeout 600 python tools/bench_kernels.py > gpurun_out/kernels_k3.txt 2>&1
grep "K3" gpurun_out/kernels_k3.txt
