mkdir -p gpurun_out
timeout 600 python tools/bench_kernels.py > gpurun_out/kernels_late.txt 2>&1; grep "dla\|K3" gpurun_out/kernels_late.txt | cut -c1-140
timeout 900 python -m pytest tests -m gpu -x -q -k "pairwise or prsrank or lambdarank or dla or golden" 2>&1 | tail -2
