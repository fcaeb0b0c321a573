mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "l2 or ipw_small or dla_small or pairdebias_small" 2>&1 | tail -25 > gpurun_out/pytest_l2.log
tail -25 gpurun_out/pytest_l2.log | cut -c1-220
