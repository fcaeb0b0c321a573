mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu20.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu20.log
tail -15 gpurun_out/pytest_gpu20.log | cut -c1-400
