# GPU test-suite only
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_r2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_r2.log
grep -E "passed|failed|rc=|FAILED|Error" gpurun_out/pytest_r2.log | tail -30 | cut -c1-300
