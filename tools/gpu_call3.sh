mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu3.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu3.log
tail -3 gpurun_out/pytest_gpu3.log
python tests/debug_hostpack.py > gpurun_out/hostpack3.log 2>&1
python bench.py --steps 400 --warmup 20 --no-cpu-baseline > gpurun_out/bench3.json 2> gpurun_out/bench3.err
UB200_PDL=0 python tools/trace_step.py > gpurun_out/trace3_nopdl.txt 2>&1
UB200_PDL=1 python tools/trace_step.py > gpurun_out/trace3_pdl.txt 2>&1
cat gpurun_out/bench3.json | cut -c1-900
