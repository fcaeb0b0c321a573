mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu4.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu4.log
tail -3 gpurun_out/pytest_gpu4.log
python bench.py --steps 400 --warmup 20 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err
python tools/trace_step.py > gpurun_out/trace4.txt 2>&1
python tools/trace_step.py c3_dla_yahoo > gpurun_out/trace4_c3.txt 2>&1
python tools/trace_step.py c4_lambdarank_mslr30k > gpurun_out/trace4_c4.txt 2>&1
cat gpurun_out/bench4.json | cut -c1-900
tail -16 gpurun_out/trace4.txt
