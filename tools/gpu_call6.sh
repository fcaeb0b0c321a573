mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu6.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu6.log
tail -15 gpurun_out/pytest_gpu6.log
for w in c4_lambdarank_mslr30k c4_pairdebias_mslr30k; do python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline >> gpurun_out/bench6.json 2>> gpurun_out/bench6.err; done
python tools/trace_step.py c4_lambdarank_mslr30k > gpurun_out/trace6_c4.txt 2>&1
python tools/trace_step.py c4_pairdebias_mslr30k > gpurun_out/trace6_c4pd.txt 2>&1
cut -c1-300 gpurun_out/bench6.json
grep pairwise gpurun_out/trace6_c4.txt gpurun_out/trace6_c4pd.txt
