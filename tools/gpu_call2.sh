mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu2.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu2.log
tail -3 gpurun_out/pytest_gpu2.log
for cfg in "1 1" "0 1" "1 0" "0 0"; do set -- $cfg; echo "PDL=$1 OPT_FUSED=$2" >> gpurun_out/bench_ab2.txt; UB200_PDL=$1 UB200_OPT_FUSED=$2 python bench.py --steps 400 --warmup 20 --no-cpu-baseline >> gpurun_out/bench_ab2.txt 2>> gpurun_out/bench_ab2.err; done
python tests/debug_hostpack.py > gpurun_out/hostpack2.log 2>&1
cat gpurun_out/bench_ab2.txt | cut -c1-400
