# final measurement script of round 1 (one B200): tests, bench lines, ncu captures, timeline -> gpurun_out/*_r1g*
mkdir -p gpurun_out
T=r1g
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_$T.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_$T.log
python bench.py --steps 1000 --warmup 20 > gpurun_out/bench_${T}_c2.json 2> gpurun_out/bench_${T}_c2.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_${T}_ref.json 2> gpurun_out/bench_${T}_ref.err
for w in c1_na_toy c3_dla_yahoo c4_lambdarank_mslr30k c4_pairdebias_mslr30k c5_dla_istella; do python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline >> gpurun_out/bench_${T}_others.json 2>> gpurun_out/bench_${T}_others.err; done
python bench.py --batch 16384 --steps 20 --warmup 3 --no-cpu-baseline >> gpurun_out/bench_${T}_others.json 2>> gpurun_out/bench_${T}_others.err
python tools/bench_kernels.py > gpurun_out/kernels_$T.txt 2>&1
python tools/trace_step.py > gpurun_out/trace_${T}_c2.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"^(?!.*at::).*" -s 56 -c 56 --csv --log-file gpurun_out/launches_$T.csv python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_l_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"^(?!.*at::).*" -s 56 -c 14 -o gpurun_out/prof_${T}_full python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_f_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"softmax_ce_reg" -s 24 -c 1 -o gpurun_out/prof_${T}_k2 python tools/bench_kernels.py > gpurun_out/ncu_k2_$T.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pairwise_kernel" -s 4 -c 1 -o gpurun_out/prof_${T}_k3 python bench.py --workload c4_lambdarank_mslr30k --steps 4 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_k3_$T.log 2>&1
tail -3 gpurun_out/pytest_$T.log; cut -c1-400 gpurun_out/bench_${T}_c2.json
