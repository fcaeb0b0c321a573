mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"^(?!.*at::).*" -s 60 -c 60 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"^(?!.*at::).*" -s 60 -c 15 -o gpurun_out/prof_r1b_full python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_f.log 2>&1
for w in c1_na_toy c3_dla_yahoo c4_lambdarank_mslr30k c4_pairdebias_mslr30k c5_dla_istella; do python bench.py --workload $w --steps 50 --warmup 5 --no-cpu-baseline >> gpurun_out/bench_others.json 2>> gpurun_out/bench_others.err; done
python bench.py --batch 16384 --steps 20 --warmup 3 --no-cpu-baseline >> gpurun_out/bench_others.json 2>> gpurun_out/bench_others.err
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_c2.json
