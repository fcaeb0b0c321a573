mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"^(?!.*at::).*" -s 56 -c 56 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_l8.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"^(?!.*at::).*" -s 56 -c 14 -o gpurun_out/prof_r1c_full python bench.py --steps 4 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_f8.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pairwise_kernel" -s 4 -c 1 -o gpurun_out/prof_r1c_k3 python bench.py --workload c4_lambdarank_mslr30k --steps 4 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/ncu_k3.log 2>&1
python bench.py --steps 1000 --warmup 20 > gpurun_out/bench8_c2.json 2> gpurun_out/bench8_c2.err
for w in c1_na_toy c3_dla_yahoo c4_lambdarank_mslr30k c4_pairdebias_mslr30k c5_dla_istella; do python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline >> gpurun_out/bench8_others.json 2>> gpurun_out/bench8_others.err; done
python bench.py --batch 16384 --steps 20 --warmup 3 --no-cpu-baseline >> gpurun_out/bench8_others.json 2>> gpurun_out/bench8_others.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench8_ref.json 2> gpurun_out/bench8_ref.err
cat gpurun_out/bench8_c2.json
