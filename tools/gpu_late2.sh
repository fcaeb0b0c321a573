mkdir -p gpurun_out
timeout 600 python tools/bench_kernels.py > gpurun_out/kernels_late.txt 2>&1; grep "dla\|K3" gpurun_out/kernels_late.txt | cut -c1-140
timeout 200 python tools/trace_step.py > gpurun_out/trace_late_c2.txt 2>&1; tail -5 gpurun_out/trace_late_c2.txt | cut -c1-150
