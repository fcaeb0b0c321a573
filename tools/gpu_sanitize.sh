# compute-sanitizer over a subset of the GPU tests (the tools slow the kernels 10-100x): memcheck on the parity tests of
# every kernel family, racecheck + synccheck on the K1 kernels (hand-rolled mbarrier / TMEM / grid-barrier protocols)
mkdir -p gpurun_out
SEL='mlp_forward_backward_vs_oracle or ipw_c2like or dla_wide or rank_metrics or (softmax_ce_vs_oracle and 33) or clip_update or resident_dataset'
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 77 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|Out-of-range|misaligned" gpurun_out/sanitizer_memcheck.log | tail -8
SEL2='(mlp_forward_backward_vs_oracle and (136-hidden1 or 13-hidden0)) or ipw_small'
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL2" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/sanitizer_racecheck.log
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitizer_racecheck.log | tail -8
timeout 1200 compute-sanitizer --tool synccheck --print-limit 20 --error-exitcode 77 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL2" > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" | tee -a gpurun_out/sanitizer_synccheck.log
grep -E "ERROR SUMMARY|passed|failed|Barrier|divergent" gpurun_out/sanitizer_synccheck.log | tail -8
