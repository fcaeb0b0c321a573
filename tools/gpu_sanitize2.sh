# memcheck over the kernels changed late in round 2: K1 with the weight-gradient operand images forced on for every
# shape (UB200_IMG=1), the DLA loss, the pairwise losses, staging (ipw golden through train())
mkdir -p gpurun_out
SEL='mlp_forward_backward_vs_oracle or dla_loss or (pairwise_vs_oracle and 64-40) or prsrank or ipw_c2like or dla_wide or lambdarank_c4like'
UB200_IMG=1 timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 77 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/sanitizer2_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/sanitizer2_memcheck.log
grep -E "ERROR SUMMARY|passed|failed|Invalid|Out-of-range|misaligned" gpurun_out/sanitizer2_memcheck.log | tail -8
