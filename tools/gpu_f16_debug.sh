# A/B of the fp16-split K1 kernels against the CUDA-core path on a few shapes (each in its own process + timeout)
mkdir -p gpurun_out
for cfg in "136 40 64 256,128,64" "136 40 256 256,128,64" "700 20 24 512,256,128" "136 200 16 512,256,128" "220 10 33 512,256,128" "136 1 300 64" "136 7 100 128,64"; do
  echo "######## $cfg" 
  timeout 120 python tools/debug_f16.py $cfg ${MODES:-127,255} 2>&1 | tail -40
done > gpurun_out/f16_debug.log 2>&1
grep -c "nan=[1-9]" gpurun_out/f16_debug.log
grep "####\|mode .* done\|Traceback\|rror" gpurun_out/f16_debug.log | head -60
