# A/B of the fp16-split K1 kernels against the CUDA-core path on a few shapes (each in its own process + timeout)
mkdir -p gpurun_out
for cfg in "136 40 64 256,128,64" "136 40 256 256,128,64" "700 20 24 512,256,128" "136 200 16 512,256,128" "220 10 33 512,256,128" "136 1 300 64"; do
  echo "######## $cfg" 
  timeout 120 python tools/debug_f16.py $cfg ${MODES:-15,31,63} 2>&1 | tail -60
done > gpurun_out/f16_debug.log 2>&1
tail -150 gpurun_out/f16_debug.log
