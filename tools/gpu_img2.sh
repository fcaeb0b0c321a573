mkdir -p gpurun_out
for img in 0 1; do
echo "### UB200_IMG=$img"
UB200_IMG=$img timeout 120 python tools/debug_f16.py 136 40 256 256,128,64 255 2>&1 | grep "done"
UB200_IMG=$img timeout 120 python tools/debug_f16.py 136 200 256 512,256,128 255 2>&1 | grep "done"
done > gpurun_out/img2.log 2>&1
cat gpurun_out/img2.log
