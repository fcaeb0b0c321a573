mkdir -p gpurun_out
for args in "136 40 256 256,128,64" "136 40 2048 256,128,64" "136 200 256 512,256,128"; do
  echo "######## $args"
  timeout 200 python tools/timeline_f16.py $args --bwd --wgrad 2>&1 | tail -70
done > gpurun_out/tl_r2b.log 2>&1
tail -5 gpurun_out/tl_r2b.log
