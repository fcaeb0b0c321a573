for cfg in "136 40 64 256,128,64" "700 20 24 512,256,128" "220 10 33 512,256,128"; do
    echo "######## $cfg"
    timeout 120 python tools/debug_f16.py $cfg 127 2>&1 | grep -E "mode|scores|linear.\.weight|linear.\.bias|Error|timed" | head -40
done
python tools/timeline_f16.py 136 40 256 256,128,64 --wgrad 2>&1 | tail -18
for c in "136 40 256 256,128,64" "136 40 16384 256,128,64" "136 200 256 512,256,128" "700 20 256 512,256,128"; do echo "## $c"; timeout 200 python tools/time_k1.py $c 127 2>&1 | tail -1; done
