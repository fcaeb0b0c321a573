mkdir -p gpurun_out
for env in "UB200_PACK_NT=1 UB200_PACK_ISSUER=0" "UB200_PACK_NT=0 UB200_PACK_ISSUER=0" "UB200_PACK_NT=1 UB200_PACK_ISSUER=1" "UB200_PACK_NT=0 UB200_PACK_ISSUER=1"; do
  echo "#### $env"
  env $env timeout 200 python tools/host_profile2.py 2>&1 | grep "stage_feed\|convert rot"
  env $env timeout 200 python tools/host_profile.py 2>&1 | grep "median\|train() loop\|stage alone"
done > gpurun_out/host2.log 2>&1
