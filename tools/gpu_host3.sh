mkdir -p gpurun_out
for env in "UB200_X=1"; do
  echo "#### $env"
  env $env timeout 200 python tools/host_profile.py 2>&1 | grep "median\|train() loop\|stage alone\|inside"
done > gpurun_out/host3b.log 2>&1
timeout 600 python bench.py --steps 1000 --warmup 20 --no-all-configs > gpurun_out/bench_s3.json 2> gpurun_out/bench_s3.err
python - <<'PY' >> gpurun_out/host3b.log
import json
d=json.loads(open('gpurun_out/bench_s3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['pipeline'])
PY
