mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "golden or resident or graph or dropin or train_steps" 2>&1 | tail -2
timeout 600 python bench.py --steps 1000 --warmup 20 --no-cpu-baseline --no-all-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['pipeline'])"
timeout 200 python tools/host_profile.py 2>&1 | grep "median\|train() loop\|inside"
