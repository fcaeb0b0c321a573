mkdir -p gpurun_out
timeout 200 python tools/trace_step.py > gpurun_out/trace_late_c2.txt 2>&1; tail -6 gpurun_out/trace_late_c2.txt | cut -c1-150
timeout 900 python -m pytest tests -m gpu -x -q -k "parity or golden" 2>&1 | tail -2
timeout 600 python bench.py --steps 1000 --warmup 20 --no-pipeline --no-cpu-baseline --no-all-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'])"
