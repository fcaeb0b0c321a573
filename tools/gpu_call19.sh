mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_resident.py tests/test_online_feed.py -m gpu -q > gpurun_out/pytest_gpu19.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu19.log
tail -15 gpurun_out/pytest_gpu19.log | cut -c1-300
timeout 300 python bench.py --steps 400 --warmup 20 --no-cpu-baseline > gpurun_out/bench19_c2.json 2> gpurun_out/bench19_c2.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench19_c2.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], json.dumps(d.get('pipeline')))"
tail -3 gpurun_out/bench19_c2.err
