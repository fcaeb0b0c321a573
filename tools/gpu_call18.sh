mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu18.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu18.log
tail -12 gpurun_out/pytest_gpu18.log
timeout 300 python bench.py --steps 1000 --warmup 20 > gpurun_out/bench18_c2.json 2> gpurun_out/bench18_c2.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench18_c2.json') if l.startswith('{')][-1])
print(d['value'], d['e2e'], d.get('pipeline'))"
tail -3 gpurun_out/bench18_c2.err
timeout 200 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench18_ref.json 2> gpurun_out/bench18_ref.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench18_ref.json') if l.startswith('{')][-1])
print(d['value'], d.get('pipeline'))"
