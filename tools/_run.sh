mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_v5.json 2> gpurun_out/bench_v5.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_v5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v5.json').read().strip().splitlines()[0])
print(d['value'], d['e2e'], d['roofline'])
for r in d['sweep']: print(r)
print(d['independent']['value'])
for r in d['independent']['sweep']: print(r)
PY
