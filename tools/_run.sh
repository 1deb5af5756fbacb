mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
( time timeout 900 python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
( time timeout 400 python bench.py --impl reference --steps 10 --warmup 3 ) > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "bench ref rc=$?"
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[0])
print(d['value'], d['e2e']['value'], d['roofline'], d['clocks'])
for r in d['sweep']: print(r)
print(d['independent']['value'])
for r in d['independent']['sweep']: print(r)
r=json.loads(open('gpurun_out/bench_ref_final.json').read().strip().splitlines()[0])
print('REF', r['value'], r['e2e'])
PY
