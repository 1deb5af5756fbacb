mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
( time timeout 900 python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
( time timeout 400 python bench.py --impl reference --steps 10 --warmup 3 ) > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "bench ref rc=$?"
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:umma -c 900 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-model > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
for m in 1 256 512; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma -s 10 -c 2 -f -o gpurun_out/full_M$m python tools/ncu_one.py $m > gpurun_out/ncu_full_M$m.log 2>&1; echo "ncu full M=$m rc=$?"; done
for m in 1 256; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma -s 10 -c 2 -f -o gpurun_out/full_indep_M$m python tools/ncu_one.py $m indep > gpurun_out/ncu_full_indep_M$m.log 2>&1; echo "ncu full indep M=$m rc=$?"; done
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[0])
print(d['value'], d['e2e']['value'], d['independent']['value'], d['clocks'], d['llama2_7b_tokens_per_s'])
for r in d['sweep']: print(r)
r=json.loads(open('gpurun_out/bench_ref_final.json').read().strip().splitlines()[0])
print('REF', r['value'], r['e2e'])
PY
