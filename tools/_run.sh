mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
( time timeout 400 python bench.py --impl reference --steps 10 --warmup 3 ) > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; echo "bench ref rc=$?"
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:umma -c 900 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
for m in 1 256 512; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma -s 10 -c 2 -f -o gpurun_out/full_M$m python tools/ncu_one.py $m > gpurun_out/ncu_full_M$m.log 2>&1; echo "ncu full M=$m rc=$?"; done
for m in 1 256; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma -s 10 -c 2 -f -o gpurun_out/full_indep_M$m python tools/ncu_one.py $m indep > gpurun_out/ncu_full_indep_M$m.log 2>&1; echo "ncu full indep M=$m rc=$?"; done
timeout 600 python tools/bench_model.py --model llama-2-7b --out gpurun_out/model_llama2_7b.json > gpurun_out/model_llama2_7b.log 2>&1; echo "model rc=$?"; tail -4 gpurun_out/model_llama2_7b.log | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[0])
print(d['value'], d['e2e']['value'], d['roofline'], d['clocks'])
for r in d['sweep']: print(r)
print(d['independent']['value'])
for r in d['independent']['sweep']: print(r)
r=json.loads(open('gpurun_out/bench_ref_final.json').read().strip().splitlines()[0])
print('REF', r['value'], r['e2e'], [ (x['M'], x['us']) for x in r.get('sweep', [])])
PY
