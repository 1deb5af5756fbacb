mkdir -p gpurun_out
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final2.json').read().strip().splitlines()[0])
print(d['value'], d['e2e']['value'], d['independent']['value'], d['llama2_7b_tokens_per_s'])
PY
