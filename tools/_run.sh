mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus=$N"
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/tp_check.py ) > gpurun_out/tp_check_n$N.log 2>&1; echo "tp_check rc=$?"; grep "TP_CHECK\|False" gpurun_out/tp_check_n$N.log | head
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 10 --warmup 3 ) > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines() if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['e2e']['value'], d['independent']['value'], d['llama2_7b_tokens_per_s'])
PY
