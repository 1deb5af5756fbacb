mkdir -p gpurun_out
( time timeout 100 python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -12 gpurun_out/pytest_gpu.log | cut -c1-400
( timeout 60 python tools/bench_model.py --model llama-2-7b --batch 1 8 64 --gen 64 --out gpurun_out/model_llama2_7b_attn_decode.json ) > gpurun_out/model_7b_attn.log 2>&1; echo "model attn rc=$?"; grep '^{' gpurun_out/model_7b_attn.log | cut -c1-210
( QB200_ATTN_DECODE=0 timeout 40 python tools/bench_model.py --model llama-2-7b --batch 1 64 --gen 64 ) > gpurun_out/model_7b_sdpa.log 2>&1; echo "model sdpa rc=$?"; grep '^{' gpurun_out/model_7b_sdpa.log | cut -c1-210
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
