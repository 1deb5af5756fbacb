mkdir -p gpurun_out
( time timeout 70 python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-400
( timeout 40 python tools/bench_model.py --model llama-2-7b --batch 1 8 64 --gen 64 --out gpurun_out/model_llama2_7b_attn_decode.json ) > gpurun_out/model_7b_attn.log 2>&1; echo "model attn rc=$?"; grep '^{' gpurun_out/model_7b_attn.log | cut -c1-210
