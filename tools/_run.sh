mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_model.py --model llama-2-7b --batch 1 8 32 64 --out gpurun_out/model_llama2_7b_fused_pdl.json > gpurun_out/model_7b_fused_pdl.log 2>&1; echo "rc=$?"; grep '^{' gpurun_out/model_7b_fused_pdl.log | cut -c1-200
QB200_NO_PDL=1 timeout 600 python tools/bench_model.py --model llama-2-7b --batch 1 64 > gpurun_out/model_7b_fused_nopdl.log 2>&1; echo "rc=$?"; grep '^{' gpurun_out/model_7b_fused_nopdl.log | cut -c1-200
