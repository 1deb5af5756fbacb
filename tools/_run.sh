mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_model.py --model llama-2-7b --impl reference --batch 1 8 32 64 --out gpurun_out/model_llama2_7b_refkernel_fused.json > gpurun_out/model_7b_ref_fused.log 2>&1; echo "rc=$?"; grep '^{' gpurun_out/model_7b_ref_fused.log | cut -c1-200
timeout 600 python tools/bench_model.py --model mistral-7b --batch 1 8 32 64 --out gpurun_out/model_mistral_7b_fused.json > gpurun_out/model_mistral_fused.log 2>&1; echo "rc=$?"; grep '^{' gpurun_out/model_mistral_fused.log | cut -c1-200
timeout 600 python tools/bench_model.py --model llama-2-70b --batch 1 8 --out gpurun_out/model_llama2_70b_fused.json > gpurun_out/model_70b_fused.log 2>&1; echo "rc=$?"; grep '^{' gpurun_out/model_70b_fused.log | cut -c1-200
