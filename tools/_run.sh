mkdir -p gpurun_out
( time timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -k "gemv" ) > gpurun_out/pytest_gemv.log 2>&1; echo "pytest gemv rc=$?"; tail -30 gpurun_out/pytest_gemv.log | cut -c1-300
( time timeout 150 python tools/gemv_check.py --out gpurun_out/gemv_check.json ) > gpurun_out/gemv_check.log 2>&1; echo "gemv_check rc=$?"; tail -20 gpurun_out/gemv_check.log | cut -c1-300
for mm in 0 1 2 4; do ( QB200_GEMV_MAX_M=$mm timeout 100 python tools/bench_model.py --model llama-2-7b --batch 1 2 4 --gen 64 ) > gpurun_out/model_7b_gemv$mm.log 2>&1; echo "model gemv_max_m=$mm rc=$?"; grep '^{' gpurun_out/model_7b_gemv$mm.log | cut -c1-200; done
