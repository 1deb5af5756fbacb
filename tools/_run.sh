mkdir -p gpurun_out
( time timeout 200 python -m pytest tests/test_gpu_awq_surface.py -x -q -rs ) > gpurun_out/pytest_gpu_awq.log 2>&1; echo "pytest awq rc=$?"; tail -25 gpurun_out/pytest_gpu_awq.log | cut -c1-300
( time timeout 300 python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
( timeout 200 python tools/bench_model.py --model llama-2-7b --batch 1 8 --generate --out gpurun_out/model_llama2_7b_generate.json ) > gpurun_out/model_7b_generate.log 2>&1; echo "generate rc=$?"; grep '^{' gpurun_out/model_7b_generate.log | cut -c1-400
( timeout 200 python bench.py --steps 10 --warmup 3 ) > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_r1d.json; tail -2 gpurun_out/bench_r1d.err
