mkdir -p gpurun_out
( time timeout 200 python -m pytest tests/test_gpu_awq_surface.py -x -q -rs ) > gpurun_out/pytest_gpu_awq.log 2>&1; echo "pytest awq rc=$?"; tail -25 gpurun_out/pytest_gpu_awq.log | cut -c1-300
( time timeout 200 python tools/e2e_checkpoint.py --out gpurun_out/e2e_checkpoint_llama.json ) > gpurun_out/e2e_llama.log 2>&1; echo "e2e llama rc=$?"; tail -3 gpurun_out/e2e_llama.log | cut -c1-1500
( time timeout 120 python examples/benchmark.py --random_init llama-2-7b --batch_size 1 64 --out gpurun_out/example_benchmark_7b.json ) > gpurun_out/example_benchmark_7b.log 2>&1; echo "example rc=$?"; tail -5 gpurun_out/example_benchmark_7b.log | cut -c1-400
