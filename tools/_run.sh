# Round-end style check on a GPU box: smoke, the GPU suite, both bench arms, model tokens/s (edit as needed).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
( time timeout 300 python -m pytest tests -m gpu -x -q -rs ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
( timeout 300 python bench.py --impl reference --steps 10 --warmup 3 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
( timeout 300 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench.json
( timeout 200 python tools/bench_model.py --model llama-2-7b --batch 1 8 32 64 --generate --out gpurun_out/model_llama2_7b.json ) > gpurun_out/model_7b.log 2>&1; echo "model rc=$?"; grep '^{' gpurun_out/model_7b.log | cut -c1-260
