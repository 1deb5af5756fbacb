mkdir -p gpurun_out
timeout 600 python tools/bench_model.py --model llama-2-7b --batch 1 8 64 --out gpurun_out/model_llama2_7b_planner.json > gpurun_out/model_7b_planner.log 2>&1; echo "rc=$?"; grep '^{' gpurun_out/model_7b_planner.log | cut -c1-200
export QB200_LIB=$PWD/quick_b200/libquick_b200_dev.so
K=4096 N=12288 MS=1,16,64 VARS=0 SPLITS=1,2,4 OUT=tune_qkv.json timeout 300 python tools/tune.py > gpurun_out/tune_qkv.log 2>&1
K=4096 N=22016 MS=1,16,64 VARS=0 SPLITS=1,2,4,8 OUT=tune_gu.json timeout 300 python tools/tune.py > gpurun_out/tune_gu.log 2>&1
K=11008 N=4096 MS=1,16,64 VARS=0 SPLITS=2,4,8 OUT=tune_down.json timeout 300 python tools/tune.py > gpurun_out/tune_down.log 2>&1
for f in qkv gu down; do echo $f; cut -c1-150 gpurun_out/tune_$f.log | grep -v '"tok": 256\|"tok": 128'; done
