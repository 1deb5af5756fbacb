export QB200_LIB=$PWD/quick_b200/libquick_b200_dev.so
VARS=0,1 timeout 300 python tools/check_fast.py > gpurun_out/check3.log 2>&1; echo "check rc=$?"; tail -1 gpurun_out/check3.log; grep '"ok": false' gpurun_out/check3.log | head
MS=1,16,32,64,128,256,512 VARS=0,1 SPLITS=1,2,4,8 OUT=tune_v4c.json timeout 600 python tools/tune.py > gpurun_out/tune_v4c.log 2>&1; echo "tune rc=$?"
for cfg in "16 4 1 4096 4096 3000" "64 4 64 4096 4096 2000" "128 2 256 4096 4096 2000" "0 0 16 4096 11008 2000"; do timeout 120 python tools/stress.py $cfg; done > gpurun_out/stress3.log 2>&1; tail -4 gpurun_out/stress3.log
export QB200_LIB=$PWD/quick_b200/libquick_b200_trace.so
for cfg in "16 4 1 4096 4096 0" "128 2 256 4096 4096 0" "256 4 256 4096 4096 0"; do timeout 120 python tools/trace.py $cfg; done > gpurun_out/trace_v4c.log 2>&1
