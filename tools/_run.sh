mkdir -p gpurun_out
export QB200_LIB=$PWD/quick_b200/libquick_b200_dev.so
VARS=0,1 timeout 300 python tools/check_fast.py > gpurun_out/check13.log 2>&1; rc=$?; echo "check rc=$rc"; tail -1 gpurun_out/check13.log; grep '"ok": false' gpurun_out/check13.log | head -5; grep -i "error\|Traceback" gpurun_out/check13.log | head -5
[ $rc -ne 0 ] && exit 1
MS=1,16,64,128,256,512 VARS=0,1 SPLITS=1,2,4 OUT=tune_v13.json timeout 400 python tools/tune.py > gpurun_out/tune_v13.log 2>&1; echo "tune rc=$?"
