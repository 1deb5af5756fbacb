mkdir -p gpurun_out
export QB200_LIB=$PWD/quick_b200/libquick_b200_trace.so
for cfg in "1 4096 4096" "1 4096 4096 16 8" "16 4096 4096 32 4" "256 4096 4096" "256 4096 4096 128 2" "512 4096 4096"; do timeout 120 python tools/timeline.py $cfg; done > gpurun_out/timeline_v4.log 2>&1
QB200_NO_PDL=1 timeout 120 python tools/timeline.py 1 4096 4096 >> gpurun_out/timeline_v4.log 2>&1
for cfg in "16 4 1 4096 4096 0" "64 1 256 4096 4096 0" "128 1 512 4096 4096 0"; do timeout 120 python tools/trace.py $cfg; done > gpurun_out/trace_v4b.log 2>&1
cat gpurun_out/timeline_v4.log
