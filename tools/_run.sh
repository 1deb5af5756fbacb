mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 tools/tp_check.py ) > gpurun_out/tp_check_mc.log 2>&1; rc=$?; echo "tp_check rc=$rc"; grep "TP_CHECK\|False\|multicast\|rror" gpurun_out/tp_check_mc.log | head
[ $rc -ne 0 ] && exit 1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 tools/bench_model.py --model llama-2-70b --batch 1 32 --out gpurun_out/model_llama2_70b_tp2_mc.json ) > gpurun_out/model_70b_tp2_mc.log 2>&1; echo "70b mc rc=$?"; grep '^{' gpurun_out/model_70b_tp2_mc.log | cut -c1-230
( time QB200_TP_MULTICAST=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 tools/bench_model.py --model llama-2-70b --batch 1 32 ) > gpurun_out/model_70b_tp2_nomc.log 2>&1; echo "70b nomc rc=$?"; grep '^{' gpurun_out/model_70b_tp2_nomc.log | cut -c1-230
