#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_multi_gpu.py > gpurun_out/r2_multi_gpu_check_n2.txt 2>&1; echo "check rc=$?"; grep "^world" gpurun_out/r2_multi_gpu_check_n2.txt
timeout 600 python tools/c5_convergence.py --gpus 1,2 --rows 64 > gpurun_out/r2_c5_conv_n2.log 2>&1; echo "c5 rc=$?"; tail -6 gpurun_out/r2_c5_conv_n2.log | cut -c1-500
( time python -m pytest tests -m gpu -q -k "not (queues or chunk-stages or pool-fused or c4)" ) 2>&1 | tail -6
for cfg in "--config c2 --arith 0" "--config c2 --arith 1"; do python tools/quick_bench.py $cfg --batch 8 --launches 3 2>&1 | grep "Msegments" | cut -c1-100; done
