#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast_mode.py tests/test_gpu_configs.py -x -q -m gpu -k "bvh or wide or c4" ) 2>&1 | tail -3
L=gpurun_out/r2_c4_greedy.log; : > $L
for ar in 1 0; do
  echo "c4 arith=$ar:" >> $L
  timeout 600 python tools/quick_bench.py --config c4 --batch 8 --arith $ar --launches 4 2>&1 | grep -v "^mean frame" | cut -c1-220 >> $L
done
timeout 300 python tools/quick_bench.py --config c4 --batch 1 --arith 1 --launches 1 --count 1 2>&1 | grep "nodes/seg" >> $L
for cfg in "--config c5" "--config c2 --camera monkey_close"; do
  echo "width 4 $cfg:" >> $L
  timeout 600 python tools/quick_bench.py $cfg --batch 8 --arith 1 --bvh-width 4 --launches 4 2>&1 | grep "ms/launch" | cut -c1-200 >> $L
  timeout 300 python tools/quick_bench.py $cfg --batch 1 --arith 1 --bvh-width 4 --launches 1 --count 1 2>&1 | grep "nodes/seg" >> $L
done
cat $L
