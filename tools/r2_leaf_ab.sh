#!/bin/bash
# A/B: leaf size of the 2-/4-wide tree; 8-wide tree after the PRMT byte conversion
mkdir -p gpurun_out
L=gpurun_out/r2_leaf_ab.log; : > $L
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast_mode.py -x -q -m gpu -k "bvh or wide" ) > gpurun_out/r2_bvh8_tests.log 2>&1; tail -3 gpurun_out/r2_bvh8_tests.log
for cfg in "--config c2" "--config c2 --camera monkey_close" "--config c4"; do
  for leaf in 1 2 3 4 6 8; do
    echo "leaf=$leaf $cfg:" >> $L
    timeout 600 python tools/quick_bench.py $cfg --batch 8 --arith 1 --leaf $leaf --launches 3 2>&1 | grep -v "^mean frame" | cut -c1-200 >> $L
    timeout 300 python tools/quick_bench.py $cfg --batch 1 --arith 1 --leaf $leaf --launches 1 --count 1 2>&1 | grep "nodes/seg" >> $L
  done
  echo "width=8 $cfg:" >> $L
  timeout 600 python tools/quick_bench.py $cfg --batch 8 --arith 1 --bvh-width 8 --launches 3 2>&1 | grep -v "^mean frame" | cut -c1-200 >> $L
done
cat $L
