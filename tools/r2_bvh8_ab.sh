#!/bin/bash
# A/B of the BVH widths (2 / 4 / 8-wide quantised) on the five workloads, fast arithmetic; first the tests that cover the new tree
mkdir -p gpurun_out
L=gpurun_out/r2_bvh8_ab.log; : > $L
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast_mode.py -x -q -m gpu -k "bvh or wide" ) > gpurun_out/r2_bvh8_tests.log 2>&1; tail -5 gpurun_out/r2_bvh8_tests.log
for cfg in "--config c2" "--config c2 --camera monkey_close" "--config c5" "--config c3 --camera suitcase_close" "--config c4"; do
  for w in ${WIDTHS:-2 4 8}; do
    echo "width=$w $cfg:" >> $L
    timeout 600 python tools/quick_bench.py $cfg --batch 8 --arith 1 --bvh-width $w --launches 3 2>&1 | grep -v "^mean frame" | cut -c1-230 >> $L
    timeout 300 python tools/quick_bench.py $cfg --batch 1 --arith 1 --bvh-width $w --launches 1 --count 1 2>&1 | grep "nodes/seg" >> $L
  done
done
cat $L
