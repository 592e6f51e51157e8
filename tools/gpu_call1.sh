#!/bin/bash
# round 2, call 1: GPU test suite + baseline numbers of the round-1 kernels on this round's box
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_call1_tests.log 2>&1
tail -5 gpurun_out/r2_call1_tests.log
for cfg in "--config c2" "--config c2 --camera monkey_close" "--config c5" "--config c3 --camera suitcase_close" "--config c4"; do
  echo "== $cfg" >> gpurun_out/r2_call1_bench.log
  python tools/quick_bench.py $cfg --batch 8 --launches 3 2>&1 | grep -v "^build\|mean frame" >> gpurun_out/r2_call1_bench.log
done
cat gpurun_out/r2_call1_bench.log | cut -c1-200
