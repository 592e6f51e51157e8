#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s -k "not (queues or chunk-stages)" ) > gpurun_out/r2_call7_tests.log 2>&1
grep -n "passed\|failed\|^FAILED\|^ERROR" gpurun_out/r2_call7_tests.log | cut -c1-330
rm -f gpurun_out/r2_call7_bench.log
run() { # name lib
  for cfg in "--config c2 --arith 0" "--config c2 --arith 1" "--config c2 --camera monkey_close --arith 1" "--config c5 --arith 1" "--config c3 --camera suitcase_close --arith 1" "--config c4 --arith 1"; do
    echo -n "$1: " >> gpurun_out/r2_call7_bench.log
    PTB_LIB=$2 python tools/quick_bench.py $cfg --batch 8 --launches 3 2>&1 | grep "Msegments" | cut -c1-90 >> gpurun_out/r2_call7_bench.log
  done
}
run default szakdolgozat_pathtracer_b200/libptb.so
for v in notop q16; do run $v build/var_$v/libptb.so; done
cat gpurun_out/r2_call7_bench.log
