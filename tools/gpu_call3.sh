#!/bin/bash
# round 2, call 3: new tests on the default build + A/B variants of the fused kernel (fast and exact arithmetic)
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_fast_mode.py tests/test_gpu_configs.py -m gpu -q -s -k "not c4" ) > gpurun_out/r2_call3_tests.log 2>&1
grep -n "fast-mode gate\|passed\|failed\|^FAILED" gpurun_out/r2_call3_tests.log | cut -c1-260
rm -f gpurun_out/r2_call3_bench.log
run() { # name lib
  for cfg in "--config c2 --arith 0" "--config c2 --arith 1" "--config c2 --camera monkey_close --arith 1" "--config c5 --arith 1"; do
    echo -n "$1: " >> gpurun_out/r2_call3_bench.log
    PTB_LIB=$2 python tools/quick_bench.py $cfg --batch 8 --launches 3 2>&1 | grep "Msegments" | cut -c1-90 >> gpurun_out/r2_call3_bench.log
  done
}
run default szakdolgozat_pathtracer_b200/libptb.so
for v in excam minb10 minb12 t128 shdyn q4 q16; do run $v build/var_$v/libptb.so; done
cat gpurun_out/r2_call3_bench.log
