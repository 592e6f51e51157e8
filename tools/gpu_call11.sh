#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r2_call11_bench.log
run() { # name lib
  for cfg in "--config c2 --arith 0" "--config c2 --arith 1" "--config c5 --arith 1" "--config c2 --camera monkey_close --arith 1" "--config c2 --arith 1 --batch 1 --launches 8" "--config c2 --arith 0 --batch 1 --launches 8"; do
    echo -n "$1: " >> gpurun_out/r2_call11_bench.log
    extra="--batch 8 --launches 3"; [[ "$cfg" == *"--batch"* ]] && extra=""
    PTB_LIB=$2 python tools/quick_bench.py $cfg $extra 2>&1 | grep "Msegments" | cut -c1-90 >> gpurun_out/r2_call11_bench.log
  done
}
run default szakdolgozat_pathtracer_b200/libptb.so
run frg build/var_frg/libptb.so
cat gpurun_out/r2_call11_bench.log
PTB_LIB=build/var_frg/libptb.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_fast_mode.py -m gpu -q -k "chunk-fused or fast" 2>&1 | tail -3
