#!/bin/bash
# round 2, call 2: full GPU suite, exact vs fast arithmetic timings, ncu capture of the fast fused kernel
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -s -k "not (queues or chunk-stages or pool-fused)" ) > gpurun_out/r2_call2_tests.log 2>&1
tail -15 gpurun_out/r2_call2_tests.log
rm -f gpurun_out/r2_call2_bench.log
for cfg in "--config c2" "--config c2 --camera monkey_close" "--config c5" "--config c3 --camera suitcase_close" "--config c4"; do
  for ar in 0 1; do
    python tools/quick_bench.py $cfg --batch 8 --launches 3 --arith $ar 2>&1 | grep "Msegments" >> gpurun_out/r2_call2_bench.log
  done
done
cut -c1-150 gpurun_out/r2_call2_bench.log
ncu --set full --clock-control none --import-source on -k regex:k_chunk_fused -s 1 -c 1 -o gpurun_out/r2_fast1 python tools/quick_bench.py --pipeline 3 --batch 8 --arith 1 > gpurun_out/r2_ncu_fast1.log 2>&1
ncu -i gpurun_out/r2_fast1.ncu-rep --page raw --csv > gpurun_out/r2_fast1_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_fast1.ncu-rep --page source --csv > gpurun_out/r2_fast1_src.csv 2>/dev/null
ls -la gpurun_out/r2_fast1*
