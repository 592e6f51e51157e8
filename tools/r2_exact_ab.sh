#!/bin/bash
# A/B of build variants, exact arithmetic only: usage r2_exact_ab.sh name1 name2 ...
mkdir -p gpurun_out
L=gpurun_out/r2_exact_ab.log; : > $L
for v in "$@"; do
  export PTB_LIB=$PWD/build/var_$v/libptb.so
  for cfg in "--config c2" "--config c2 --camera monkey_close" "--config c5" "--config c3 --camera suitcase_close" "--config c4"; do
    echo -n "$v arith=0 $cfg: " >> $L
    timeout 600 python tools/quick_bench.py $cfg --batch 8 --arith 0 --launches 4 2>&1 | grep "ms/launch" | sed 's/.*depth 8: //' | cut -c1-60 >> $L
  done
done
cat $L
