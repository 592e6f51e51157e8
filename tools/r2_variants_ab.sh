#!/bin/bash
# A/B of build variants (tools/build_variants.sh): usage r2_variants_ab.sh name1 name2 ...
mkdir -p gpurun_out
L=gpurun_out/r2_variants_ab.log; : > $L
for v in "$@"; do
  export PTB_LIB=$PWD/build/var_$v/libptb.so
  for cfg in "--config c2" "--config c2 --camera monkey_close" "--config c5" "--config c3 --camera suitcase_close" "--config c4"; do
    for ar in ${ARITHS:-1}; do
      echo -n "$v arith=$ar $cfg: " >> $L
      timeout 600 python tools/quick_bench.py $cfg --batch 8 --arith $ar --launches 4 2>&1 | grep "ms/launch" | sed 's/.*depth 8: //' | cut -c1-60 >> $L
    done
  done
  echo -n "$v arith=0 --config c2: " >> $L
  timeout 600 python tools/quick_bench.py --config c2 --batch 8 --arith 0 --launches 4 2>&1 | grep "ms/launch" | sed 's/.*depth 8: //' | cut -c1-60 >> $L
done
cat $L
