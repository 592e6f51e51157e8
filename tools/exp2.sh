#!/bin/bash
mkdir -p gpurun_out
for lib in libptb libptb_ni1 libptb_ni2; do for pl in 3 4; do echo "== $lib pipeline $pl"; PTB_LIB=$PWD/szakdolgozat_pathtracer_b200/$lib.so python tools/quick_bench.py --pipeline $pl --batch 8 | grep Msegments; done; done
echo "== c5 verbose"; PTB_VERBOSE=1 python tools/quick_bench.py --config c5 --pipeline 3 --batch 4 --count 1 2>&1 | grep -E "ptb_accel|build|Msegments|nodes/seg"
echo "== c5 verbose morton30"; PTB_VERBOSE=1 python tools/quick_bench.py --config c5 --pipeline 3 --batch 4 --count 1 --morton 30 2>&1 | grep -E "ptb_accel|build|Msegments|nodes/seg"
echo "== c4 verbose"; PTB_VERBOSE=1 python tools/quick_bench.py --config c4 --pipeline 3 --batch 4 --count 1 2>&1 | grep -E "ptb_accel|build|Msegments|nodes/seg"
ncu --set full --clock-control none --import-source on -k regex:k_warp_fused -s 1 -c 1 -o gpurun_out/warp python tools/quick_bench.py --pipeline 4 --batch 8 > gpurun_out/ncu_warp.log 2>&1
