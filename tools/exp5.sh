#!/bin/bash
mkdir -p gpurun_out
for merge in 0 1; do for park in 0 24; do echo "== chunk-fused merge $merge park $park"; PTB_MERGE=$merge PTB_PARK=$park python tools/quick_bench.py --pipeline 3 --batch 8 | grep Msegments; done; done
for L in spt16 t512; do for park in 0 24; do echo "== $L park $park"; PTB_LIB=$PWD/szakdolgozat_pathtracer_b200/libptb_$L.so PTB_PARK=$park python tools/quick_bench.py --pipeline 3 --batch 8 | grep Msegments; done; done
echo "== batch 1"; python tools/quick_bench.py --pipeline 3 --batch 1 --launches 16 | grep Msegments
echo "== warp"; python tools/quick_bench.py --pipeline 4 --batch 8 | grep Msegments
ncu --set full --clock-control none --import-source on -k regex:k_chunk_fused -s 1 -c 1 -o gpurun_out/fused2 python tools/quick_bench.py --pipeline 3 --batch 8 > gpurun_out/ncu_fused2.log 2>&1
