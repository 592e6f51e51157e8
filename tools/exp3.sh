#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/tests.log 2>&1
tail -5 gpurun_out/tests.log
for park in 0 8 16 24 32; do echo "== chunk-fused park $park"; PTB_PARK=$park python tools/quick_bench.py --pipeline 3 --batch 8 | grep Msegments; done
for T in 512 1024; do for park in 24 32; do echo "== threads $T park $park"; PTB_LIB=$PWD/szakdolgozat_pathtracer_b200/libptb_t$T.so PTB_PARK=$park python tools/quick_bench.py --pipeline 3 --batch 8 | grep Msegments; done; done
echo "== batch 1"; python tools/quick_bench.py --pipeline 3 --batch 1 --launches 16 | grep Msegments
echo "== warp-fused"; python tools/quick_bench.py --pipeline 4 --batch 8 | grep Msegments
python tools/bvh_probe.py c5
