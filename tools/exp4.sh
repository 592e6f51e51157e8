#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/tests.log 2>&1
tail -5 gpurun_out/tests.log
for mb in 8 6; do for park in 0 24; do echo "== chunk-fused minb $mb park $park"; PTB_MINB=$mb PTB_PARK=$park python tools/quick_bench.py --pipeline 3 --batch 8 | grep Msegments; done; done
for L in spt16 t512; do for park in 0 24; do echo "== $L park $park"; PTB_LIB=$PWD/szakdolgozat_pathtracer_b200/libptb_$L.so PTB_PARK=$park python tools/quick_bench.py --pipeline 3 --batch 8 | grep Msegments; done; done
echo "== batch 1"; python tools/quick_bench.py --pipeline 3 --batch 1 --launches 16 | grep Msegments
for c in c5 c4; do
echo "== $c aniso"; PTB_MORTON_ANISO=1 python tools/quick_bench.py --config $c --pipeline 3 --batch 4 --count 1 | grep -E "build|Msegments|nodes/seg"
echo "== $c cubic"; python tools/quick_bench.py --config $c --pipeline 3 --batch 4 --count 1 | grep -E "build|Msegments|nodes/seg"
echo "== $c cubic norefine"; python tools/quick_bench.py --config $c --pipeline 3 --batch 4 --count 1 --refine 0 | grep -E "build|Msegments|nodes/seg"
done
echo "== c2 cubic count"; python tools/quick_bench.py --config c2 --pipeline 3 --batch 8 --count 1 | grep -E "build|Msegments|nodes/seg"
echo "== c2 aniso count"; PTB_MORTON_ANISO=1 python tools/quick_bench.py --config c2 --pipeline 3 --batch 8 --count 1 | grep -E "build|Msegments|nodes/seg"
echo "== c3 cubic"; python tools/quick_bench.py --config c3 --pipeline 3 --batch 8 --count 1 | grep -E "build|Msegments|nodes/seg"
echo "== c3 aniso"; PTB_MORTON_ANISO=1 python tools/quick_bench.py --config c3 --pipeline 3 --batch 8 --count 1 | grep -E "build|Msegments|nodes/seg"
