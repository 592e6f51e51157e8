#!/bin/bash
# experiment driver (development aid): run inside gpurun
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/tests.log 2>&1
tail -4 gpurun_out/tests.log
echo "== chunk-fused baseline"; python tools/quick_bench.py --pipeline 3 --batch 8 | grep Msegments
for spl in 8 16 32; do echo "== warp-fused spl $spl minb 8"; PTB_WF_SPL=$spl PTB_WF_MINB=8 python tools/quick_bench.py --pipeline 4 --batch 8 | grep Msegments; done
for mb in 7 6; do echo "== warp-fused spl 8 minb $mb"; PTB_WF_SPL=8 PTB_WF_MINB=$mb python tools/quick_bench.py --pipeline 4 --batch 8 | grep Msegments; done
echo "== warp-fused spl 16 minb 6"; PTB_WF_SPL=16 PTB_WF_MINB=6 python tools/quick_bench.py --pipeline 4 --batch 8 | grep Msegments
echo "== warp-fused batch 1"; python tools/quick_bench.py --pipeline 4 --batch 1 --launches 16 | grep Msegments
echo "== chunk-fused batch 1"; python tools/quick_bench.py --pipeline 3 --batch 1 --launches 16 | grep Msegments
for c in c4 c5; do for mbits in 30 63; do echo "== $c morton $mbits"; python tools/quick_bench.py --config $c --pipeline 4 --batch 4 --count 1 --morton $mbits | grep -E "build|Msegments|nodes/seg"; done; done
