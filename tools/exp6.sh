#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/tests.log 2>&1
tail -5 gpurun_out/tests.log
for pl in 3 4; do for b in 8 1; do echo "== pipeline $pl batch $b"; python tools/quick_bench.py --pipeline $pl --batch $b --launches $((32/b)) | grep Msegments; done; done
for pl in 3 4; do echo "== c2_close pipeline $pl"; python tools/quick_bench.py --pipeline $pl --batch 8 --camera monkey_close | grep Msegments; done
for pl in 3 4; do echo "== c3 4K pipeline $pl"; python tools/quick_bench.py --config c3 --width 3840 --height 2160 --pipeline $pl --batch 4 | grep Msegments; done
for pl in 3 4; do echo "== c5 pipeline $pl"; python tools/quick_bench.py --config c5 --pipeline $pl --batch 8 | grep Msegments; done
for pl in 3 4; do echo "== c4 pipeline $pl"; python tools/quick_bench.py --config c4 --pipeline $pl --batch 8 | grep Msegments; done
ncu --set full --clock-control none --import-source on -k regex:k_pool_fused -s 1 -c 1 -o gpurun_out/pool python tools/quick_bench.py --pipeline 4 --batch 8 > gpurun_out/ncu_pool.log 2>&1
