#!/bin/bash
mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/tests.log 2>&1
tail -5 gpurun_out/tests.log
for hf in 0 0.0625; do for pl in 3 4; do echo "== huge $hf pipeline $pl"; PTB_HUGE_FRAC=$hf python tools/quick_bench.py --pipeline $pl --batch 8 | grep -E "build|Msegments"; done; done
for hf in 0 0.0625; do echo "== huge $hf count"; PTB_HUGE_FRAC=$hf python tools/quick_bench.py --pipeline 3 --batch 8 --count 1 | grep -E "nodes/seg"; done
for c in "c2 --camera monkey_close" "c3 --width 3840 --height 2160" c4 c5; do for hf in 0 0.0625; do echo "== $c huge $hf"; PTB_HUGE_FRAC=$hf python tools/quick_bench.py --config $c --pipeline 3 --batch 4 --count 1 | grep -E "build|nodes/seg"; PTB_HUGE_FRAC=$hf python tools/quick_bench.py --config $c --pipeline 3 --batch 4 | grep -E "Msegments"; done; done
