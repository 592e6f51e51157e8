#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -3 gpurun_out/r2_bench_a.err; cut -c1-600 gpurun_out/r2_bench_a.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_a_ref.json 2>> gpurun_out/r2_bench_a.err; cut -c1-300 gpurun_out/r2_bench_a_ref.json
for ar in 1 0; do
ncu --set full --clock-control none --import-source on -k regex:k_chunk_fused -s 1 -c 1 -o gpurun_out/r2_c2_arith$ar python tools/quick_bench.py --pipeline 3 --batch 8 --arith $ar > gpurun_out/r2_ncu_c2_arith$ar.log 2>&1
ncu -i gpurun_out/r2_c2_arith$ar.ncu-rep --page raw --csv > gpurun_out/r2_c2_arith${ar}_raw.csv 2>/dev/null
done
python -m pytest tests/test_gpu_multi.py tests/test_cli.py -m gpu -q 2>&1 | tail -5
