mkdir -p gpurun_out
python bench.py > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err
python bench.py --pipeline 4 --no-cpu-baseline --steps 10 > gpurun_out/bench_v6_pool.json 2>> gpurun_out/bench_v6.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_v6_ref.json 2>> gpurun_out/bench_v6.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r1_v6_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chunk_fused -s 1 -c 1 -o gpurun_out/fused6 python tools/quick_bench.py --pipeline 3 --batch 8 > gpurun_out/ncu_fused6.log 2>&1
python tools/run_configs.py > gpurun_out/run_configs.log 2>&1
cut -c1-300 gpurun_out/bench_v6.json; tail -3 gpurun_out/run_configs.log | cut -c1-300
