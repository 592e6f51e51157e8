#!/bin/bash
# one ncu --set full capture of the final C2 fast kernel (same command as tools/profile_configs.sh, without the plain pre-run:
# the same program ran without ncu in the validation call before)
mkdir -p gpurun_out
name=c2_fast_pipeline3; args="--config c2 --batch 8"
ncu --set full --clock-control none --import-source on -k regex:k_chunk_fused -s 1 -c 1 -o gpurun_out/r2_$name -f python tools/quick_bench.py $args --arith 1 --launches 1 > gpurun_out/r2_prof_$name.log 2>&1
ncu -i gpurun_out/r2_$name.ncu-rep --page raw --csv > gpurun_out/r2_${name}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_$name.ncu-rep --page source --csv --print-source cuda,sass > /tmp/r2_${name}_src.csv 2>/dev/null
{ echo "ncu --set full --clock-control none --import-source on -k regex:k_chunk_fused -s 1 -c 1, tools/quick_bench.py $args --arith 1"; python tools/ncu_raw_summary.py gpurun_out/r2_${name}_raw.csv;
  echo; echo "per source line (tools/ncu_source_summary.py):"; python tools/ncu_source_summary.py /tmp/r2_${name}_src.csv 30; } > gpurun_out/r2_${name}_ncu_full.txt 2>&1
rm -f gpurun_out/r2_$name.ncu-rep
head -2 gpurun_out/r2_${name}_ncu_full.txt | cut -c1-400
