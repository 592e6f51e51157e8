#!/bin/bash
# ncu --set full capture of the fused kernel on every BASELINE configuration, at the geometry of ONE pool batch of the
# bench.py workload, both arithmetic modes for C2; writes the .ncu-rep / raw csv to gpurun_out/ and the summary entries bench.py
# reads to profiles/r2_ncu_summary.json (tools/ncu_to_json.py).  One GPU; every workload has run without ncu before.
mkdir -p gpurun_out
cap() { # key config-args arith
  name=$(echo "$1" | tr ':' '_')
  python tools/quick_bench.py $2 --arith $3 --launches 1 > gpurun_out/r2_prof_$name.plain.log 2>&1 || { echo "plain run failed: $1"; return; }
  ncu --set full --clock-control none --import-source on -k regex:k_chunk_fused -s 1 -c 1 -o gpurun_out/r2_$name -f python tools/quick_bench.py $2 --arith $3 --launches 1 > gpurun_out/r2_prof_$name.log 2>&1
  ncu -i gpurun_out/r2_$name.ncu-rep --page raw --csv > gpurun_out/r2_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r2_$name.ncu-rep --page source --csv --print-source cuda,sass > /tmp/r2_${name}_src.csv 2>/dev/null
  { echo "ncu --set full --clock-control none --import-source on -k regex:k_chunk_fused -s 1 -c 1, tools/quick_bench.py $2 --arith $3"; python tools/ncu_raw_summary.py gpurun_out/r2_${name}_raw.csv;
    echo; echo "per source line (tools/ncu_source_summary.py):"; python tools/ncu_source_summary.py /tmp/r2_${name}_src.csv 30; } > gpurun_out/r2_${name}_ncu_full.txt 2>&1
  rm -f gpurun_out/r2_$name.ncu-rep /tmp/r2_${name}_src.csv   # the report itself is ~10 MB: gpurun_out/ is capped at 64 MiB
  seg=$(grep -o "last-launch segments [0-9]*" gpurun_out/r2_prof_$name.log | grep -o "[0-9]*$")
  python tools/ncu_to_json.py gpurun_out/r2_${name}_raw.csv "$1" "$seg" "ncu --set full --clock-control none, tools/quick_bench.py $2 --arith $3 (one launch = one pool batch of the bench.py workload); raw export: profiles/r2_${name}_raw.csv" | cut -c1-250
}
cap "c2:fast:pipeline3"       "--config c2 --batch 8" 1
cap "c2:exact:pipeline3"      "--config c2 --batch 8" 0
cap "c2_close:fast:pipeline3" "--config c2 --camera monkey_close --batch 8" 1
cap "c3:fast:pipeline3"       "--config c3 --width 3840 --height 2160 --batch 2" 1
cap "c4:fast:pipeline3"       "--config c4 --batch 8" 1
cap "c5:fast:pipeline3"       "--config c5 --width 3840 --height 2160 --spp 64 --batch 2" 1
