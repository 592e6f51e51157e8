#!/bin/bash
# after the greedy 4-wide collapse: the C4 capture and bench line again (same commands as tools/profile_configs.sh / gpu_single_final.sh)
mkdir -p gpurun_out
name=c4_fast_pipeline3; args="--config c4 --batch 8"
python tools/quick_bench.py $args --arith 1 --launches 1 > gpurun_out/r2_prof_$name.plain.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chunk_fused -s 1 -c 1 -o gpurun_out/r2_$name -f python tools/quick_bench.py $args --arith 1 --launches 1 > gpurun_out/r2_prof_$name.log 2>&1
ncu -i gpurun_out/r2_$name.ncu-rep --page raw --csv > gpurun_out/r2_${name}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_$name.ncu-rep --page source --csv --print-source cuda,sass > /tmp/r2_${name}_src.csv 2>/dev/null
{ echo "ncu --set full --clock-control none --import-source on -k regex:k_chunk_fused -s 1 -c 1, tools/quick_bench.py $args --arith 1"; python tools/ncu_raw_summary.py gpurun_out/r2_${name}_raw.csv;
  echo; echo "per source line (tools/ncu_source_summary.py):"; python tools/ncu_source_summary.py /tmp/r2_${name}_src.csv 30; } > gpurun_out/r2_${name}_ncu_full.txt 2>&1
rm -f gpurun_out/r2_$name.ncu-rep
seg=$(grep -o "last-launch segments [0-9]*" gpurun_out/r2_prof_$name.log | grep -o "[0-9]*$")
python tools/ncu_to_json.py gpurun_out/r2_${name}_raw.csv "c4:fast:pipeline3" "$seg" "ncu --set full --clock-control none, tools/quick_bench.py $args --arith 1 (one launch = one pool batch of the bench.py workload); raw export: profiles/r2_${name}_raw.csv" | cut -c1-200
python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err; echo "bench c4 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_c4_n1.json")); r = d["roofline"]
print("c4 value %.0f ms %.2f e2e %.0f other %.0f | frac %.3f dram_frac %s l2_frac %.3f issue %s lanes %s inst/seg %s nodes %.2f tris %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["other_arith"]["value"], r["frac"], r["dram_frac"], r["l2_frac"], r["issue_slot_util"], r["lanes_per_inst"], r["thread_inst_per_segment"], r["nodes_per_segment"], r["tris_per_segment"]))
PY
