#!/bin/bash
# single-GPU evidence of the round: GPU test suite, bench lines of every BASELINE configuration, reference arm, ncu captures
mkdir -p gpurun_out
if [ "${SKIP_TESTS:-0}" != 1 ]; then
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/r2_gpu_tests.log 2>&1
grep -n "fast-mode gate\|passed\|failed\|^FAILED\|^ERROR" gpurun_out/r2_gpu_tests.log | cut -c1-330
fi
bash tools/profile_configs.sh 2>&1 | cut -c1-260
python bench.py > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/r2_bench_c2_n1.err; echo "bench c2 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench_c2_n1.err
for c in c2_close c3 c4 c5; do
  st=5; wu=3; [ $c = c5 ] && { st=1; wu=1; }   # one C5 step is a 4096-spp 4K frame (6.4 s)
  python bench.py --config $c --steps $st --warmup $wu --no-cpu-baseline > gpurun_out/r2_bench_${c}_n1.json 2> gpurun_out/r2_bench_${c}_n1.err; echo "bench $c rc=$?"
done
python - <<'PY'
import json
for c in ("c2", "c2_close", "c3", "c4", "c5"):
    try:
        d = json.load(open(f"gpurun_out/r2_bench_{c}_n1.json"))
        r = d["roofline"]
        print(c, "value %.0f ms %.2f e2e %.0f other %s | frac %.3f dram_frac %s l2_frac %.3f issue %s lanes %s inst/seg %s nodes %.2f tris %.2f" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], d["other_arith"] and round(d["other_arith"]["value"]), r["frac"], r["dram_frac"], r["l2_frac"],
            r["issue_slot_util"], r["lanes_per_inst"], r["thread_inst_per_segment"], r["nodes_per_segment"], r["tris_per_segment"]))
    except Exception as e:
        print(c, "no line", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1; echo "launch list rc=$?"
