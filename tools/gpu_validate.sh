#!/bin/bash
# what the driver runs at round end, in small: GPU test suite, smoke(), the bench line
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_validate_tests.log 2>&1; tail -6 gpurun_out/r2_validate_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --steps 10 > gpurun_out/r2_validate_bench.json 2> gpurun_out/r2_validate_bench.err ) 2>&1 | grep real; echo "bench rc=$?"; tail -3 gpurun_out/r2_validate_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_validate_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['other_arith']['value'], d['roofline']['frac'], d['roofline']['dram_frac'], d['clocks'])
for k,v in (d.get('configs') or {}).items(): print(k, round(v['value'],1), round(v['ms_per_step'],2), v['bvh'])
PY
