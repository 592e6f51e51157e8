#!/bin/bash
# bench.py exactly as the driver launches it for N > 1 (NG ranks), plus the reference arm under torchrun
NG=${NG:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533"
( time timeout 900 $TR bench.py --gpus $NG --steps 10 --warmup 3 2> gpurun_out/r2_bench_default_n$NG.err | grep '^{' > gpurun_out/r2_bench_default_n$NG.json ) 2>&1 | grep real; echo "rc=$?"
tail -5 gpurun_out/r2_bench_default_n$NG.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_default_n$NG.json"))
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "gpu_launches")}, d["e2e"]["value"], d.get("exchange_error"))
print("strong", d["strong"])
for k, v in (d.get("configs") or {}).items(): print(k, round(v["value"], 1), round(v["ms_per_step"], 2), v.get("exchange_error"))
PY
