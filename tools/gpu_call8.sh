#!/bin/bash
# 2 GPUs: the multi-GPU paths (one process per GPU under torchrun with the flag-ordered exchange; ptb_multi in one process)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/check_multi_gpu.py > gpurun_out/r2_multi_gpu_check_n2.txt 2>&1; echo "check rc=$?"; grep "world" gpurun_out/r2_multi_gpu_check_n2.txt
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','strong','other_arith','e2e','gpu_launches')})
print(d.get('exchange_error'))
PY
timeout 600 python tools/c5_convergence.py --gpus 1,2 --rows 64 > gpurun_out/r2_c5_conv_n2.log 2>&1; echo "c5 rc=$?"; tail -4 gpurun_out/r2_c5_conv_n2.log | cut -c1-400
python -m pytest tests/test_cli.py tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -3
