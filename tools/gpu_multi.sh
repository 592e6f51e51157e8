#!/bin/bash
# The multi-GPU session of a round on one box with NG GPUs (NG=8 by default): on-hardware parity of both exchanges and of
# ptb_multi, bench.py lines (weak + strong legs) for C2 and C3, the C5 convergence run at 1/2/4/NG GPUs, NVLink byte counts
# of the fused exchange kernel.  Everything lands in gpurun_out/ (copied to profiles/r2_* afterwards).
NG=${NG:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511"
echo "== check_multi_gpu N=$NG"
timeout 400 $TR tools/check_multi_gpu.py > gpurun_out/r2_multi_gpu_check_n$NG.txt 2>&1; echo "rc=$?"; grep "^world" gpurun_out/r2_multi_gpu_check_n$NG.txt
echo "== bench c2 N=$NG"
timeout 600 $TR bench.py --gpus $NG --steps 10 --warmup 3 2> gpurun_out/r2_bench_c2_n$NG.err | grep '^{' > gpurun_out/r2_bench_c2_n$NG.json; echo "rc=$?"
echo "== bench c3 N=$NG"
timeout 900 $TR bench.py --gpus $NG --config c3 --steps 3 --warmup 3 2> gpurun_out/r2_bench_c3_n$NG.err | grep '^{' > gpurun_out/r2_bench_c3_n$NG.json; echo "rc=$?"
python - <<PY
import json
for c in ("c2", "c3"):
    try:
        d = json.load(open(f"gpurun_out/r2_bench_{c}_n$NG.json"))
        print(c, {k: d[k] for k in ("value", "ms_per_step", "n_gpus", "strong", "other_arith")}, d.get("exchange_error"))
    except Exception as e:
        print(c, "no bench line:", e)
PY
echo "== c5 convergence"
GP="1,2,4,8"; [ "$NG" -lt 8 ] && GP="1,2"
timeout 900 python tools/c5_convergence.py --gpus $GP --rows 128 > gpurun_out/r2_c5_conv.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/r2_c5_conv.log | cut -c1-400
echo "== NVLink traffic of the exchange kernel (single process, ptb_multi through the CLI)"
python -c "
import sys; sys.path[:0]=['.','tools']
import make_assets; print(make_assets.ensure('c2'))" > gpurun_out/c2_assets.txt
OBJ=assets/_gen/c2/monkey.obj; ENV=assets/_gen/c2/env2.exr
CLI="szakdolgozat_pathtracer_b200/ptb_render -f gpurun_out/r2_cli_n$NG.png --dim=1920x1080 -s 8 --depth 8 --batch $((NG*2)) --launches 2 --scene $OBJ --env $ENV --scale 1.0 --gpus $NG --fast"
$CLI; echo "cli rc=$?"
ncu --query-metrics 2>/dev/null | grep -i "^nvl" | head -20 > gpurun_out/r2_nvlink_metrics_available.txt
timeout 600 ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_aperture_peer.sum,lts__t_sectors_aperture_peer_op_read.sum,lts__t_sectors_aperture_peer_op_write.sum \
    --clock-control none -k regex:k_resolve_peers --csv --log-file gpurun_out/r2_nvlink_resolve_peers_n$NG.csv $CLI > gpurun_out/r2_nvlink_ncu.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/r2_nvlink_resolve_peers_n$NG.csv | cut -c1-300
nvidia-smi topo -m > gpurun_out/r2_topo_n$NG.txt 2>&1
