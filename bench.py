#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native path-tracing core.

Workload (BASELINE.json configs[1], "C2"): monkey.obj + monkey_albedo.png under the
synthetic env2.exr (2048x1024, seed 2), 1920x1080, 64 spp, depth 8, reference camera.
One STEP = one full 64-spp frame = 8 subframes of 8 samples per pixel, rendered by ONE C-ABI call
ptb_launch() with subframes_per_launch = 8 (bit-identical to 8 consecutive calls, i.e. to 8 optixLaunch
iterations of the reference's render loop, optixSphere.cpp:1390-1437; tests/test_gpu_parity.py checks it).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path; --arith fast (default, the headline: the
                                                                 reference's own build mode, gated by tests/test_gpu_fast_mode.py)
                                                                 or --arith exact (bit-identical to the CPU oracle); the other
                                                                 mode is timed alongside and reported under "other_arith"
  python bench.py --impl reference [--gpus N] [--steps K] ...    the reference's own device file compiled for the
                                                                 host (oracle/_ref), timed on the box's CPU cores

N > 1 (one process per GPU under torchrun): the scene is replicated, rank r renders
a contiguous block of subframes of a 64*N-spp frame (weak scaling) into a sum-mode accumulator; the exchange is ONE
fused kernel per rank over peer memory (reduce-scatter -> tonemap -> gather, NVLink), ordered by epoch flags in
peer-mapped memory -- no NCCL on the data path or for ordering (`--exchange nccl` keeps the ncclReduce variant).
Strong scaling (the fixed 64-spp frame split by samples and by tiles) is timed too and reported under "strong".

Prints ONE JSON line (rank 0).  `value` = segments of all ranks / max-over-ranks device time.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    sys.path.insert(0, str(p))

import numpy as np

# BASELINE.json configs.  The headline metric is quoted on C2; the others are selectable
# with --config (one step = ONE batched launch of `subframes` subframes of `spp` samples per pixel).
CONFIGS = {
    "c2": dict(scene="c2", res=(1920, 1080), spp=8, subframes=8, depth=8, camera="default",
               text="C2: monkey.obj+monkey_albedo.png, env2 2048x1024 (synthetic, seed 2), 1920x1080, 64 spp (8 subframes x 8), depth 8, reference camera, DoF on"),
    "c2_close": dict(scene="c2", res=(1920, 1080), spp=8, subframes=8, depth=8, camera="monkey_close",
                     text="C2 scene with the close camera of SURVEY.md section 8d (eye (0,1.2,3.2) -> (0,0.7,0)), 1920x1080, 64 spp, depth 8"),
    "c3": dict(scene="c3", res=(3840, 2160), spp=8, subframes=32, depth=8, camera="default",
               text="C3: suitcase.obj full PBR (real metallic/normal/roughness, synthetic albedo), env3 4096x2048, 3840x2160, 256 spp (32 subframes x 8), depth 8"),
    "c4": dict(scene="c4", res=(1920, 1080), spp=8, subframes=16, depth=8, camera="default",
               text="C4: fish+tower+5 synthetic 0.33-1.3 M triangle stand-ins (4.6 M triangles), env4, 1920x1080, 128 spp (16 subframes x 8), depth 8"),
    "c5": dict(scene="c5", res=(3840, 2160), spp=64, subframes=64, depth=8, camera="default",
               text="C5: synthetic model.obj (0.33 M triangles, uvs) + albedo, env5 4096x2048, 3840x2160, 4096 spp (64 subframes x 64), depth 8"),
}
W, H, SPP_PER_LAUNCH, LAUNCHES_PER_STEP, DEPTH, SCENE, CAMERA, WORKLOAD = 1920, 1080, 8, 8, 8, "c2", "default", CONFIGS["c2"]["text"]


def select_config(name):
    global W, H, SPP_PER_LAUNCH, LAUNCHES_PER_STEP, DEPTH, SCENE, CAMERA, WORKLOAD
    c = CONFIGS[name]
    (W, H), SPP_PER_LAUNCH, LAUNCHES_PER_STEP, DEPTH = c["res"], c["spp"], c["subframes"], c["depth"]
    SCENE, CAMERA, WORKLOAD = c["scene"], c["camera"], c["text"]


def measured_peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms while the timed region runs."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.lines, self.proc, self.first = gpu_index, [], None, 0

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_running(self, timeout=3.0):
        """Blocks until nvidia-smi has printed its first sample (its start-up can take longer than a short timed region)."""
        t0 = time.time()
        while self.proc and not self.lines and time.time() - t0 < timeout:
            time.sleep(0.01)

    def mark(self):
        """Samples from here on belong to the timed region."""
        self.first = len(self.lines)

    def stop(self):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
        # the samples taken inside the timed region; a region shorter than the sampling period falls back to the samples taken
        # under the same load during the warm-up steps
        lines = self.lines[self.first:] if len(self.lines) > self.first else self.lines[-4:]
        sm, mx, reasons = [], [], set()
        for ln in lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_sample(oh, ptb, osc, rows, threads=0):
    """Times the reference's CPU implementation on a band of the C2 frame.
    oracle/_ref (the reference's optixSphere.cu compiled for the host) when present, else the oracle port."""
    from scenes import CAMERAS
    if threads <= 0:
        # all host cores this process may use; explicit because torchrun exports OMP_NUM_THREADS=1 to its workers
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    kind = "reference" if oh.have_ref() else "port"
    which = "ref" if kind == "reference" else "oracle"
    p = ptb.make_params(W, H, subframe_index=0, dof=True, **CAMERAS[CAMERA])
    y0 = max(0, min(H - rows, H // 2 - rows // 2 - 40))  # a band around the mesh; rows == H is the whole frame
    if kind == "reference":
        cfg = oh.default_config("ref", threads=threads)
        lits = "reference literals 10 spp/launch, depth 20 (compile-time constants, optixSphere.cu:323,360)"
    else:
        cfg = oh.default_config("oracle", spp_per_launch=SPP_PER_LAUNCH, max_depth=DEPTH, threads=threads)
        lits = f"{SPP_PER_LAUNCH} spp/launch, depth {DEPTH}"
    accum = np.zeros((H, W, 4), np.float32)
    _, _, _, st, rc = oh.render(which, osc, oh.params_from_ptb(p), cfg, accum=accum, window=(0, y0, W, y0 + rows), want_hits=False)
    if rc not in (0, 3):
        raise RuntimeError(f"CPU reference render failed rc={rc}")
    sample = f"one launch over rows {y0}..{y0 + rows - 1} of the {SCENE} {W}x{H} frame ({W * rows} px), {lits}, {st.segments} segments, {st.seconds:.2f} s"
    return dict(value=st.segments / st.seconds / 1e6, unit="Msegments/s", cores=int(st.threads), kind=kind, sample=sample), st


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import make_assets
    import orchelp as oh
    import szakdolgozat_pathtracer_b200 as ptb
    from scenes import load_config
    sc = load_config(ptb, make_assets, SCENE)
    osc = oh.OracleScene.from_ptb(sc, guard=True)
    rows = min(args.ref_rows, H)
    for _ in range(args.warmup):
        cpu_reference_sample(oh, ptb, osc, max(4, rows // 8))
    seg, sec, last = 0, 0.0, None
    for _ in range(args.steps):
        last, st = cpu_reference_sample(oh, ptb, osc, rows)
        seg += st.segments; sec += st.seconds
    value = seg / sec / 1e6
    last["value"] = value
    line = {
        "impl": "reference", "metric": "Msegments/s", "value": value, "unit": "Msegments/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "sample_per_step": last["sample"]},
        "cpu_baseline": last, "e2e": {"value": value, "unit": "Msegments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-rows", type=int, default=1080, help="rows of the frame per step of the CPU reference arm")
    ap.add_argument("--cpu-rows", type=int, default=1080, help="rows of the frame for the cpu_baseline sample of our arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline", type=int, default=3, help="1 global queues, 2 chunked stage kernels, 3 chunked fused (default), 4 persistent pool")
    ap.add_argument("--config", default="c2", choices=list(CONFIGS))
    ap.add_argument("--arith", default="fast", choices=["fast", "exact"], help="arithmetic mode of the headline number (ptb_render_cfg.arith_mode)")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the strong-scaling legs")
    ap.add_argument("--split", default="samples", choices=["samples", "tiles"],
                    help="N > 1: split the frame's subframes across ranks (weak scaling, default) or its rows (tile partitioning, strong scaling)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N > 1: fused peer-memory exchange (default) or NCCL reduce")
    ap.add_argument("--other-configs", default="default",
                    help="comma list of the other BASELINE workloads timed briefly after the headline one and reported under \"configs\" "
                         "(default: c2_close,c3,c4,c5 when --config is c2; \"none\" to skip)")
    ap.add_argument("--other-steps", type=int, default=2, help="timed steps of each --other-configs workload (after 1 warm-up step)")
    args = ap.parse_args()
    select_config(args.config)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    import make_assets
    import szakdolgozat_pathtracer_b200 as ptb
    from scenes import CAMERAS, load_config
    from szakdolgozat_pathtracer_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the ptb path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    ARITH = {"exact": ptb.PTB_ARITH_EXACT, "fast": ptb.PTB_ARITH_FAST}
    arith, other = args.arith, ("exact" if args.arith == "fast" else "fast")
    if args.pipeline not in (2, 3):   # the fast build exists for the chunked pipelines only
        arith, other = "exact", None

    others = []
    if args.other_configs == "default":
        others = ["c2_close", "c3", "c4", "c5"] if args.config == "c2" else []
    elif args.other_configs != "none":
        others = [c for c in args.other_configs.split(",") if c]
    for c in others:
        if c not in CONFIGS:
            raise SystemExit(f"bench.py: unknown config {c!r} in --other-configs")

    line = run_config(args, args.config, False, torch, dist, ptb, parallel, make_assets, CAMERAS, load_config, world, rank, local_rank, dev, ARITH, arith, other)
    # the other BASELINE workloads (north_star: every named scene at 1 / 2 / 4 / 8 GPUs), same launch path, briefly
    per_config = {}
    for c in others:
        select_config(c)
        per_config[c] = run_config(args, c, True, torch, dist, ptb, parallel, make_assets, CAMERAS, load_config, world, rank, local_rank, dev, ARITH, arith, None)
    if rank == 0:
        if others:
            line["configs"] = per_config
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_config(args, config_name, brief, torch, dist, ptb, parallel, make_assets, CAMERAS, load_config, world, rank, local_rank, dev, ARITH, arith, other):
    """One workload on all ranks.  brief: only the headline timing (1 warm-up + --other-steps steps), returned as a small dict."""
    cpu_rows = min(args.cpu_rows, H)
    ctx = ptb.Context(local_rank)
    if rank == 0:
        make_assets.ensure(SCENE)
    if world > 1:
        dist.barrier()
    sc = load_config(ptb, make_assets, SCENE)
    handle, bst = ctx.accel_build(sc)

    n = W * H
    stream = torch.cuda.current_stream().cuda_stream
    accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    frame = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
    multi = world > 1
    tiles = multi and args.split == "tiles"

    def make_cfg(mode, subframes=LAUNCHES_PER_STEP, tiled=False, **kw):
        il = dict(row_interleave_count=world, row_interleave_index=rank, row_interleave_height=16) if tiled else {}
        base = dict(spp_per_launch=SPP_PER_LAUNCH, max_depth=DEPTH, subframes_per_launch=subframes, pipeline=args.pipeline,
                    accumulate_mode=1 if multi else 0, write_frame=0 if multi else 1, arith_mode=ARITH[mode])
        base.update(il); base.update(kw)
        return ptb.default_render_cfg(**base)

    # N > 1 exchange.  "p2p": every rank reduces + tonemaps its slice of the frame straight out of the peers' accumulators
    # (CUDA IPC mappings over NVLink) and stores it into rank 0's buffers: ONE kernel per rank, ordered by epoch flags in
    # peer memory (parallel.PeerExchange).  "nccl": ncclReduce of the float4 accumulator, then ptb_resolve on rank 0.
    exchange, exchange_note = None, "none (single GPU)"
    if multi:
        exchange_note = "nccl reduce + ptb_resolve on rank 0"
        if args.exchange == "p2p":
            try:
                exchange = parallel.PeerExchange(ctx, rank, world, n)   # plain cudaMalloc buffers: exportable through CUDA IPC
                exchange_note = ("fused peer-memory reduce-scatter -> tonemap -> gather (ptb_resolve_peers_sync over CUDA IPC / NVLink), ordered by epoch "
                                 "flags in peer memory (no NCCL call per step), double-buffered accumulators")
            except Exception as e:  # capability probe at set-up time, outside every timed region
                print(f"[rank {rank}] peer-memory exchange unavailable ({e}); using the NCCL reduce", file=sys.stderr)
                exchange = None
        flag = torch.tensor([1 if exchange is not None else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and exchange is not None:
            exchange.close(); exchange = None
            exchange_note = "nccl reduce + ptb_resolve on rank 0 (a peer could not map the IPC handles)"
    frame_ptr = exchange.out_frame if (exchange is not None and rank == 0) else frame.data_ptr()
    result_accum_ptr = exchange.out_accum if (exchange is not None and rank == 0) else accum.data_ptr()

    def one_step(cfg_used, frame_subframes, tiled=False):
        """One frame of `frame_subframes` subframes over all ranks.  samples: rank r renders a contiguous block of them;
        tiles: every rank renders all of them for ITS strips (the rest of its accumulator stays zero, so the same
        sum-exchange doubles as the gather)."""
        if multi:
            acc_ptr = exchange.begin_step() if exchange is not None else accum.data_ptr()
            ctx.memset(acc_ptr, 0, n * 16, stream=stream)
            first = 0 if tiled else parallel.subframe_block_for_rank(rank, world, frame_subframes)[0]
        else:
            acc_ptr, first = accum.data_ptr(), 0   # a fresh frame: subframe 0 restarts the running average (cpp:267-278)
        p = ptb.make_params(W, H, subframe_index=first, dof=True, **CAMERAS[CAMERA])
        p.accum_buffer, p.frame_buffer, p.handle = acc_ptr, (None if multi else frame_ptr), handle
        if cfg_used.subframes_per_launch > 0:
            ctx.launch(p, cfg_used, stream=stream)
        if multi and exchange is not None:
            exchange.resolve(frame_subframes, cfg_used, stream)
        elif multi:
            parallel.reduce_accumulator(accum, dst=0)
            if rank == 0:
                ctx.resolve(acc_ptr, acc_ptr, frame_ptr, n, parallel.resolve_scale(frame_subframes), cfg_used, stream=stream)

    def sync_all():
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(cfg_used, frame_subframes, steps, warmup, tiled=False, sample_clocks=False):
        """W warm-up steps, then K steps between CUDA events on the launch stream; max over ranks; device-counted segments."""
        sampler = ClockSampler(local_rank) if (sample_clocks and rank == 0) else None
        if sampler:
            sampler.start()          # before the warm-up: nvidia-smi is up and sampling by the time the timed region starts
            sampler.wait_running()
        for _ in range(warmup):
            one_step(cfg_used, frame_subframes, tiled)
        sync_all()
        ctx.totals(reset=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        if sampler:
            sampler.mark()
        e0.record()
        for _ in range(steps):
            one_step(cfg_used, frame_subframes, tiled)
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        tot = ctx.totals(reset=True)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        segs = torch.tensor([float(tot["segments"])], dtype=torch.float64, device=dev)
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(segs, op=dist.ReduceOp.SUM)
        return float(t.item()), float(segs.item()), clocks

    # ---- headline: weak scaling (64 spp per GPU), arithmetic mode `arith` ----
    total_subframes = LAUNCHES_PER_STEP * (1 if (tiles or not multi) else world)   # subframes in the frame all ranks produce together
    cfg = make_cfg(arith, tiled=tiles)
    if brief:
        k_b = max(1, args.other_steps)
        bms, bseg, _ = timed(cfg, total_subframes, k_b, 1, tiled=tiles)
        timed_out = bool(exchange.timed_out(stream)) if exchange is not None else False
        if exchange is not None:
            sync_all()
            exchange.close()
        del accum, frame
        sc.close()
        ctx.close()
        out = {"workload": WORKLOAD, "value": bseg / (bms * 1e-3) / 1e6, "unit": "Msegments/s", "ms_per_step": bms / k_b, "steps": k_b, "warmup": 1,
               "arith_mode": arith, "segments_per_step": bseg / k_b, "spp_per_step_per_gpu": SPP_PER_LAUNCH * LAUNCHES_PER_STEP,
               "spp_per_s_1080p": SPP_PER_LAUNCH * total_subframes * k_b / (bms * 1e-3) * (W * H / (1920.0 * 1080.0)),
               "scaling": "strong" if tiles else "weak",
               "bvh": {"triangles": bst.num_triangles, "width": bst.bvh_width, "build_ms": bst.build_ms}}
        if timed_out:
            out["exchange_error"] = "a peer signal timed out: numbers of this workload are invalid"
        return out
    ms_max, seg_total, clocks = timed(cfg, total_subframes, args.steps, args.warmup, tiled=tiles, sample_clocks=True)
    value = seg_total / (ms_max * 1e-3) / 1e6
    gpu_launches = int(ctx.launch_stats().kernel_launches) * args.steps + (3 * args.steps if multi else 0)

    # the other arithmetic mode, same workload, fewer steps: reported beside the headline
    other_arith = None
    if other is not None:
        k_other = max(3, min(args.steps, 8))
        oms, oseg, _ = timed(make_cfg(other, tiled=tiles), total_subframes, k_other, 2, tiled=tiles)
        other_arith = {"arith": other, "value": oseg / (oms * 1e-3) / 1e6, "unit": "Msegments/s", "ms_per_step": oms / k_other, "steps": k_other,
                       "note": ("bit-identical to the CPU oracle (tests/test_gpu_parity.py)" if other == "exact" else
                                "FMA + MUFU in the shading code, gated by tests/test_gpu_fast_mode.py")}

    # ---- strong scaling (N > 1): the SAME 64-spp frame split over the ranks, by samples and by tiles ----
    strong = None
    if multi and not args.no_strong:
        strong = {"frame": f"{SPP_PER_LAUNCH * LAUNCHES_PER_STEP} spp of {W}x{H} in total, all ranks together", "arith": arith}
        k_s = max(3, min(args.steps, 8))
        mine = parallel.subframe_block_for_rank(rank, world, LAUNCHES_PER_STEP)
        sms, sseg, _ = timed(make_cfg(arith, subframes=len(mine)), LAUNCHES_PER_STEP, k_s, 2)
        strong["samples"] = {"ms_per_frame": sms / k_s, "value": sseg / (sms * 1e-3) / 1e6, "subframes_per_rank": -(-LAUNCHES_PER_STEP // world)}
        tms, tseg, _ = timed(make_cfg(arith, tiled=True), LAUNCHES_PER_STEP, k_s, 2, tiled=True)
        strong["tiles"] = {"ms_per_frame": tms / k_s, "value": tseg / (tms * 1e-3) / 1e6, "strip_rows": 16}

    # ---- roofline of the dominant kernel + stage shares (CUDA events inside ptb_launch, same stream) ----
    def profiled(pipeline, mode):
        cfgp = make_cfg(mode, tiled=tiles, pipeline=pipeline, profile_stages=1)
        acc = {k: 0.0 for k in ("raygen", "trace", "shade", "miss", "resolve", "total")}
        reps = max(1, min(args.steps, 3))
        scratch_acc = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
        for _ in range(reps):
            first = 0 if (tiles or not multi) else parallel.subframe_block_for_rank(rank, world, total_subframes)[0]
            pp = ptb.make_params(W, H, subframe_index=first, dof=True, **CAMERAS[CAMERA])
            pp.accum_buffer, pp.frame_buffer, pp.handle = scratch_acc.data_ptr(), (None if multi else frame.data_ptr()), handle
            ctx.launch(pp, cfgp, stream=stream)
            for kk, v in ctx.stage_ms().items():
                acc[kk] += v / reps
        return acc
    stage = profiled(args.pipeline, arith)
    # per-stage split from the same stages run as separate kernels (pipeline 2); explains where the time goes
    stage_split = profiled(2, arith) if args.pipeline in (3, 4) else stage
    # traversal work per segment, from one instrumented launch of subframe 0
    cfg_cnt = ptb.default_render_cfg(spp_per_launch=SPP_PER_LAUNCH, max_depth=DEPTH, write_frame=0, count_traversal=1, pipeline=args.pipeline, arith_mode=ARITH[arith])
    p = ptb.make_params(W, H, subframe_index=0, dof=True, **CAMERAS[CAMERA])
    scratch = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
    p.accum_buffer, p.frame_buffer, p.handle = scratch.data_ptr(), frame.data_ptr(), handle
    ctx.launch(p, cfg_cnt, stream=stream)
    lst = ctx.launch_stats()
    nodes_per_seg = lst.nodes_visited / max(lst.segments, 1)
    tris_per_seg = lst.tris_tested / max(lst.segments, 1)
    hit_frac = lst.hits / max(lst.segments, 1)
    seg_per_step = seg_total / (args.steps * world)
    # Algorithmic bytes per segment:
    #   traversal stage: list 4 + ray 32 read, hit record 16 + status 1 written, node bytes per node visit, 48 B per triangle
    #   shade (per hit): 80 state read + 64 written + 120 attributes;  miss: 48 read + 64 env taps + 32 pixsum + 64 regenerated ray
    node_bytes = 128.0 if bst.bvh_width == 4 else 64.0
    trace_bytes_per_seg = 4 + 32 + 16 + 1 + node_bytes * nodes_per_seg + 48.0 * tris_per_seg
    shade_bytes_per_seg = hit_frac * (80 + 64 + 120) + (1.0 - hit_frac) * (48 + 64 + 32 + 64)
    peak, peak_src = measured_peaks()
    l2_peak = ctx.microbench_read(32 << 20, 40)    # 32 MiB working set: L2 -> SM read bandwidth
    hbm_read = ctx.microbench_read(2 << 30, 4)     # 2 GiB working set: HBM read bandwidth of the same kernel
    if args.pipeline in (3, 4):
        kernel, kernel_ms, launches_k = ("k_chunk_fused" if args.pipeline == 3 else "k_pool_fused"), stage["trace"], 1
        bytes_per_seg = trace_bytes_per_seg + shade_bytes_per_seg
    else:
        kernel, kernel_ms = ("k_chunk_trace" if args.pipeline == 2 else "k_trace"), stage["trace"]
        launches_k = SPP_PER_LAUNCH * (DEPTH + 1)
        bytes_per_seg = trace_bytes_per_seg
    achieved = bytes_per_seg * seg_per_step / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    # what ncu measured for this kernel on this workload (profiles/r2_ncu_summary.json, written by tools/ncu_to_json.py from
    # the --set full capture kept next to it): DRAM bytes per launch, issue-slot use, lanes per instruction, instruction count
    ncu = {}
    nf = ROOT / "profiles" / "r2_ncu_summary.json"
    if nf.exists():
        try:
            ncu = json.loads(nf.read_text()).get(f"{args.config}:{arith}:pipeline{args.pipeline}", {})
        except Exception:
            ncu = {}
    traffic = ncu.get("dram_bytes_per_launch")
    ncu_seg = ncu.get("segments_per_launch")
    roofline = {
        "kernel": kernel, "arith": arith,
        "bound": "latency/issue",
        "bound_note": "no memory level is near its peak (see dram_frac, l2_frac): the kernel is bound by instruction issue and dependent-load latency "
                      "at 8-9 warps per scheduler; `frac` below is ALGORITHMIC bytes over the HBM copy peak (SURVEY.md section 8d) -- most of those bytes "
                      "are served by L1/L2, so it is not an HBM utilisation",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "traffic": traffic,
        "dram_frac": (traffic * (seg_per_step / ncu_seg) / (kernel_ms * 1e-3) / 1e9 / peak) if (traffic and ncu_seg and kernel_ms > 0) else None,
        "l2_peak_gbs": l2_peak, "l2_frac": achieved / l2_peak if l2_peak > 0 else None, "hbm_read_gbs_own_microbench": hbm_read,
        "issue_slot_util": ncu.get("issue_slot_util"), "lanes_per_inst": ncu.get("lanes_per_inst"),
        "thread_inst_per_segment": ncu.get("thread_inst_per_segment"), "warp_inst_per_launch": ncu.get("warp_inst"),
        "ncu_l2_throughput_pct": ncu.get("l2_throughput_pct"), "ncu_dram_throughput_pct": ncu.get("dram_throughput_pct"),
        "ncu_stalls_per_issue": ncu.get("stalls_per_issue"), "ncu_source": ncu.get("source"),
        "algorithmic_bytes_per_segment": bytes_per_seg, "algorithmic_bytes_per_launch": bytes_per_seg * seg_per_step / launches_k,
        "avg_launch_ms": kernel_ms / launches_k, "launches_per_step": launches_k,
        "nodes_per_segment": nodes_per_seg, "tris_per_segment": tris_per_seg, "hit_fraction": hit_frac,
        "share_of_step": kernel_ms / stage["total"] if stage["total"] > 0 else None,
        "stage_ms_per_step_separate_kernels": {k: v for k, v in stage_split.items()},
        "stage_share_separate_kernels": {k: (stage_split[k] / stage_split["total"] if stage_split["total"] > 0 else 0.0) for k in ("raygen", "trace", "shade", "miss", "resolve")},
    }

    # ---- end to end through the C ABI with HOST buffers (pinned): accum in, accum + frame out, every step ----
    h_accum = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
    h_frame = torch.zeros((H, W, 4), dtype=torch.uint8).pin_memory()
    import ctypes as C
    L = ptb.lib()

    def e2e_step():
        # the frame's accumulator goes up from the host (the reference keeps it on the device across launches; a host
        # caller of the C ABI owns it), the step renders, accumulator + 8-bit frame come back
        if rank == 0:
            L.ptb_copy_to_device(ctx._h, C.c_void_p(result_accum_ptr), C.c_void_p(h_accum.data_ptr()), C.c_size_t(n * 16), C.c_void_p(stream))
        one_step(cfg, total_subframes, tiles)
        if rank == 0:
            L.ptb_copy_to_host(ctx._h, C.c_void_p(h_accum.data_ptr()), C.c_void_p(result_accum_ptr), C.c_size_t(n * 16), C.c_void_p(stream))
            L.ptb_copy_to_host(ctx._h, C.c_void_p(h_frame.data_ptr()), C.c_void_p(frame_ptr), C.c_size_t(n * 4), C.c_void_p(stream))

    e2e_timing = "host wall clock around K steps incl. pinned h2d/d2h copies of every step, one stream, a device synchronize at the end"
    if not multi:
        # One GPU: the copies run on their own streams and the device / host buffers are double-buffered, so the upload of step
        # k + 1 and the download of step k overlap the render of the neighbouring step (every step still uploads its own
        # accumulator and downloads its own accumulator + frame inside the timed region; a step's frame restarts the running
        # average at subframe 0, so consecutive steps are independent frames).
        sets = [(accum, frame, h_accum, h_frame),
                (torch.zeros_like(accum), torch.zeros_like(frame), torch.zeros_like(h_accum).pin_memory(), torch.zeros_like(h_frame).pin_memory())]
        s_up, s_down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        ev = [[torch.cuda.Event() for _ in range(3)] for _ in range(2)]   # per buffer set: upload, render, download finished
        used = [False, False]
        cur = torch.cuda.current_stream()

        def e2e_step(k=0):
            b = k & 1
            d_acc, d_frm, h_acc, h_frm = sets[b]
            if used[b]:
                s_up.wait_event(ev[b][2])      # this set's previous download has finished
            L.ptb_copy_to_device(ctx._h, C.c_void_p(d_acc.data_ptr()), C.c_void_p(h_acc.data_ptr()), C.c_size_t(n * 16), C.c_void_p(s_up.cuda_stream))
            ev[b][0].record(s_up)
            cur.wait_event(ev[b][0])
            pe = ptb.make_params(W, H, subframe_index=0, dof=True, **CAMERAS[CAMERA])
            pe.accum_buffer, pe.frame_buffer, pe.handle = d_acc.data_ptr(), d_frm.data_ptr(), handle
            ctx.launch(pe, cfg, stream=stream)
            ev[b][1].record(cur)
            s_down.wait_event(ev[b][1])
            L.ptb_copy_to_host_async(ctx._h, C.c_void_p(h_acc.data_ptr()), C.c_void_p(d_acc.data_ptr()), C.c_size_t(n * 16), C.c_void_p(s_down.cuda_stream))
            L.ptb_copy_to_host_async(ctx._h, C.c_void_p(h_frm.data_ptr()), C.c_void_p(d_frm.data_ptr()), C.c_size_t(n * 4), C.c_void_p(s_down.cuda_stream))
            ev[b][2].record(s_down)
            used[b] = True
        e2e_timing = ("host wall clock around K steps incl. the pinned h2d (accumulator) and d2h (accumulator + frame) copies of every step; copies on their own "
                      "streams with double-buffered device and host buffers, so a step's copies overlap the neighbouring step's render; device synchronize at the end")
        e2e_step(0); e2e_step(1)
    else:
        e2e_step()
    sync_all()
    ctx.totals(reset=True)
    t0 = time.perf_counter()
    for s_ in range(args.steps):
        e2e_step(s_) if not multi else e2e_step()
    sync_all()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    segs = torch.tensor([float(ctx.totals(reset=True)["segments"])], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(segs, op=dist.ReduceOp.SUM)
    e2e_value = float(segs.item()) / float(t.item()) / 1e6
    e2e = {"value": e2e_value, "unit": "Msegments/s", "h2d_bytes_per_step": n * 16, "d2h_bytes_per_step": n * 16 + n * 4,
           "timing": e2e_timing}
    timed_out = bool(exchange.timed_out(stream)) if exchange is not None else False

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import orchelp as oh
        osc = oh.OracleScene.from_ptb(sc, guard=True)
        cpu_baseline, _ = cpu_reference_sample(oh, ptb, osc, cpu_rows)

    line = None
    if rank == 0:
        pool_gb = min(LAUNCHES_PER_STEP * W * H * 97, 2 << 30) / 1e9
        line = {
            "metric": "Msegments/s", "value": value, "unit": "Msegments/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if tiles else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "arith_mode": arith,
                       "arith_note": ("fast arithmetic: FMA contraction + MUFU rcp/sqrt/sin/cos in the shading code (the reference's own build is --use_fast_math); camera rays of the "
                                      "pinhole model and all ray-triangle tests stay IEEE-exact; gate: tests/test_gpu_fast_mode.py (primary-hit IDs, image error bounds vs the CPU oracle)"
                                      if arith == "fast" else "exact arithmetic: accumulation buffer and frame bit-identical to the CPU oracle (tests/test_gpu_parity.py)"),
                       "spp_per_step_per_gpu": SPP_PER_LAUNCH * LAUNCHES_PER_STEP, "split": ("tiles (16-row strips dealt round-robin)" if tiles else "samples") if multi else "none",
                       "l2": f"no flush: the path pool of one step is {pool_gb:.1f} GB of path state (> 126 MB L2) and is rewritten every wavefront iteration",
                       "pipeline": {1: "global queues", 2: "block-local wavefront, one kernel per stage and iteration", 3: "block-local wavefront, fused persistent kernel",
                                    4: "persistent path pool (blocks own positions, slots handed out from one counter)"}[args.pipeline],
                       "subframes_per_launch": LAUNCHES_PER_STEP,
                       "multi_gpu": ("scene replicated, subframes split by rank; exchange: " + exchange_note) if multi else "single GPU, reference accumulate mode",
                       "bvh": {"triangles": bst.num_triangles, "nodes": bst.num_nodes, "max_depth": bst.max_depth, "width": bst.bvh_width, "sah": bst.sah_cost, "build_ms": bst.build_ms}},
            "spp_per_s_1080p": SPP_PER_LAUNCH * total_subframes * args.steps / (ms_max * 1e-3) * (W * H / (1920.0 * 1080.0)),
            "segments_per_step": seg_total / args.steps,
            "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
            "other_arith": other_arith, "strong": strong,
        }
        if timed_out:
            line["exchange_error"] = "a peer signal timed out (flag wait gave up after 4 s): numbers of this run are invalid"
    if exchange is not None:
        torch.cuda.synchronize()
        if multi:
            dist.barrier()
        exchange.close()
    del accum, frame
    sc.close()
    ctx.close()
    return line if rank == 0 else None


if __name__ == "__main__":
    sys.exit(main())
