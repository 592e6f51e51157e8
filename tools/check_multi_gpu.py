#!/usr/bin/env python3
"""Run under torchrun on N GPUs: renders a small C1 frame split over the ranks with BOTH exchanges (NCCL reduce and the
fused peer-memory kernel ordered by epoch flags) and compares them with each other and with the single-process sum rendered
on rank 0; then the same check through the single-process multi-GPU context (ptb_multi) on rank 0.  The output is kept
under profiles/ (r2_multi_gpu_check_n*.txt)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    sys.path.insert(0, str(p))
import numpy as np
import torch
import torch.distributed as dist

import make_assets
import szakdolgozat_pathtracer_b200 as ptb
from scenes import load_config
from szakdolgozat_pathtracer_b200 import parallel

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
ctx = ptb.Context(local)
if rank == 0:
    make_assets.ensure("c1", small=True)
dist.barrier()
sc = load_config(ptb, make_assets, "c1", small=True)
handle, _ = ctx.accel_build(sc)
W, H, SUB = 192, 128, 3
n = W * H
stream = torch.cuda.current_stream().cuda_stream
cfg = ptb.default_render_cfg(spp_per_launch=4, max_depth=6, subframes_per_launch=SUB, accumulate_mode=1, write_frame=0)


def render_into(ptr, first):
    ctx.memset(ptr, 0, n * 16, stream=stream)
    p = ptb.make_params(W, H, subframe_index=first, dof=True)
    p.accum_buffer, p.frame_buffer, p.handle = ptr, None, handle
    ctx.launch(p, cfg, stream=stream)


first = parallel.subframe_block_for_rank(rank, world, SUB * world)[0]
# NCCL path
accum = torch.zeros((H, W, 4), dtype=torch.float32, device=dev)
frame = torch.zeros((H, W, 4), dtype=torch.uint8, device=dev)
render_into(accum.data_ptr(), first)
parallel.reduce_accumulator(accum, dst=0)
if rank == 0:
    ctx.resolve(accum.data_ptr(), accum.data_ptr(), frame.data_ptr(), n, parallel.resolve_scale(SUB * world), cfg, stream=stream)
torch.cuda.synchronize()
# fused peer-memory path, ordered by epoch flags (no NCCL): three steps, so that both accumulators of the double buffer and
# the re-use of the first one are exercised; every step renders the same subframes and must give the same frame
ex = parallel.PeerExchange(ctx, rank, world, n)
for step in range(3):
    render_into(ex.begin_step(), first)
    ex.resolve(SUB * world, cfg, stream)
torch.cuda.synchronize()
assert not ex.timed_out(stream), "a peer signal never arrived"
raw_out, raw_frame = ex.out_accum, ex.out_frame
dist.barrier()
ok = True
if rank == 0:
    a_nccl, f_nccl = accum.cpu().numpy(), frame.cpu().numpy()
    a_p2p = ctx.to_host(raw_out, (H, W, 4), np.float32)
    f_p2p = ctx.to_host(raw_frame, (H, W, 4), np.uint8)
    # single-process reference: all subframes summed on this GPU in rank order
    tot = np.zeros((H, W, 4), np.float32)
    tmp = ctx.alloc(n * 16)
    for r in range(world):
        render_into(tmp, parallel.subframe_block_for_rank(r, world, SUB * world)[0])
        ctx.synchronize(stream)
        tot[..., :3] += ctx.to_host(tmp, (H, W, 4), np.float32)[..., :3]
    want = tot[..., :3] * np.float32(parallel.resolve_scale(SUB * world))
    d1 = float(np.abs(a_p2p[..., :3] - want).max()); d2 = float(np.abs(a_nccl[..., :3] - want).max())
    df = int(np.abs(f_p2p.astype(int) - f_nccl.astype(int)).max())
    exact = bool(np.array_equal(a_p2p[..., :3], want))
    print(f"world {world}: p2p vs single-process sum max abs diff {d1:.3e} (bit-exact: {exact}), nccl {d2:.3e}, frame p2p vs nccl max LSB diff {df}, mean {want.mean():.4f}")
    ok = exact and d2 < 1e-4 * max(1.0, float(want.max())) and df <= 1 and want.mean() > 0.01
ex.close()
dist.barrier()
# the single-process multi-GPU context (ptb_multi, csrc/multi.cpp) over the same `world` devices, driven by rank 0 alone
if rank == 0:
    def single(launches, kw):
        d_a, d_f = ctx.alloc(n * 16), ctx.alloc(n * 4)
        ctx.memset(d_a, 0, n * 16)
        sf = 0
        for k in launches:
            p = ptb.make_params(W, H, subframe_index=sf, dof=True)
            p.accum_buffer, p.frame_buffer, p.handle = d_a, d_f, handle
            ctx.launch(p, ptb.default_render_cfg(subframes_per_launch=k, **kw)); sf += k
        ctx.synchronize()
        return ctx.to_host(d_a, (H, W, 4), np.float32), ctx.to_host(d_f, (H, W, 4), np.uint8)

    def multi(launches, kw, split):
        m = ptb.Multi(list(range(world)))
        m.accel_build(sc)
        root = m.root
        d_a, d_f = root.alloc(n * 16), root.alloc(n * 4)
        root.memset(d_a, 0, n * 16); root.synchronize()
        sf = 0
        for k in launches:
            p = ptb.make_params(W, H, subframe_index=sf, dof=True)
            p.accum_buffer, p.frame_buffer = d_a, d_f
            m.launch(p, ptb.default_render_cfg(subframes_per_launch=k, **kw), split); sf += k
        m.synchronize()
        out = root.to_host(d_a, (H, W, 4), np.float32), root.to_host(d_f, (H, W, 4), np.uint8)
        m.close()
        return out
    kw = dict(spp_per_launch=4, max_depth=6)
    launches = [world + 1, 2 * world, 1]
    a1, f1 = single(launches, kw)
    at, ft = multi(launches, kw, ptb.PTB_SPLIT_TILES)
    as_, fs = multi(launches, kw, ptb.PTB_SPLIT_SAMPLES)
    tiles_exact = bool(np.array_equal(a1.view(np.uint32), at.view(np.uint32)) and np.array_equal(f1, ft))
    rel = float((np.abs(as_[..., :3] - a1[..., :3]) / (np.abs(a1[..., :3]) + 1e-7)).max())
    dfs = int(np.abs(fs.astype(int) - f1.astype(int)).max())
    print(f"world {world}: ptb_multi tiles vs one GPU bit-identical: {tiles_exact}; ptb_multi samples vs one GPU max rel diff {rel:.3e}, frame max LSB diff {dfs}")
    ok = ok and tiles_exact and rel < 3e-6 and dfs <= 1
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
