#!/usr/bin/env python3
"""Probe: how much do single-subframe launches gain when consecutive launches may overlap on the device?
Two contexts (two path pools) of one GPU on two streams, launches dealt round-robin, against one context on one stream."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests"), str(ROOT / "tools")]
import torch

import make_assets
import szakdolgozat_pathtracer_b200 as ptb
from scenes import CAMERAS, load_config

W, H, SPP, DEPTH, N = int(sys.argv[1]) if len(sys.argv) > 1 else 1920, int(sys.argv[2]) if len(sys.argv) > 2 else 1080, int(sys.argv[3]) if len(sys.argv) > 3 else 8, int(sys.argv[4]) if len(sys.argv) > 4 else 8, 32
sc = load_config(ptb, make_assets, "c2")
ctxs = [ptb.Context(0) for _ in range(3)]
handles = [c.accel_build(sc)[0] for c in ctxs]
streams = [torch.cuda.Stream() for _ in range(3)]
n = W * H
acc = [torch.zeros((H, W, 4), dtype=torch.float32, device="cuda") for _ in range(3)]
for arith in (1, 0):
    cfg = ptb.default_render_cfg(spp_per_launch=SPP, max_depth=DEPTH, accumulate_mode=1, write_frame=0, arith_mode=arith)
    for lanes in (1, 2, 3):
        for rep in range(2):
            torch.cuda.synchronize()
            for c in ctxs: c.totals(reset=True)
            t0 = time.perf_counter()
            for k in range(N):
                l = k % lanes
                p = ptb.make_params(W, H, subframe_index=k, dof=True, **CAMERAS["default"])
                p.accum_buffer, p.frame_buffer, p.handle = acc[l].data_ptr(), None, handles[l]
                ctxs[l].launch(p, cfg, stream=streams[l].cuda_stream)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        seg = sum(c.totals(reset=True)["segments"] for c in ctxs)
        print(f"{W}x{H} spp {SPP} depth {DEPTH} arith {arith} lanes {lanes}: {dt / N * 1e3:.3f} ms per launch, {seg / dt / 1e6:.0f} Msegments/s")
