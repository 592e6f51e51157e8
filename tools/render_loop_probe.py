#!/usr/bin/env python3
"""The reference's render loop through the C ABI (one subframe per ptb_launch, running average + tonemapped frame,
optixSphere.cpp:1390-1437) at the reference's own frame sizes and literals: serial launches against overlapped ones
(ptb_render_cfg.overlap_lanes)."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests"), str(ROOT / "tools")]
import numpy as np

import make_assets
import szakdolgozat_pathtracer_b200 as ptb
from scenes import CAMERAS, load_config

ctx = ptb.Context(0)
sc = load_config(ptb, make_assets, "c2")
handle, _ = ctx.accel_build(sc)
out = []
for (W, H, spp, depth) in ((600, 400, 10, 20), (1600, 1200, 10, 20), (1920, 1080, 8, 8), (3840, 2160, 8, 8)):
    n = W * H
    d_accum, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
    N = 48 if n < 3_000_000 else 16
    for arith in (1, 0):
        row = {"frame": f"{W}x{H}", "spp_per_launch": spp, "max_depth": depth, "arith": "fast" if arith else "exact", "launches": N}
        ref = None
        for lanes in (1, 2, 3, 4):
            cfg = ptb.default_render_cfg(spp_per_launch=spp, max_depth=depth, arith_mode=arith, overlap_lanes=lanes)
            best = None
            for rep in range(3):
                ctx.memset(d_accum, 0, n * 16)
                ctx.synchronize(); ctx.totals(reset=True)
                t0 = time.perf_counter()
                for k in range(N):
                    p = ptb.make_params(W, H, subframe_index=k, dof=True, **CAMERAS["default"])
                    p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
                    ctx.launch(p, cfg)
                ctx.synchronize()
                dt = time.perf_counter() - t0
                seg = ctx.totals(reset=True)["segments"]
                best = dt if best is None or dt < best else best
            frame = ctx.to_host(d_frame, (H, W, 4), np.uint8)
            if ref is None:
                ref = frame
            row[f"lanes{lanes}"] = {"ms_per_launch": round(best / N * 1e3, 3), "msegments_per_s": round(seg / best / 1e6), "identical_to_serial": bool(np.array_equal(ref, frame))}
        out.append(row)
        print(json.dumps(row))
    ctx.free(d_accum); ctx.free(d_frame)
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "r2_render_loop_overlap.json").write_text(json.dumps(out, indent=1) + "\n")
