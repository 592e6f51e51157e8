#!/usr/bin/env python3
"""Runs the BASELINE.json configurations C1..C5 on one GPU at their named resolution with a reduced
sample count, and spot-checks each against the CPU oracle on a crop (accum bit-exact).  Writes gpurun_out/configs.json."""
import argparse
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    sys.path.insert(0, str(p))
import numpy as np

import make_assets
import orchelp as oh
import szakdolgozat_pathtracer_b200 as ptb
from scenes import CAMERAS, load_config

CONFIGS = {
    "c1": dict(res=(512, 512), spp=1, depth=4, batch=1, camera="default"),
    "c2": dict(res=(1920, 1080), spp=8, depth=8, batch=8, camera="default"),
    "c2_close": dict(scene="c2", res=(1920, 1080), spp=8, depth=8, batch=8, camera="monkey_close"),
    "c3": dict(res=(3840, 2160), spp=8, depth=8, batch=4, camera="default"),
    "c3_close": dict(scene="c3", res=(3840, 2160), spp=8, depth=8, batch=4, camera="suitcase_close"),
    "c4": dict(res=(1920, 1080), spp=8, depth=8, batch=8, camera="default"),
    "c5": dict(res=(3840, 2160), spp=8, depth=8, batch=4, camera="default"),
}

ap = argparse.ArgumentParser()
ap.add_argument("names", nargs="*", default=list(CONFIGS))
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--crop", type=int, default=48)
a = ap.parse_args()
ctx = ptb.Context(0)
results = {}
for name in a.names:
    c = CONFIGS[name]
    t0 = time.time()
    sc = load_config(ptb, make_assets, c.get("scene", name))
    t_load = time.time() - t0
    best_build = None
    for refine in (0, 1):
        for _ in range(2):  # second build is warm
            handle, bst = ctx.accel_build(sc, ptb.default_build_cfg(sah_refine=refine))
        b = dict(nodes=bst.num_nodes, leaves=bst.num_leaves, depth=bst.max_depth, sah=bst.sah_cost, build_ms=bst.build_ms)
        results.setdefault(name, {})["build_refine%d" % refine] = b
    W, H = c["res"]
    n = W * H
    d_accum, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
    cfg = ptb.default_render_cfg(spp_per_launch=c["spp"], max_depth=c["depth"], subframes_per_launch=c["batch"], count_traversal=0)
    p = ptb.make_params(W, H, subframe_index=0, dof=True, **CAMERAS[c["camera"]])
    p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
    ctx.memset(d_accum, 0, n * 16)
    ctx.launch(p, cfg); ctx.synchronize()
    ctx.totals(reset=True)
    t0 = time.time()
    for _ in range(a.reps):
        ctx.launch(p, cfg)
    ctx.synchronize()
    dt = (time.time() - t0) / a.reps
    tot = ctx.totals(reset=True)
    seg = tot["segments"] / a.reps
    cfgc = ptb.default_render_cfg(spp_per_launch=c["spp"], max_depth=c["depth"], count_traversal=1, write_frame=0)
    ctx.launch(p, cfgc)
    st = ctx.launch_stats()
    frame = ctx.to_host(d_frame, (H, W, 4), np.uint8)
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    ptb.save_image(ROOT / "gpurun_out" / f"cfg_{name}.png", frame[::max(1, H // 540), ::max(1, W // 960)].copy())
    # parity crop: one subframe, same spp/depth, centre window
    hs = a.crop // 2
    win = (W // 2 - hs, H // 2 - hs - H // 8, W // 2 + hs, H // 2 + hs - H // 8)
    cfg1 = ptb.default_render_cfg(spp_per_launch=c["spp"], max_depth=c["depth"], write_frame=0)
    ctx.memset(d_accum, 0, n * 16)
    ctx.launch(p, cfg1)
    ga = ctx.to_host(d_accum, (H, W, 4), np.float32)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    t0 = time.time()
    ca, _, _, ost, rc = oh.render("oracle", osc, oh.params_from_ptb(p), oh.default_config("oracle", spp_per_launch=c["spp"], max_depth=c["depth"]),
                                  window=win, want_hits=False)
    x0, y0, x1, y1 = win
    bad = int((ga[y0:y1, x0:x1].view(np.uint32) != ca[y0:y1, x0:x1].view(np.uint32)).any(axis=2).sum())
    results[name].update(dict(triangles=sc.num_triangles, load_s=t_load, res=[W, H], spp=c["spp"] * c["batch"], depth=c["depth"],
                              camera=c["camera"], ms_per_launch=dt * 1e3, msegments_per_s=seg / dt / 1e6,
                              spp_per_s_1080p=c["spp"] * c["batch"] / dt * (n / (1920 * 1080)),
                              nodes_per_segment=st.nodes_visited / max(st.segments, 1), tris_per_segment=st.tris_tested / max(st.segments, 1),
                              parity_crop=dict(window=list(win), pixels=(x1 - x0) * (y1 - y0), accum_mismatch_pixels=bad,
                                               oracle_segments=int(ost.segments), oracle_seconds=ost.seconds)))
    print(name, json.dumps(results[name]), flush=True)
    ctx.free(d_accum); ctx.free(d_frame)
    sc.close()
(ROOT / "gpurun_out" / "configs.json").write_text(json.dumps(results, indent=1))
