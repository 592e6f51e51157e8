#!/usr/bin/env python3
"""Quick device-side timing of one configuration (development aid; bench.py is the contract)."""
import argparse
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT / "tools"))
import numpy as np

import make_assets
import szakdolgozat_pathtracer_b200 as ptb
from scenes import CAMERAS, load_config

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="c2")
ap.add_argument("--camera", default="default")
ap.add_argument("--width", type=int, default=1920)
ap.add_argument("--height", type=int, default=1080)
ap.add_argument("--spp", type=int, default=8)
ap.add_argument("--depth", type=int, default=8)
ap.add_argument("--launches", type=int, default=4)
ap.add_argument("--count", type=int, default=0)
ap.add_argument("--leaf", type=int, default=4)
ap.add_argument("--batch", type=int, default=1, help="subframes_per_launch")
ap.add_argument("--stages", type=int, default=0)
ap.add_argument("--pipeline", type=int, default=0)
ap.add_argument("--refine", type=int, default=1)
ap.add_argument("--morton", type=int, default=30)
ap.add_argument("--treelet", type=int, default=256)
ap.add_argument("--bvh-width", type=int, default=0)
ap.add_argument("--arith", type=int, default=0, help="0 exact, 1 fast")
a = ap.parse_args()

ctx = ptb.Context(0)
sc = load_config(ptb, make_assets, a.config)
t0 = time.time()
handle, bst = ctx.accel_build(sc, ptb.default_build_cfg(max_leaf_size=a.leaf, sah_refine=a.refine, morton_bits=a.morton, treelet_size=a.treelet, bvh_width=a.bvh_width))
print(f"build: {bst.num_triangles} tris, {bst.num_nodes} nodes, {bst.num_leaves} leaves, depth {bst.max_depth}, sah {bst.sah_cost:.2f} (mesh subtree {bst.sah_cost_mesh:.2f}), width {bst.bvh_width}, "
      f"{bst.build_ms:.3f} ms device, {1e3 * (time.time() - t0):.1f} ms wall incl. upload")
W, H = a.width, a.height
n = W * H
d_accum, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
ctx.memset(d_accum, 0, n * 16)
cfg = ptb.default_render_cfg(spp_per_launch=a.spp, max_depth=a.depth, count_traversal=a.count, subframes_per_launch=a.batch,
                             profile_stages=a.stages, pipeline=a.pipeline, arith_mode=a.arith)
for rep in range(2):
    seg = 0
    ctx.synchronize()
    t0 = time.time()
    for sf in range(0, a.launches * a.batch, a.batch):
        p = ptb.make_params(W, H, subframe_index=sf, dof=True, **CAMERAS[a.camera])
        p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
        ctx.launch(p, cfg)
    ctx.synchronize()
    dt = time.time() - t0
st = ctx.launch_stats()
seg = st.segments * a.launches
print(f"{a.config}/{a.camera} arith {a.arith} {W}x{H} spp {a.spp} depth {a.depth}: {dt / a.launches * 1e3:.2f} ms/launch, "
      f"~{seg / dt / 1e6:.1f} Msegments/s (last-launch segments {st.segments}, iterations {st.iterations}, "
      f"hits {st.hits}, misses {st.misses}), {a.spp * a.launches * a.batch / dt * (n / (1920 * 1080)):.1f} 1080p-spp/s")
if a.stages:
    print("stage ms (last launch):", {k: round(v, 3) for k, v in ctx.stage_ms().items()})
if a.count:
    print(f"nodes/seg {st.nodes_visited / st.segments:.2f}, tris/seg {st.tris_tested / st.segments:.2f}")
frame = ctx.to_host(d_frame, (H, W, 4), np.uint8)
out = ROOT / "gpurun_out"
out.mkdir(exist_ok=True)
ptb.save_image(out / f"{a.config}_{a.camera}.png", frame)
print("mean frame", frame[..., :3].mean(axis=(0, 1)))
