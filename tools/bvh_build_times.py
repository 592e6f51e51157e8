#!/usr/bin/env python3
"""BVH build times (kernel-only CUDA-event timing of ptb_accel_build) for the BASELINE configurations: plain LBVH and LBVH + treelet SAH.
usage: bvh_build_times.py [configs...]  (default c2 c3 c5 c4)"""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    sys.path.insert(0, str(p))
import make_assets
import szakdolgozat_pathtracer_b200 as ptb
from scenes import load_config

ctx = ptb.Context(0)
out = {}
for name in (sys.argv[1:] or ["c2", "c3", "c5", "c4"]):
    sc = load_config(ptb, make_assets, name)
    for refine in (0, 1):
        best = None
        for _ in range(4):
            _, st = ctx.accel_build(sc, ptb.default_build_cfg(sah_refine=refine))
            best = st.build_ms if best is None else min(best, st.build_ms)
        out.setdefault(name, {})["lbvh+sah" if refine else "lbvh"] = dict(ms=best, nodes=st.num_nodes, depth=st.max_depth, triangles=st.num_triangles,
                                                                           mtris_per_s=st.num_triangles / best / 1e3)
    print(name, json.dumps(out[name]), flush=True)
    sc.close()
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / "bvh_build.json").write_text(json.dumps(out, indent=1))
