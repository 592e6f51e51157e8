#!/usr/bin/env python3
"""Development aid: builds one configuration's BVH under several builder settings and compares the flattened trees."""
import ctypes as C
import hashlib
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    sys.path.insert(0, str(p))
import numpy as np

import make_assets
import szakdolgozat_pathtracer_b200 as ptb
from scenes import load_config

name = sys.argv[1] if len(sys.argv) > 1 else "c5"
ctx = ptb.Context(0)
sc = load_config(ptb, make_assets, name)
print("sizeof(BuildCfg)", C.sizeof(ptb.BuildCfg))
for morton in (30, 63):
    for refine in (0, 1):
        cfg = ptb.default_build_cfg(sah_refine=refine, morton_bits=morton)
        for _ in range(2):
            handle, st = ctx.accel_build(sc, cfg)
        nodes, tris = ctx.accel_read(handle)
        codes = nodes[:, 12:14].view(np.int32)
        live = codes[:, 0] != -1
        # surface-area cost of the mesh part only: nodes whose box is smaller than 10 units across (the floor is 400)
        lo = np.minimum(nodes[:, [0, 2, 8]], nodes[:, [4, 6, 10]]); hi = np.maximum(nodes[:, [1, 3, 9]], nodes[:, [5, 7, 11]])
        ext = hi - lo
        small = live & np.all(np.isfinite(ext), axis=1) & (ext.max(axis=1) < 10.0)
        area = (ext[:, 0] * ext[:, 1] + ext[:, 1] * ext[:, 2] + ext[:, 2] * ext[:, 0])
        print(f"{name} morton {cfg.morton_bits} refine {refine}: nodes {st.num_nodes} leaves {st.num_leaves} depth {st.max_depth} sah {st.sah_cost:.4f} "
              f"build {st.build_ms:.2f} ms, mesh-node area sum {area[small].sum():.4f} ({int(small.sum())} nodes), "
              f"md5 {hashlib.md5(nodes.tobytes()).hexdigest()[:10]} prim-order md5 {hashlib.md5(tris[:, 3].tobytes()).hexdigest()[:10]}")
