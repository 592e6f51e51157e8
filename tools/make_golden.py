#!/usr/bin/env python3
"""Generates the committed golden fixtures under tests/golden/ from the REFERENCE ITSELF, run here:
  * abi_layout.json        sizeof/offsetof of optixSphere.h, printed by oracle/_ref/ref_probe (real CUDA vector types)
  * obj_<name>.json        triangle count + sha256 of the face-vertex stream tiny_obj_loader.h produces for each OBJ
  * ref_c1_small.npz       accum / frame / primary-hit buffers rendered by oracle/_ref/libref_pt.so (the reference's
                           optixSphere.cu compiled for the host) on the C1-small scene, reference literals, 2 subframes
  * ref_c2_crop.npz,       the same for a window of the C2 scene (monkey + albedo map, close camera) and of the C3 scene
    ref_c3_crop.npz        (suitcase with albedo/normal/roughness/metallic maps, close camera): the textured closest-hit
                           branches of optixSphere.cu:682-714
  * ref_glass_demo.npz      the procedural scene with a transparent sphere: the glass branch of optixSphere.cu:803-856
Needs /root/reference (oracle/_ref is built from it); the fixtures then travel to boxes that do not have it.
"""
import hashlib
import json
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "tools"):
    sys.path.insert(0, str(p))
import numpy as np

import make_assets
import orchelp as oh
import szakdolgozat_pathtracer_b200 as ptb
from scenes import load_config

GOLD = ROOT / "tests" / "golden"
GOLD.mkdir(parents=True, exist_ok=True)
oh.build_oracle()
assert oh.have_ref(), "oracle/_ref is missing: /root/reference must be present to regenerate the fixtures"

layout = json.loads(subprocess.run([str(oh.REF_PROBE), "layout"], capture_output=True, text=True, check=True).stdout)
layout.pop("end")
(GOLD / "abi_layout.json").write_text(json.dumps(layout, indent=1) + "\n")

for name in ("test", "monkey", "suitcase", "fish", "tower"):
    with tempfile.NamedTemporaryFile(suffix=".bin") as tf:
        info = json.loads(subprocess.run([str(oh.REF_PROBE), "obj", str(ROOT / "assets" / f"{name}.obj"), tf.name],
                                         capture_output=True, text=True, check=True).stdout)
        raw = Path(tf.name).read_bytes()
    info["sha256_face_vertex_stream"] = hashlib.sha256(raw).hexdigest()
    (GOLD / f"obj_{name}.json").write_text(json.dumps(info, indent=1) + "\n")

sc = load_config(ptb, make_assets, "c1", small=True)
osc = oh.OracleScene.from_ptb(sc, guard=True)
W, H = 96, 64
accum = np.zeros((H, W, 4), np.float32)
hits0 = None
for sf in range(2):
    p = ptb.make_params(W, H, subframe_index=sf, dof=True)
    accum, frame, hits, st, rc = oh.render("ref", osc, oh.params_from_ptb(p), oh.default_config("ref"), accum=accum)
    assert rc == 0
    if sf == 0:
        hits0 = hits.copy()
        seg0 = int(st.segments)
np.savez_compressed(GOLD / "ref_c1_small.npz", accum=accum, frame=frame, hits=hits0, segments0=np.int64(seg0))

# window crops of the textured scenes (the windows tests/test_oracle_pins.py renders with the oracle)
from scenes import CAMERAS
CROPS = {"c2": dict(res=(160, 90), camera="monkey_close", window=(40, 20, 120, 60)),
         "c3": dict(res=(192, 108), camera="suitcase_close", window=(60, 30, 132, 78))}
for name, c in CROPS.items():
    sc = load_config(ptb, make_assets, name)
    osc = oh.OracleScene.from_ptb(sc, guard=True)
    (W, H), (x0, y0, x1, y1) = c["res"], c["window"]
    p = ptb.make_params(W, H, subframe_index=0, dof=True, **CAMERAS[c["camera"]])
    accum, frame, hits, st, rc = oh.render("ref", osc, oh.params_from_ptb(p), oh.default_config("ref"), window=c["window"])
    assert rc == 0
    np.savez_compressed(GOLD / f"ref_{name}_crop.npz", accum=accum[y0:y1, x0:x1], frame=frame[y0:y1, x0:x1], hits=hits[y0:y1, x0:x1],
                        segments=np.int64(st.segments), window=np.array(c["window"]), res=np.array(c["res"]))

# the glass branch (optixSphere.cu:803-856): procedural scene with a transparent middle sphere, reference literals
from scenes import glass_demo_scene
sc = glass_demo_scene(ptb, make_assets)
osc = oh.OracleScene.from_ptb(sc, guard=True)
W, H = 96, 64
p = ptb.make_params(W, H, subframe_index=0, dof=True)
accum, frame, hits, st, rc = oh.render("ref", osc, oh.params_from_ptb(p), oh.default_config("ref"))
assert rc == 0
np.savez_compressed(GOLD / "ref_glass_demo.npz", accum=accum, frame=frame, hits=hits, segments=np.int64(st.segments))
print("golden fixtures written to", GOLD)
