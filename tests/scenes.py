"""Scene/configuration helpers shared by the tests, smoke() and bench.py."""
from __future__ import annotations

import numpy as np


def load_config(ptb, assets, name, small=False, material_seed=4):
    """BASELINE.json configs c1..c5 as a ptb Scene."""
    cfg = assets.ensure(name, small=small)
    sc = ptb.Scene.load_obj(cfg["files"], scale=cfg["scale"], material_seed=material_seed)
    sc.set_env_file(cfg["env"])
    return sc


CAMERAS = {
    # the reference camera (optixSphere.cpp:104-107)
    "default": dict(eye=(0.0, 2.0, 6.0), lookat=(0.0, 0.0, 0.0)),
    # close cameras of SURVEY.md section 8d
    "monkey_close": dict(eye=(0.0, 1.2, 3.2), lookat=(0.0, 0.7, 0.0)),
    "suitcase_close": dict(eye=(0.0, 1.6, 4.0), lookat=(0.0, 0.5, 1.0)),
}


def random_rays(rng, n, lo, hi):
    """Rays from random origins in a box around the scene toward random points inside it."""
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    ext = hi - lo
    o = (lo - 0.5 * ext + rng.random((n, 3), dtype=np.float32) * 2.0 * ext).astype(np.float32)
    tgt = (lo + rng.random((n, 3), dtype=np.float32) * ext).astype(np.float32)
    d = tgt - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o, d.astype(np.float32)


def glass_demo_scene(ptb, assets):
    """The reference's procedural scene (ground + three spheres) with a material table that makes the middle sphere GLASS
    (HitGroupData.transparent, optixSphere.cu:803-856), the left one rough red and the right one metallic, under a small
    synthetic environment: the scene of the glass-branch parity tests and of tests/golden/ref_glass_demo.npz."""
    sc = ptb.Scene.demo()
    sc.set_materials([dict(diffuse_color=(0.5, 0.5, 0.5), specular=(1, 1, 1), roughness=0.8),
                      dict(diffuse_color=(0.9, 0.1, 0.1), specular=(1, 0, 0), roughness=0.3),
                      dict(diffuse_color=(0.9, 0.95, 1.0), specular=(1, 1, 1), roughness=0.12, transparent=True),
                      dict(diffuse_color=(0.8, 0.7, 0.3), specular=(1, 1, 1), roughness=0.25, metallic=True)])
    env = assets.make_env(7, 256, 128).astype(np.float32)
    sc.set_env_pixels(np.concatenate([env, np.ones_like(env[..., :1])], -1))
    return sc
