"""Pins of the CPU oracle: known-answer values, and bit-exact agreement with the REFERENCE ITSELF
(/root/reference/optixSphere.cu compiled for the host, oracle/_ref) -- live where it is built, and through the
committed fixture tests/golden/ref_c1_small.npz everywhere else."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

from scenes import load_config

ROOT = Path(__file__).resolve().parent.parent


def test_rng_known_answers(oh):
    """optixSphere.cu:24-35 from seed 0."""
    L = oh.load("oracle")
    states = [129708000, 3259027200, 1053809664, 4075414272, 626777792, 1724044800]
    uniforms = [0.0301999971, 0.758801401, 0.245359182, 0.948881328, 0.145933077, 0.40141046]
    s = 0
    for st, un in zip(states, uniforms):
        u = C.c_float()
        s = L.orc_rng_next(C.c_uint32(s), 1, C.byref(u))
        assert s == st
        assert abs(u.value - un) < 1e-9 * max(1.0, un) + 1e-8
    assert np.float32(4294967295) == np.float32(4294967296.0)  # (float)UINT_MAX == 2^32


def _pcg_word(x):
    state = (x * 747796405 + 2891336453) & 0xFFFFFFFF
    w = (((state >> ((state >> 28) + 4)) ^ state) * 277803737) & 0xFFFFFFFF
    return ((w >> 22) ^ w) & 0xFFFFFFFF


def _pcg_preimage(out):
    """Invert the (bijective) integer hash of optixSphere.cu:26-29."""
    w = out ^ (out >> 22)
    y = (w * pow(277803737, -1, 2**32)) & 0xFFFFFFFF
    k = (y >> 28) + 4
    s = y
    for _ in range(8):
        s = y ^ (s >> k)
    return ((s - 2891336453) * pow(747796405, -1, 2**32)) & 0xFFFFFFFF


def test_rng_saturation_rule(oh):
    """Exactly the 128 hashes >= 2^32-128 round to 2^32 as float; CUDA's cvt saturates them to 0xFFFFFFFF
    (u == 1.0), x86 wraps to 0.  Oracle rule R1 is CUDA's."""
    L = oh.load("oracle")
    assert sum(1 for v in range(2**32 - 300, 2**32) if np.float32(v) == np.float32(2**32)) == 128
    u = C.c_float()
    for out in (2**32 - 1, 2**32 - 128):
        x = _pcg_preimage(out)
        assert _pcg_word(x) == out
        assert L.orc_rng_next(C.c_uint32(x), 1, C.byref(u)) == 0xFFFFFFFF and u.value == 1.0
        assert L.orc_rng_next(C.c_uint32(x), 0, C.byref(u)) == 0 and u.value == 0.0
    x = _pcg_preimage(2**32 - 129)  # rounds down to 2^32 - 256: representable, no saturation
    assert L.orc_rng_next(C.c_uint32(x), 1, C.byref(u)) == 2**32 - 256 and u.value < 1.0


def test_detmath_accuracy(oh):
    L = oh.load("oracle")
    x = np.linspace(0, 6.2831855, 5001, dtype=np.float32)
    s, c = C.c_float(), C.c_float()
    es = ec = 0.0
    for v in x:
        L.orc_sincos(C.c_float(float(v)), C.byref(s), C.byref(c))
        es = max(es, abs(s.value - np.sin(np.float64(v)))); ec = max(ec, abs(c.value - np.cos(np.float64(v))))
    assert es < 2.5e-7 and ec < 2.5e-7
    ys = np.linspace(-1, 1, 2001, dtype=np.float32)
    assert max(abs(L.orc_asin(float(v)) - np.arcsin(np.float64(v))) for v in ys) < 5e-7
    rng = np.random.default_rng(3)
    yx = rng.random((4000, 2), dtype=np.float32) * 2 - 1
    assert max(abs(L.orc_atan2(float(a), float(b)) - np.arctan2(np.float64(a), np.float64(b))) for a, b in yx) < 6e-7


def test_tonemap_analytic(oh):
    out = (C.c_uint8 * 4)()
    cfg = oh.default_config("oracle")
    oh.load("oracle").orc_tonemap_pixel((C.c_float * 3)(0, 0, 0), C.byref(cfg), out)
    # tonemap(0) = DE/DF - E/F ~ 0, then contrast 1.25 about 0.5 pushes it below 0 -> clamped to black
    assert list(out) == [0, 0, 0, 255]
    oh.load("oracle").orc_tonemap_pixel((C.c_float * 3)(1e6, 1e6, 1e6), C.byref(cfg), out)
    assert list(out)[:3] == [255, 255, 255]


def _oracle_c1_small(ptb, oh, assets, sat_cuda):
    sc = load_config(ptb, assets, "c1", small=True)
    osc = oh.OracleScene.from_ptb(sc, guard=True)
    W, H = 96, 64
    accum = np.zeros((H, W, 4), np.float32)
    hits0 = seg0 = None
    for sf in range(2):
        p = ptb.make_params(W, H, subframe_index=sf, dof=True)
        accum, frame, hits, st, rc = oh.render("oracle", osc, oh.params_from_ptb(p), oh.default_config("oracle", sat_cuda=sat_cuda), accum=accum)
        assert rc == 0
        if sf == 0:
            hits0, seg0 = hits.copy(), int(st.segments)
    return osc, accum, frame, hits0, seg0


def test_oracle_matches_reference_fixture(ptb, oh, assets):
    """The restated oracle against buffers the reference's own optixSphere.cu produced (compiled for the host, run in
    the build container by tools/make_golden.py).  x86 float->uint semantics on both sides (sat_cuda=0)."""
    g = np.load(ROOT / "tests" / "golden" / "ref_c1_small.npz")
    _, accum, frame, hits, seg0 = _oracle_c1_small(ptb, oh, assets, sat_cuda=0)
    assert np.array_equal(hits, g["hits"])
    assert seg0 == int(g["segments0"])
    assert np.array_equal(accum.view(np.uint32), g["accum"].view(np.uint32))
    assert np.array_equal(frame, g["frame"])


def test_oracle_matches_live_reference(ptb, oh, assets):
    if not oh.have_ref():
        pytest.skip("oracle/_ref/libref_pt.so not built (no /root/reference on this box)")
    osc, accum, frame, hits, seg0 = _oracle_c1_small(ptb, oh, assets, sat_cuda=0)
    W, H = 96, 64
    racc = np.zeros((H, W, 4), np.float32)
    for sf in range(2):
        p = ptb.make_params(W, H, subframe_index=sf, dof=True)
        racc, rframe, rhits, st, rc = oh.render("ref", osc, oh.params_from_ptb(p), oh.default_config("ref"), accum=racc)
        assert rc == 0
    assert np.array_equal(accum.view(np.uint32), racc.view(np.uint32)) and np.array_equal(frame, rframe)
    # no-DoF and brute-force variants
    p = ptb.make_params(W, H, subframe_index=3, dof=False)
    a1, f1, h1, s1, _ = oh.render("oracle", osc, oh.params_from_ptb(p), oh.default_config("oracle", sat_cuda=0, use_bvh=0))
    a2, f2, h2, s2, _ = oh.render("ref", osc, oh.params_from_ptb(p), oh.default_config("ref", use_bvh=0))
    assert np.array_equal(a1.view(np.uint32), a2.view(np.uint32)) and np.array_equal(h1, h2) and s1.segments == s2.segments


def test_oracle_monkey_matches_live_reference(ptb, oh, assets):
    """Same check on the C2 scene (albedo map only, 15 746 triangles) on a crop."""
    if not oh.have_ref():
        pytest.skip("oracle/_ref/libref_pt.so not built")
    sc = load_config(ptb, assets, "c2")
    osc = oh.OracleScene.from_ptb(sc, guard=True)
    p = ptb.make_params(160, 90, subframe_index=0, dof=True, eye=(0.0, 1.2, 3.2), lookat=(0.0, 0.7, 0.0))
    win = (40, 20, 120, 60)
    a1, _, h1, s1, _ = oh.render("oracle", osc, oh.params_from_ptb(p), oh.default_config("oracle", sat_cuda=0), window=win)
    a2, _, h2, s2, _ = oh.render("ref", osc, oh.params_from_ptb(p), oh.default_config("ref"), window=win)
    assert s1.segments == s2.segments and np.array_equal(h1, h2)
    assert np.array_equal(a1.view(np.uint32), a2.view(np.uint32))
    assert (h1[20:60, 40:120] < 15744).mean() > 0.1


CROPS = {"c2": dict(res=(160, 90), camera="monkey_close", window=(40, 20, 120, 60), mesh_tris=15744),
         "c3": dict(res=(192, 108), camera="suitcase_close", window=(60, 30, 132, 78), mesh_tris=2204)}


def _oracle_crop(ptb, oh, assets, name, which="oracle"):
    from scenes import CAMERAS
    c = CROPS[name]
    sc = load_config(ptb, assets, name)
    osc = oh.OracleScene.from_ptb(sc, guard=True)
    (W, H), (x0, y0, x1, y1) = c["res"], c["window"]
    p = ptb.make_params(W, H, subframe_index=0, dof=True, **CAMERAS[c["camera"]])
    cfg = oh.default_config("oracle", sat_cuda=0) if which == "oracle" else oh.default_config("ref")
    a, f, h, st, rc = oh.render(which, osc, oh.params_from_ptb(p), cfg, window=c["window"])
    assert rc == 0
    return a[y0:y1, x0:x1], f[y0:y1, x0:x1], h[y0:y1, x0:x1], int(st.segments)


@pytest.mark.parametrize("name", ["c2", "c3"])
def test_oracle_matches_reference_crop_fixture(ptb, oh, assets, name):
    """Textured scenes: the oracle against buffers the reference's own optixSphere.cu produced (host build, generated by
    tools/make_golden.py).  C3 exercises all four maps: albedo, normal (decode + swizzle + blend, cu:687-701), roughness,
    metallic (cu:705-714)."""
    g = np.load(ROOT / "tests" / "golden" / f"ref_{name}_crop.npz")
    a, f, h, seg = _oracle_crop(ptb, oh, assets, name)
    assert seg == int(g["segments"]) and np.array_equal(h, g["hits"])
    assert np.array_equal(a.view(np.uint32), g["accum"].view(np.uint32))
    assert np.array_equal(f, g["frame"])
    assert (h < CROPS[name]["mesh_tris"]).mean() > 0.1  # the textured mesh is in the window


def test_oracle_c3_matches_live_reference(ptb, oh, assets):
    """Full-PBR scene against the reference compiled here (live counterpart of the c3 fixture)."""
    if not oh.have_ref():
        pytest.skip("oracle/_ref/libref_pt.so not built")
    a1, f1, h1, s1 = _oracle_crop(ptb, oh, assets, "c3", "oracle")
    a2, f2, h2, s2 = _oracle_crop(ptb, oh, assets, "c3", "ref")
    assert s1 == s2 and np.array_equal(h1, h2) and np.array_equal(f1, f2)
    assert np.array_equal(a1.view(np.uint32), a2.view(np.uint32))


def _oracle_glass(ptb, oh, assets, which):
    from scenes import glass_demo_scene
    sc = glass_demo_scene(ptb, assets)
    assert sc.material(2).transparent == 1
    osc = oh.OracleScene.from_ptb(sc, guard=True)
    p = ptb.make_params(96, 64, subframe_index=0, dof=True)
    cfg = oh.default_config("oracle", sat_cuda=0) if which == "oracle" else oh.default_config("ref")
    a, f, h, st, rc = oh.render(which, osc, oh.params_from_ptb(p), cfg)
    assert rc == 0
    return a, f, h, int(st.segments), sc


def test_oracle_glass_branch_matches_reference_fixture(ptb, oh, assets):
    """The transparent branch (optixSphere.cu:803-856: Schlick reflect / sutil refract + 0.8 alpha random_in_unit_sphere) against
    buffers the reference's own code produced on the host (tools/make_golden.py)."""
    g = np.load(ROOT / "tests" / "golden" / "ref_glass_demo.npz")
    a, f, h, seg, sc = _oracle_glass(ptb, oh, assets, "oracle")
    assert seg == int(g["segments"]) and np.array_equal(h, g["hits"]) and np.array_equal(f, g["frame"])
    assert np.array_equal(a.view(np.uint32), g["accum"].view(np.uint32))
    ids = sc.material_ids()
    assert (ids[h[h >= 0]] == 2).mean() > 0.03   # the glass sphere is in frame


def test_oracle_glass_branch_matches_live_reference(ptb, oh, assets):
    if not oh.have_ref():
        pytest.skip("oracle/_ref/libref_pt.so not built")
    a1, f1, h1, s1, _ = _oracle_glass(ptb, oh, assets, "oracle")
    a2, f2, h2, s2, _ = _oracle_glass(ptb, oh, assets, "ref")
    assert s1 == s2 and np.array_equal(h1, h2) and np.array_equal(f1, f2) and np.array_equal(a1.view(np.uint32), a2.view(np.uint32))


def test_det_pow_accuracy(oh):
    """det_powf (display transform) is the correctly rounded power on the exponents the reference uses."""
    L = oh.load("oracle")
    rng = np.random.default_rng(9)
    xs = np.concatenate([rng.random(20000, dtype=np.float32), np.linspace(0, 1, 2001, dtype=np.float32)])
    for y in (np.float32(1.0) / np.float32(2.2), np.float32(1.0) / np.float32(2.4)):
        got = np.array([L.orc_pow(float(x), float(y)) for x in xs], np.float32)
        want = np.power(xs.astype(np.float64), np.float64(y)).astype(np.float32)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert L.orc_pow(0.0, 0.4) == 0.0 and L.orc_pow(1.0, 0.4) == 1.0 and L.orc_pow(2.0, 10.0) == 1024.0


def test_one_row_textures_wrap_inside_the_image(oh):
    """Texel index rule R3 on 1x1 and one-row images: x0 = y0 = -1 gives a linear index of -w-1, which must wrap into the
    image (+w*h once leaves -1 for h == 1)."""
    L = oh.load("oracle")
    out = (C.c_float * 4)()
    for w, h in ((1, 1), (2, 1), (5, 1), (1, 3)):
        img = np.arange(w * h * 4, dtype=np.float32) + 1.0
        guard = np.concatenate([np.full(64, np.nan, np.float32), img, np.full(64, np.nan, np.float32)])
        tex = oh.OrcTexture(guard[64:].ctypes.data, w, h, 1, 0)
        for u, v in ((0.0, 0.0), (0.01, 0.99), (0.99, 0.01), (0.5, 0.5)):
            L.orc_sample_texture(C.byref(tex), C.c_float(u), C.c_float(v), out)
            assert all(np.isfinite(list(out))), (w, h, u, v)
        env = (C.c_float * 3)(0.3, -0.9, 0.1)
        L.orc_sample_env(guard[64:].ctypes.data_as(C.c_void_p), w, h, env, out)
        assert all(np.isfinite(list(out)))


def test_oracle_bvh_equals_brute_force(ptb, oh, assets):
    from scenes import random_rays
    sc = load_config(ptb, assets, "c2")
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    v = osc.vertices[:, :3]
    o, d = random_rays(np.random.default_rng(11), 3000, v[:-6].min(0), v[:-6].max(0))
    hit = 0
    for i in range(len(o)):
        a = oh.closest_hit("oracle", osc, o[i], d[i], use_bvh=1)
        b = oh.closest_hit("oracle", osc, o[i], d[i], use_bvh=0)
        assert a == b
        hit += a[0] >= 0
    assert hit > 600
