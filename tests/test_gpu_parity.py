"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star): primary-hit triangle IDs bit-exact, accumulation
buffer bit-exact (the kernels and the oracle share one fixed IEEE operation order
and the same deterministic sin/cos/atan2/asin), 8-bit frame bit-exact (the display
transform's pow is the double-precision det_powf on both sides).
"""
import ctypes as C

import numpy as np
import pytest

from scenes import CAMERAS, load_config, random_rays

pytestmark = pytest.mark.gpu

# every parity test runs on all four pipelines (global queues / chunked stages / chunked fused / persistent pool): the
# results must be bit-identical because a slot's arithmetic never depends on the order in which slots are processed
PIPELINE = 0


@pytest.fixture(autouse=True, params=[1, 2, 3, 4], ids=["queues", "chunk-stages", "chunk-fused", "pool-fused"])
def _pipeline(request):
    global PIPELINE
    PIPELINE = request.param
    yield
    PIPELINE = 0


def _render_gpu(ptb, ctx, handle, W, H, cfg_kw, subframes=1, dof=True, camera="default", accum0=None):
    n = W * H
    d_accum, d_frame, d_hits = ctx.alloc(n * 16), ctx.alloc(n * 4), ctx.alloc(n * 4)
    try:
        if accum0 is not None:
            ctx.to_device(d_accum, accum0)
        else:
            ctx.memset(d_accum, 0, n * 16)
        ctx.memset(d_hits, 0xFF, n * 4)
        stats = []
        for sf in range(subframes):
            p = ptb.make_params(W, H, subframe_index=sf, dof=dof, **CAMERAS[camera])
            p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
            cfg = ptb.default_render_cfg(aux_primary_hit=d_hits if sf == 0 else None, pipeline=PIPELINE, **cfg_kw)
            ctx.launch(p, cfg)
            stats.append(ctx.launch_stats())
        accum = ctx.to_host(d_accum, (H, W, 4), np.float32)
        frame = ctx.to_host(d_frame, (H, W, 4), np.uint8)
        hits = ctx.to_host(d_hits, (H, W), np.int32)
    finally:
        for b in (d_accum, d_frame, d_hits):
            ctx.free(b)
    return accum, frame, hits, stats


def _render_cpu(oh, ptb, osc, W, H, cfg_kw, subframes=1, dof=True, camera="default"):
    accum = np.zeros((H, W, 4), np.float32)
    hits0 = None
    seg = 0
    for sf in range(subframes):
        p = ptb.make_params(W, H, subframe_index=sf, dof=dof, **CAMERAS[camera])
        a, frame, hits, st, rc = oh.render("oracle", osc, oh.params_from_ptb(p), oh.default_config("oracle", **cfg_kw), accum=accum)
        assert rc == 0
        seg += st.segments
        if sf == 0:
            hits0 = hits
    return accum, frame, hits0, seg


def test_device_math_bit_exact(ptb, ctx, oh):
    rng = np.random.default_rng(7)
    L = oh.load("oracle")
    # RNG: random seeds, plus ALL 128 seeds whose hash rounds to 2^32 as float (cvt.rzi.u32.f32 saturates: next state
    # 0xFFFFFFFF, u == 1.0) and their neighbours just below the saturation range
    from test_oracle_pins import _pcg_preimage, _pcg_word
    sat = np.array([_pcg_preimage(v) for v in range(2**32 - 128, 2**32)], dtype=np.uint32)
    near = np.array([_pcg_preimage(v) for v in range(2**32 - 400, 2**32 - 128)], dtype=np.uint32)
    assert all(_pcg_word(int(x)) == v for x, v in zip(sat, range(2**32 - 128, 2**32)))
    seeds = np.concatenate([sat, near, np.arange(0, 4096, dtype=np.uint32), rng.integers(0, 2**32, 60000, dtype=np.uint32)])
    out = ctx.test_device_math(0, seeds.view(np.float32).reshape(-1, 1), 2)
    assert np.all(out[:128, 0:1].view(np.uint32) == 0xFFFFFFFF) and np.all(out[:128, 1] == 1.0)
    assert np.all(out[128:400, 1] < 1.0)
    for i in list(range(0, 400)) + list(range(400, len(seeds), 97)):
        u = C.c_float()
        nxt = L.orc_rng_next(C.c_uint32(int(seeds[i])), 1, C.byref(u))
        assert out[i, 0:1].view(np.uint32)[0] == nxt and out[i, 1] == u.value
    x = np.concatenate([np.linspace(0, 6.2831855, 20001, dtype=np.float32), rng.random(20000, dtype=np.float32) * 6.2831855])
    sc = ctx.test_device_math(1, x.reshape(-1, 1), 2)
    s, c = C.c_float(), C.c_float()
    ref = np.zeros_like(sc)
    for i, v in enumerate(x):
        L.orc_sincos(C.c_float(float(v)), C.byref(s), C.byref(c))
        ref[i] = (s.value, c.value)
    assert np.array_equal(sc.view(np.uint32), ref.view(np.uint32))
    assert np.abs(ref[:, 0] - np.sin(x.astype(np.float64))).max() < 3e-7
    yx = (rng.random((20000, 2), dtype=np.float32) * 2 - 1).astype(np.float32)
    at = ctx.test_device_math(2, yx, 1)[:, 0]
    ref = np.array([L.orc_atan2(float(a), float(b)) for a, b in yx], np.float32)
    assert np.array_equal(at.view(np.uint32), ref.view(np.uint32))
    assert np.abs(ref - np.arctan2(yx[:, 0].astype(np.float64), yx[:, 1].astype(np.float64))).max() < 5e-7
    xs = np.concatenate([np.linspace(-1, 1, 10001, dtype=np.float32), (rng.random(10000, dtype=np.float32) * 2 - 1)])
    asn = ctx.test_device_math(3, xs.reshape(-1, 1), 1)[:, 0]
    ref = np.array([L.orc_asin(float(a)) for a in xs], np.float32)
    assert np.array_equal(asn.view(np.uint32), ref.view(np.uint32))
    # texel widening: the division-free b / 255.0f (kernels.cuh: unit_from_u8) against the device's own IEEE division and numpy
    b = np.arange(256, dtype=np.float32)
    u8 = ctx.test_device_math(4, b.reshape(-1, 1), 2)
    assert np.array_equal(u8[:, 0].view(np.uint32), u8[:, 1].view(np.uint32))
    assert np.array_equal(u8[:, 0].view(np.uint32), (b / np.float32(255.0)).view(np.uint32))
    # display-transform pow: device det_powf == oracle det_powf == correctly rounded x^y
    px = np.concatenate([rng.random(30000, dtype=np.float32), np.linspace(0, 1, 4001, dtype=np.float32), np.float32([0.0, 1.0, 1e-30, 1e-42, 2.5, 1e30])])
    for y in (np.float32(1.0) / np.float32(2.2), np.float32(1.0) / np.float32(2.4)):
        pw = ctx.test_device_math(5, np.stack([px, np.full_like(px, y)], 1), 1)[:, 0]
        ref = np.array([L.orc_pow(float(v), float(y)) for v in px], np.float32)
        assert np.array_equal(pw.view(np.uint32), ref.view(np.uint32))
        inr = px <= 1.0
        assert np.array_equal(pw[inr].view(np.uint32), np.power(px[inr].astype(np.float64), np.float64(y)).astype(np.float32).view(np.uint32))


def _check_bvh(nodes, tris, n_tris, max_leaf):
    """Structural validation of the flattened BVH read back from the device."""
    codes = nodes[:, 12:14].view(np.int32)
    seen = np.zeros(n_tris, np.int32)
    stack = [0]
    visited = 0
    prim_ids = tris[:, 3].view(np.int32)
    tv = tris.reshape(-1, 3, 4)[:, :, :3]
    while stack:
        ni = stack.pop()
        visited += 1
        n = nodes[ni]
        boxes = [(np.array([n[0], n[2], n[8]]), np.array([n[1], n[3], n[9]])), (np.array([n[4], n[6], n[10]]), np.array([n[5], n[7], n[11]]))]
        for c in range(2):
            code = int(codes[ni, c])
            lo, hi = boxes[c]
            if code >= 0:
                # child box must contain the grandchildren boxes
                cn = nodes[code]
                clo = np.minimum([cn[0], cn[2], cn[8]], [cn[4], cn[6], cn[10]])
                chi = np.maximum([cn[1], cn[3], cn[9]], [cn[5], cn[7], cn[11]])
                assert np.all(clo >= lo) and np.all(chi <= hi)
                stack.append(code)
            else:
                k = ~code
                first, cnt = k >> 3, (k & 7) + 1
                assert cnt <= max(max_leaf, 1) and first + cnt <= n_tris
                seen[first:first + cnt] += 1
                v = tv[first:first + cnt].reshape(-1, 3)
                assert np.all(v.min(0) >= lo) and np.all(v.max(0) <= hi)
    assert np.all(seen == 1), "every triangle must be referenced by exactly one leaf"
    assert sorted(prim_ids.tolist()) == list(range(n_tris))
    return visited


def _check_bvh8(nodes8, tris8, n_tris):
    """Structural validation of the 8-wide quantised tree (csrc/bvh8.cuh) read back from the device: every triangle in
    exactly one leaf slot, internal children contiguous, and every child's quantised box contains what lies below it."""
    seen = np.zeros(n_tris, np.int32)
    prim_ids = tris8[:, 3].view(np.int32)
    tv = tris8.reshape(-1, 3, 4)[:, :, :3]
    visited, children = 0, 0
    def bounds(ni):
        """(lo, hi) of the triangles under node ni, checking the node on the way (a child's own quantised boxes may stick out
        of the box its parent holds for it: only the geometry has to be inside)"""
        nonlocal visited, children
        visited += 1
        q = nodes8[ni]
        p = q[0:3].view(np.float32).astype(np.float64)
        ew = int(q[3])
        step = np.array([2.0 ** (((ew >> (8 * d)) & 0xff) - 127) for d in range(3)])
        imask = ew >> 24
        child_base, tri_base = int(q[4]), int(q[5])
        by = q[6:20].view(np.uint8).reshape(7, 8)   # meta, lox, loy, loz, hix, hiy, hiz (little endian: byte k of a word pair = slot k)
        lo_all, hi_all = np.full(3, np.inf), np.full(3, -np.inf)
        rank = 0
        for s in range(8):
            meta = int(by[0, s])
            qlo, qhi = by[1:4, s].astype(np.float64), by[4:7, s].astype(np.float64)
            if meta == 0:
                assert not (imask >> s) & 1 and np.all(qlo == 255) and np.all(qhi == 0)
                continue
            children += 1
            blo, bhi = p + qlo * step, p + qhi * step
            if (meta & 0x1f) >= 24:
                assert (meta >> 5) == 1 and (meta & 0x1f) == 24 + s and (imask >> s) & 1
                clo, chi = bounds(child_base + rank)
                rank += 1
            else:
                assert not (imask >> s) & 1
                cnt = {1: 1, 3: 2, 7: 3}[meta >> 5]
                first = tri_base + (meta & 0x1f)
                seen[prim_ids[first:first + cnt]] += 1
                v = tv[first:first + cnt].reshape(-1, 3).astype(np.float64)
                clo, chi = v.min(0), v.max(0)
            assert np.all(blo <= clo) and np.all(bhi >= chi), (ni, s)
            lo_all, hi_all = np.minimum(lo_all, clo), np.maximum(hi_all, chi)
        assert rank == bin(imask).count("1")
        return lo_all, hi_all
    import sys
    sys.setrecursionlimit(10000)
    bounds(0)
    assert np.all(seen == 1), "every triangle must be referenced by exactly one leaf slot"
    return visited, children


@pytest.mark.parametrize("name,small", [("c1", True), ("c2", False)])
def test_bvh8_structure_and_ray_queries(ptb, ctx, oh, assets, name, small):
    """ptb_build_cfg.bvh_width = 8: the quantised 8-wide tree is structurally valid and gives the brute-force hits bit for bit."""
    if PIPELINE != 3:
        pytest.skip("the BVH does not depend on the render pipeline")
    sc = load_config(ptb, assets, name, small=small)
    handle, st = ctx.accel_build(sc, ptb.default_build_cfg(bvh_width=8))
    assert st.bvh_width == 8 and st.num_nodes8 > 0
    nodes8, tris8 = ctx.accel_read8(handle)
    visited, children = _check_bvh8(nodes8, tris8, sc.num_triangles)
    assert visited == st.num_nodes8 == len(nodes8)
    assert children / visited > 3.0, "the greedy collapse should fill most nodes"
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    v = osc.vertices[:, :3]
    lo, hi = v[:-6].min(0), v[:-6].max(0)
    rng = np.random.default_rng(321)
    o, d = random_rays(rng, 20000, lo, hi)
    d[::7, 0] = 0.0          # axis-parallel rays: 1 / d = inf
    d[3::11, 1] = -0.0
    prim, t, b1, b2 = ctx.trace_rays(handle, o, d)
    handle2, _ = ctx.accel_build(sc, ptb.default_build_cfg(bvh_width=2, max_leaf_size=3))
    for a, b in zip((prim, t, b1, b2), ctx.trace_rays(handle2, o, d)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    mism = 0
    for i in range(0, len(o), 2 if name == "c1" else 20):
        rp, rt, rb1, rb2 = oh.closest_hit("oracle", osc, o[i], d[i], use_bvh=0)
        if rp != prim[i]:
            mism += 1
            continue
        if rp >= 0:
            assert np.float32(rt) == t[i] and np.float32(rb1) == b1[i] and np.float32(rb2) == b2[i]
    assert mism == 0
    assert (prim >= 0).mean() > 0.2


@pytest.mark.parametrize("refine", [0, 1], ids=["lbvh", "lbvh+sah"])
@pytest.mark.parametrize("name,small", [("c1", True), ("c2", False)])
def test_bvh_structure_and_ray_queries(ptb, ctx, oh, assets, name, small, refine):
    if PIPELINE != 2:
        pytest.skip("the BVH does not depend on the render pipeline")
    sc = load_config(ptb, assets, name, small=small)
    handle, st = ctx.accel_build(sc, ptb.default_build_cfg(sah_refine=refine))
    assert st.num_triangles == sc.num_triangles and st.max_depth < 128
    nodes, tris = ctx.accel_read(handle)
    assert _check_bvh(nodes, tris, sc.num_triangles, 4) == st.num_nodes
    if refine and name == "c2":
        _, st0 = ctx.accel_build(ptb_scene_again := load_config(ptb, assets, name, small=small), ptb.default_build_cfg(sah_refine=0))
        assert st.sah_cost < st0.sah_cost, (st.sah_cost, st0.sah_cost)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    v = osc.vertices[:, :3]
    # leave the 400-unit floor out of the sampling box so that rays concentrate on the mesh
    lo, hi = v[:-6].min(0), v[:-6].max(0)
    rng = np.random.default_rng(123)
    o, d = random_rays(rng, 20000, lo, hi)
    prim, t, b1, b2 = ctx.trace_rays(handle, o, d)
    mism = 0
    for i in range(0, len(o), 1 if name == "c1" else 10):
        rp, rt, rb1, rb2 = oh.closest_hit("oracle", osc, o[i], d[i], use_bvh=0)
        if rp != prim[i]:
            mism += 1
            continue
        if rp >= 0:
            assert np.float32(rt) == t[i] and np.float32(rb1) == b1[i] and np.float32(rb2) == b2[i]
    assert mism == 0
    assert (prim >= 0).mean() > 0.2


def test_bvh_builder_options(ptb, ctx, oh, assets):
    """Builder variants must all give a valid tree and the same hits: 63-bit Morton keys, small treelets, leaf size 1 and 8,
    the 4-wide collapsed tree;
    and the huge-primitive split: the two floor triangles (the last two prims, optixSphere.cpp:598-646) sit in ONE leaf that
    is a child of the root."""
    if PIPELINE != 3:
        pytest.skip("the BVH does not depend on the render pipeline")
    sc = load_config(ptb, assets, "c2")
    n = sc.num_triangles
    rng = np.random.default_rng(5)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    v = osc.vertices[:, :3]
    o, d = random_rays(rng, 4000, v[:-6].min(0), v[:-6].max(0))
    ref = None
    for kw in (dict(), dict(morton_bits=63), dict(treelet_size=32), dict(max_leaf_size=1), dict(max_leaf_size=8), dict(sah_refine=0, morton_bits=63),
               dict(bvh_width=4), dict(bvh_width=4, sah_refine=0), dict(bvh_width=4, max_leaf_size=1), dict(bvh_width=2),
               dict(bvh_width=8), dict(bvh_width=8, sah_refine=0), dict(bvh_width=8, max_leaf_size=1), dict(bvh_width=8, morton_bits=63, treelet_size=64)):
        handle, st = ctx.accel_build(sc, ptb.default_build_cfg(**kw))
        nodes, tris = ctx.accel_read(handle)
        leaf_cap = min(kw.get("max_leaf_size", 4), 3) if kw.get("bvh_width") == 8 else kw.get("max_leaf_size", 4)   # the 8-wide tree holds leaves of <= 3
        assert _check_bvh(nodes, tris, n, leaf_cap) == st.num_nodes, kw
        assert st.bvh_width == kw.get("bvh_width", 2), kw
        if kw.get("max_leaf_size", 4) != 1:  # with leaf size 1 the two floor triangles are two leaves under one node next to the root
            codes = nodes[0, 12:14].view(np.int32)
            leaf = [c for c in codes if c < 0]
            assert len(leaf) == 1, f"root must have exactly one leaf child (the huge primitives): {codes} {kw}"
            first, cnt = (~leaf[0]) >> 3, ((~leaf[0]) & 7) + 1
            assert cnt == 2 and sorted(tris[first:first + 2, 3].view(np.int32).tolist()) == [n - 2, n - 1], kw
        got = ctx.trace_rays(handle, o, d)
        if ref is None:
            ref = got
            for i in range(0, len(o), 40):
                rp, rt, _, _ = oh.closest_hit("oracle", osc, o[i], d[i], use_bvh=0)
                assert rp == got[0][i] and (rp < 0 or np.float32(rt) == got[1][i])
        else:
            for a, b in zip(ref, got):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), kw


def test_c1_image_bit_exact_default_config(ptb, ctx, oh, assets):
    """Reference literals (10 spp, depth 20, DoF on), two subframes: accum bit-exact, frame within 1 LSB."""
    sc = load_config(ptb, assets, "c1", small=True)
    handle, _ = ctx.accel_build(sc)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    W, H = 160, 96
    ga, gf, gh, gst = _render_gpu(ptb, ctx, handle, W, H, {}, subframes=2)
    ca, cf, ch, cseg = _render_cpu(oh, ptb, osc, W, H, {}, subframes=2)
    assert np.array_equal(gh, ch), f"{(gh != ch).sum()} primary-hit mismatches"
    assert sum(s.segments for s in gst) == cseg
    bad = (ga.view(np.uint32) != ca.view(np.uint32)).any(axis=2)
    assert bad.sum() == 0, f"{bad.sum()} of {W * H} accum pixels differ; max abs diff {np.abs(ga - ca).max()}"
    assert np.array_equal(gf, cf), "8-bit frame must be bit-exact (det_powf on both sides)"


@pytest.mark.parametrize("dof", [False, True])
def test_c1_hit_ids_512(ptb, ctx, oh, assets, dof):
    """BASELINE config 1: 512x512, 1 spp, depth 4."""
    sc = load_config(ptb, assets, "c1", small=False)
    handle, _ = ctx.accel_build(sc)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    kw = dict(spp_per_launch=1, max_depth=4)
    ga, gf, gh, gst = _render_gpu(ptb, ctx, handle, 512, 512, kw, dof=dof)
    ca, cf, ch, cseg = _render_cpu(oh, ptb, osc, 512, 512, kw, dof=dof)
    assert np.array_equal(gh, ch)
    assert gst[0].segments == cseg
    assert np.array_equal(ga.view(np.uint32), ca.view(np.uint32))


def test_c2_monkey_parity_crop(ptb, ctx, oh, assets):
    """BASELINE config 2 scene (monkey + albedo map), close camera, 8 spp depth 8 at 320x180: bit-exact."""
    sc = load_config(ptb, assets, "c2")
    handle, st = ctx.accel_build(sc)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    kw = dict(spp_per_launch=8, max_depth=8)
    W, H = 320, 180
    ga, gf, gh, gst = _render_gpu(ptb, ctx, handle, W, H, kw, camera="monkey_close")
    ca, cf, ch, cseg = _render_cpu(oh, ptb, osc, W, H, kw, camera="monkey_close")
    assert (gh != ch).sum() == 0
    assert gst[0].segments == cseg
    bad = (ga.view(np.uint32) != ca.view(np.uint32)).any(axis=2)
    assert bad.sum() == 0, f"{bad.sum()} of {W * H} accum pixels differ"
    assert np.array_equal(gf, cf)
    assert (gh < 15744).mean() > 0.05  # the mesh is in frame


def test_sum_mode_and_resolve(ptb, ctx, oh, assets):
    """accumulate_mode=1 (sample-split): accum holds the sum of launch means; ptb_resolve divides and tonemaps."""
    sc = load_config(ptb, assets, "c1", small=True)
    handle, _ = ctx.accel_build(sc)
    W, H, n = 96, 64, 96 * 64
    kw = dict(spp_per_launch=4, max_depth=6, accumulate_mode=1, write_frame=0)
    ga, _, _, _ = _render_gpu(ptb, ctx, handle, W, H, kw, subframes=3)
    per = [_render_gpu(ptb, ctx, handle, W, H, dict(spp_per_launch=4, max_depth=6, write_frame=0), subframes=1)[0]]
    # subframe 0 alone equals the first term of the sum
    d_a, d_o, d_f = ctx.alloc(n * 16), ctx.alloc(n * 16), ctx.alloc(n * 4)
    try:
        ctx.to_device(d_a, ga)
        ctx.resolve(d_a, d_o, d_f, n, 1.0 / 3.0)
        ctx.synchronize()
        mean = ctx.to_host(d_o, (H, W, 4), np.float32)
        frame = ctx.to_host(d_f, (H, W, 4), np.uint8)
    finally:
        for b in (d_a, d_o, d_f):
            ctx.free(b)
    assert np.allclose(mean[..., :3], ga[..., :3] * np.float32(1.0 / 3.0), rtol=0, atol=0)
    assert frame[..., 3].min() == 255 and frame[..., :3].max() > 0
    assert np.all(ga[..., :3] >= per[0][..., :3] - 1e-6)


def test_batched_subframes_bit_identical_to_consecutive_launches(ptb, ctx, assets):
    """subframes_per_launch = n renders n subframes as one wavefront; accum/frame must equal n consecutive launches."""
    sc = load_config(ptb, assets, "c1", small=True)
    handle, _ = ctx.accel_build(sc)
    W, H, n = 200, 120, 200 * 120
    kw = dict(spp_per_launch=3, max_depth=6)
    seq_a, seq_f, _, seq_st = _render_gpu(ptb, ctx, handle, W, H, kw, subframes=5)
    d_accum, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
    try:
        ctx.memset(d_accum, 0, n * 16)
        seg = 0
        for first, cnt in ((0, 3), (3, 2)):  # two batches: 3 + 2 subframes
            p = ptb.make_params(W, H, subframe_index=first, dof=True)
            p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
            ctx.launch(p, ptb.default_render_cfg(subframes_per_launch=cnt, pipeline=PIPELINE, **kw))
            seg += ctx.launch_stats().segments
        a = ctx.to_host(d_accum, (H, W, 4), np.float32)
        f = ctx.to_host(d_frame, (H, W, 4), np.uint8)
    finally:
        ctx.free(d_accum); ctx.free(d_frame)
    assert seg == sum(s.segments for s in seq_st)
    assert np.array_equal(a.view(np.uint32), seq_a.view(np.uint32))
    assert np.array_equal(f, seq_f)


def test_chunk_sizes_bit_identical(ptb, ctx, assets):
    """The fused kernel picks its chunk size from the launch size (256 .. 2048 slots per block): every choice, and a ragged
    last chunk, must give the same accumulator, frame, hit IDs and segment count."""
    if PIPELINE != 3:
        pytest.skip("chunk_slots_per_thread is a knob of pipeline 3")
    sc = load_config(ptb, assets, "c2")
    handle, _ = ctx.accel_build(sc)
    W, H = 251, 67  # 16 817 slots: not a multiple of any chunk size
    ref = None
    for spt in (0, 1, 2, 4, 8):
        a, f, h, st = _render_gpu(ptb, ctx, handle, W, H, dict(spp_per_launch=3, max_depth=6, chunk_slots_per_thread=spt), subframes=2, camera="monkey_close")
        cur = (a.view(np.uint32), f, h, [s.segments for s in st])
        if ref is None:
            ref = cur
        else:
            assert all(np.array_equal(x, y) for x, y in zip(ref[:3], cur[:3])) and ref[3] == cur[3], spt
    n = W * H
    p = ptb.make_params(W, H)
    d_accum, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
    try:
        p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
        with pytest.raises(ptb.PtbError):
            ctx.launch(p, ptb.default_render_cfg(chunk_slots_per_thread=3, pipeline=3))
    finally:
        ctx.free(d_accum); ctx.free(d_frame)


def test_wide_bvh_render_bit_identical(ptb, ctx, oh, assets):
    """4-wide traversal (the default for large scenes) and the 8-wide quantised one against the 2-wide one and the oracle: same
    image, bit for bit."""
    if PIPELINE not in (1, 3, 4):
        pytest.skip("pipeline 2 shares the traversal code of pipeline 3")
    sc = load_config(ptb, assets, "c2")
    W, H = 160, 90
    kw = dict(spp_per_launch=3, max_depth=6)
    res = []
    for width in (2, 4, 8):
        handle, st = ctx.accel_build(sc, ptb.default_build_cfg(bvh_width=width))
        assert st.bvh_width == width
        a, f, h, stl = _render_gpu(ptb, ctx, handle, W, H, kw, camera="monkey_close")
        res.append((a.view(np.uint32), f, h, stl[0].segments))
    for other in res[1:]:
        assert all(np.array_equal(x, y) for x, y in zip(res[0][:3], other[:3])) and res[0][3] == other[3]
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    ca, cf, ch, cseg = _render_cpu(oh, ptb, osc, W, H, kw, camera="monkey_close")
    assert np.array_equal(res[1][0], ca.view(np.uint32)) and np.array_equal(res[1][2], ch) and res[1][3] == cseg


def test_demo_scene_parity(ptb, ctx, oh, assets):
    """The reference's procedural scene (degenerate pole triangles, roughness-0 spheres clamped to 0.015): bit-exact."""
    sc = ptb.Scene.demo()
    sc.set_env_pixels(assets.make_env(7, 256, 128).astype(np.float32).repeat(1, axis=2)[..., [0, 1, 2, 2]])
    handle, st = ctx.accel_build(sc)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    kw = dict(spp_per_launch=4, max_depth=10)
    ga, gf, gh, gst = _render_gpu(ptb, ctx, handle, 240, 160, kw)
    ca, cf, ch, cseg = _render_cpu(oh, ptb, osc, 240, 160, kw)
    assert np.array_equal(gh, ch) and gst[0].segments == cseg
    assert np.array_equal(ga.view(np.uint32), ca.view(np.uint32)) and np.array_equal(gf, cf)
    assert len(np.unique(gh)) > 100


def _tri_scene(ptb, assets, tris):
    sc = ptb.Scene.from_triangles(np.asarray(tris, np.float32).reshape(-1, 32), None)
    env = assets.make_env(3, 64, 32)
    sc.set_env_pixels(np.concatenate([env, np.ones_like(env[..., :1])], -1))
    return sc


def _one_tri(y=0.0, s=3.0):
    t = np.zeros(32, np.float32)
    t[0:12] = [-s, y, -s, 0, -s, y, s, 0, s, y, -s, 0]
    t[12:24] = [0, 1, 0, 0] * 3
    return t


def _fuzz_mesh(rng, n):
    """n random triangles as TriangleData rows: clusters of different density, slivers, exact duplicates, zero-area
    triangles, axis-aligned (flat-box) triangles, coincident vertices across triangles and a few scene-sized ones."""
    c = (rng.random((n, 1, 3), dtype=np.float32) - 0.5) * np.float32([60, 6, 60])
    c[: n // 3] = c[rng.integers(0, max(n // 50, 1), n // 3)] + (rng.random((n // 3, 1, 3), dtype=np.float32) - 0.5) * 0.3   # dense clusters
    s = np.where(rng.random((n, 1, 1)) < 0.05, 8.0, np.where(rng.random((n, 1, 1)) < 0.3, 0.02, 0.6)).astype(np.float32)
    v = c + (rng.random((n, 3, 3), dtype=np.float32) - 0.5) * s
    k = max(n // 20, 1)
    v[rng.integers(0, n, k)] = v[rng.integers(0, n, k)]            # duplicates (ties: lowest primitive id wins)
    i = rng.integers(0, n, k); v[i, 2] = v[i, 1]                   # zero area
    i = rng.integers(0, n, k); v[i, :, 1] = v[i, :1, 1]            # flat in y
    i = rng.integers(0, n, k); v[i, 0] = v[(i + 1) % n, 1]         # shared vertices
    if n >= 8:
        v[0] = [[-300, -3, -300], [300, -3, -300], [300, -3, 300]]; v[1] = [[-300, -3, -300], [300, -3, 300], [-300, -3, 300]]
    t = np.zeros((n, 32), np.float32)
    t[:, 0:12].reshape(n, 3, 4)[:, :, :3] = v
    t[:, 12:24].reshape(n, 3, 4)[:, :, 1] = 1.0
    return t


@pytest.mark.parametrize("n,seed", [(5, 1), (33, 2), (257, 3), (1000, 4), (6000, 5), (50000, 6)])
def test_bvh_fuzz_random_meshes(ptb, ctx, oh, assets, n, seed):
    """Builder and traversals on meshes that are NOT the five named scenes (compute-sanitizer is closed on this GPU pool,
    profiles/r2_sanitizer_refused.txt; this is the substitute evidence for the builder's atomics and arrival flags):
    every build variant gives a structurally valid tree, REPEATED builds of the same input give the same hits and the same
    tree statistics (a race in k_refit / k_refine_treelets would show as run-to-run differences), all tree widths agree bit
    for bit, and the hits are the brute-force loop's."""
    if PIPELINE != 3:
        pytest.skip("the BVH does not depend on the render pipeline")
    rng = np.random.default_rng(seed)
    tris = _fuzz_mesh(rng, n)
    sc = _tri_scene(ptb, assets, tris)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    v = tris[:, 0:12].reshape(n, 3, 4)[:, :, :3].reshape(-1, 3)
    lo, hi = np.percentile(v, 2, axis=0), np.percentile(v, 98, axis=0)
    o, d = random_rays(rng, 6000, lo, hi)
    d[::9, 2] = 0.0
    o[1::13] = v[rng.integers(0, len(v), len(o[1::13]))]          # rays that start on a vertex
    ref, ref_stats = None, {}
    variants = [dict(), dict(), dict(sah_refine=0), dict(bvh_width=4), dict(bvh_width=8), dict(bvh_width=8), dict(max_leaf_size=1),
                dict(morton_bits=63, treelet_size=16), dict(bvh_width=8, sah_refine=0, max_leaf_size=2)]
    for kw in variants:
        handle, st = ctx.accel_build(sc, ptb.default_build_cfg(**kw))
        assert st.num_triangles == n and st.max_depth < 128
        key = tuple(sorted(kw.items()))
        stats = (st.num_nodes, st.num_leaves, st.max_depth, st.num_nodes8)   # (sah_cost is a float atomic sum: order-dependent in the last bits)
        assert ref_stats.setdefault(key, stats) == stats, f"two builds of the same input differ: {kw}"
        nodes, tris_dev = ctx.accel_read(handle)
        import hashlib
        digest = hashlib.sha1(nodes[:, :14].tobytes() + tris_dev.tobytes()).hexdigest()   # boxes, child codes, leaf order
        assert ref_stats.setdefault(key + ("tree",), digest) == digest, f"two builds of the same input give different trees: {kw}"
        leaf_cap = min(kw.get("max_leaf_size", 4), 3) if kw.get("bvh_width") == 8 else kw.get("max_leaf_size", 4)
        assert _check_bvh(nodes, tris_dev, n, leaf_cap) == st.num_nodes, kw
        if kw.get("bvh_width") == 8:
            assert st.bvh_width == 8
            nodes8, tris8 = ctx.accel_read8(handle)
            assert _check_bvh8(nodes8, tris8, n)[0] == st.num_nodes8, kw
        got = ctx.trace_rays(handle, o, d)
        if ref is None:
            ref = got
            step = max(1, len(o) * n // 4_000_000)   # brute force in the oracle: bounded work
            for i in range(0, len(o), step):
                rp, rt, rb1, rb2 = oh.closest_hit("oracle", osc, o[i], d[i], use_bvh=0)
                assert rp == got[0][i], (i, rp, got[0][i])
                if rp >= 0:
                    assert np.float32(rt) == got[1][i] and np.float32(rb1) == got[2][i] and np.float32(rb2) == got[3][i]
            assert (got[0] >= 0).mean() > (0.05 if n >= 200 else 0.0)
        else:
            for a, b in zip(ref, got):
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), kw


@pytest.mark.parametrize("n_tris", [0, 1, 2, 3])
def test_degenerate_scenes(ptb, ctx, oh, assets, n_tris):
    """Empty, 1-, 2- and 3-triangle scenes (the builder's n < 2 path has no Karras tree), odd frame sizes."""
    tris = [_one_tri(0.1 * k, 3.0 - 0.5 * k) for k in range(n_tris)]
    sc = _tri_scene(ptb, assets, tris if tris else np.zeros((0, 32), np.float32))
    handle, st = ctx.accel_build(sc)
    assert st.num_triangles == n_tris
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    for (W, H) in ((1, 1), (37, 23)):
        kw = dict(spp_per_launch=2, max_depth=3)
        ga, gf, gh, gst = _render_gpu(ptb, ctx, handle, W, H, kw)
        ca, cf, ch, cseg = _render_cpu(oh, ptb, osc, W, H, kw)
        assert np.array_equal(gh, ch) and gst[0].segments == cseg
        assert np.array_equal(ga.view(np.uint32), ca.view(np.uint32))
    if n_tris == 0:
        assert np.all(gh == -1)


def test_depth_zero_and_single_sample(ptb, ctx, oh, assets):
    """max_depth = 0: the first hit already sets done (optixSphere.cu:738); spp = 1."""
    sc = load_config(ptb, assets, "c1", small=True)
    handle, _ = ctx.accel_build(sc)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    for kw in (dict(spp_per_launch=1, max_depth=0), dict(spp_per_launch=1, max_depth=1)):
        ga, gf, gh, gst = _render_gpu(ptb, ctx, handle, 64, 48, kw, dof=False)
        ca, cf, ch, cseg = _render_cpu(oh, ptb, osc, 64, 48, kw, dof=False)
        assert gst[0].segments == cseg and np.array_equal(gh, ch)
        assert np.array_equal(ga.view(np.uint32), ca.view(np.uint32))


def test_launch_argument_errors(ptb, ctx, assets):
    sc = load_config(ptb, assets, "c1", small=True)
    handle, _ = ctx.accel_build(sc)
    d = ctx.alloc(64 * 64 * 16)
    try:
        p = ptb.make_params(64, 64)
        p.accum_buffer, p.frame_buffer, p.handle = d, None, handle
        with pytest.raises(ptb.PtbError):  # frame buffer missing while write_frame = 1
            ctx.launch(p, ptb.default_render_cfg())
        p.handle = 987654
        with pytest.raises(ptb.PtbError):  # unknown acceleration structure
            ctx.launch(p, ptb.default_render_cfg(write_frame=0))
        p.handle = handle
        with pytest.raises(ptb.PtbError):  # 0 = reference estimator, 1/2 = the optional linear modes
            ctx.launch(p, ptb.default_render_cfg(write_frame=0, env_importance_sampling=7))
        with pytest.raises(ptb.PtbError):
            ctx.launch(p, ptb.default_render_cfg(write_frame=0, spp_per_launch=0))
        ctx.launch(p, ptb.default_render_cfg(write_frame=0, spp_per_launch=1, max_depth=2))  # and a valid one still works
        assert ctx.launch_stats().segments > 0
    finally:
        ctx.free(d)
    bare = ptb.Scene.from_triangles(_one_tri().reshape(1, 32), None)
    with pytest.raises(ptb.PtbError):  # no environment map
        ctx.accel_build(bare)


def test_resolve_peers_single_device(ptb, ctx):
    """ptb_resolve_peers with three 'ranks' living on one device: fixed-order sum, scale, tonemap, slice handling."""
    from szakdolgozat_pathtracer_b200 import parallel
    rng = np.random.default_rng(5)
    n = 1000
    accs = [(rng.random((n, 4), dtype=np.float32) * 3).astype(np.float32) for _ in range(3)]
    ptrs = [ctx.alloc(n * 16) for _ in range(3)]
    d_out, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
    try:
        for pp, a in zip(ptrs, accs):
            ctx.to_device(pp, a)
        ctx.memset(d_out, 0, n * 16); ctx.memset(d_frame, 0, n * 4)
        for r in range(3):  # three slices, as three ranks would do it
            first, cnt = parallel.pixel_slice_for_rank(r, 3, n)
            ctx.resolve_peers(ptrs, d_out, d_frame, first, cnt, 1.0 / 3.0)
        ctx.synchronize()
        out = ctx.to_host(d_out, (n, 4), np.float32)
        frame = ctx.to_host(d_frame, (n, 4), np.uint8)
        ctx.resolve(d_out, 0, d_frame, n, 1.0)  # tonemap of the same values through the single-GPU entry point
        ctx.synchronize()
        frame2 = ctx.to_host(d_frame, (n, 4), np.uint8)
    finally:
        for pp in ptrs + [d_out, d_frame]:
            ctx.free(pp)
    want = ((accs[0][:, :3] + accs[1][:, :3]) + accs[2][:, :3]) * np.float32(1.0 / 3.0)
    assert np.array_equal(out[:, :3], want) and np.all(out[:, 3] == 1.0)
    assert np.array_equal(frame, frame2) and frame[:, 3].min() == 255


def test_c3_full_pbr_parity_crop(ptb, ctx, oh, assets):
    """BASELINE config 3 scene: suitcase.obj with all four maps (albedo, normal, roughness, metallic, 2048^2 each) under a
    4096x2048 environment; close camera so that the mesh fills the crop: bit-exact accum and hit IDs."""
    if PIPELINE not in (3, 4):
        pytest.skip("the fused pipelines are enough for the large-texture scene (the others are covered on C1/C2)")
    sc = load_config(ptb, assets, "c3")
    m = sc.material(0)
    assert m.has_albedo and m.has_normal and m.has_roughness and m.has_metallic and m.albedo_w == 2048
    handle, st = ctx.accel_build(sc)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    kw = dict(spp_per_launch=4, max_depth=6)
    W, H = 192, 108
    ga, gf, gh, gst = _render_gpu(ptb, ctx, handle, W, H, kw, camera="suitcase_close")
    ca, cf, ch, cseg = _render_cpu(oh, ptb, osc, W, H, kw, camera="suitcase_close")
    assert np.array_equal(gh, ch) and gst[0].segments == cseg
    bad = (ga.view(np.uint32) != ca.view(np.uint32)).any(axis=2)
    assert bad.sum() == 0, f"{bad.sum()} of {W * H} accum pixels differ"
    assert (gh < 2204).mean() > 0.2  # the textured mesh covers a good part of the frame
    assert np.array_equal(gf, cf)


def test_c2_full_frame_primary_hits(ptb, ctx, oh, assets):
    """BASELINE config 2 at its full size: all 2 073 600 primary-hit triangle IDs (DoF on) and the 1-sample image, bit-exact."""
    if PIPELINE not in (3, 4):
        pytest.skip("full-frame check on the fused pipelines (3 = default, 2048-slot chunks at this size; 4 = persistent pool)")
    sc = load_config(ptb, assets, "c2")
    handle, _ = ctx.accel_build(sc)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    kw = dict(spp_per_launch=1, max_depth=2)
    ga, gf, gh, gst = _render_gpu(ptb, ctx, handle, 1920, 1080, kw)
    ca, cf, ch, cseg = _render_cpu(oh, ptb, osc, 1920, 1080, kw)
    mism = int((gh != ch).sum())
    assert mism == 0, f"{mism} of 2073600 primary hits differ (edge ties would show up here)"
    assert gst[0].segments == cseg
    assert np.array_equal(ga.view(np.uint32), ca.view(np.uint32))
    assert 0.02 < (gh < 15744).mean() < 0.2 and (gh == -1).mean() > 0.1


def test_row_bands_tile_bit_identically(ptb, ctx, assets):
    """Tile partitioning: three row bands rendered by separate launches == the whole frame (accum, frame, hit IDs)."""
    from szakdolgozat_pathtracer_b200 import parallel
    sc = load_config(ptb, assets, "c1", small=True)
    handle, _ = ctx.accel_build(sc)
    W, H = 150, 100
    kw = dict(spp_per_launch=3, max_depth=5)
    full_a, full_f, full_h, full_st = _render_gpu(ptb, ctx, handle, W, H, kw, subframes=2)
    n = W * H
    d_accum, d_frame, d_hits = ctx.alloc(n * 16), ctx.alloc(n * 4), ctx.alloc(n * 4)
    try:
        ctx.memset(d_accum, 0, n * 16); ctx.memset(d_frame, 0, n * 4); ctx.memset(d_hits, 0xFF, n * 4)
        seg = 0
        for r in (2, 0, 1):  # any order
            r0, r1 = parallel.row_band_for_rank(r, 3, H)
            for sf in range(2):
                p = ptb.make_params(W, H, subframe_index=sf, dof=True)
                p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
                ctx.launch(p, ptb.default_render_cfg(row_begin=r0, row_end=r1, pipeline=PIPELINE, aux_primary_hit=d_hits if sf == 0 else None, **kw))
                seg += ctx.launch_stats().segments
        a = ctx.to_host(d_accum, (H, W, 4), np.float32); f = ctx.to_host(d_frame, (H, W, 4), np.uint8); h = ctx.to_host(d_hits, (H, W), np.int32)
        p = ptb.make_params(W, H)
        p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
        with pytest.raises(ptb.PtbError):
            ctx.launch(p, ptb.default_render_cfg(row_begin=10, row_end=H + 1, **kw))
    finally:
        for b in (d_accum, d_frame, d_hits):
            ctx.free(b)
    assert seg == sum(s.segments for s in full_st)
    assert np.array_equal(a.view(np.uint32), full_a.view(np.uint32)) and np.array_equal(f, full_f) and np.array_equal(h, full_h)


def test_interleaved_strips_tile_bit_identically(ptb, ctx, assets):
    """Load-balanced tile partitioning: strips of 7 rows dealt round-robin to 3 'ranks' (ragged: 100 rows = 14 strips + 2
    rows, so one rank's last strip is padded) == the whole frame, bit for bit."""
    if PIPELINE == 1:
        pytest.skip("row interleave is a feature of the chunked pipelines")
    sc = load_config(ptb, assets, "c1", small=True)
    handle, _ = ctx.accel_build(sc)
    W, H = 90, 100
    kw = dict(spp_per_launch=2, max_depth=5, pipeline=PIPELINE)
    full_a, full_f, full_h, full_st = _render_gpu(ptb, ctx, handle, W, H, dict(spp_per_launch=2, max_depth=5))
    n = W * H
    d_accum, d_frame, d_hits = ctx.alloc(n * 16), ctx.alloc(n * 4), ctx.alloc(n * 4)
    try:
        ctx.memset(d_accum, 0, n * 16); ctx.memset(d_frame, 0, n * 4); ctx.memset(d_hits, 0xFF, n * 4)
        seg = 0
        for r in range(3):
            p = ptb.make_params(W, H, subframe_index=0, dof=True)
            p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
            ctx.launch(p, ptb.default_render_cfg(row_interleave_count=3, row_interleave_index=r, row_interleave_height=7, aux_primary_hit=d_hits, **kw))
            seg += ctx.launch_stats().segments
        a = ctx.to_host(d_accum, (H, W, 4), np.float32); f = ctx.to_host(d_frame, (H, W, 4), np.uint8); h = ctx.to_host(d_hits, (H, W), np.int32)
        # more ranks than strips: a no-op, not an error
        p = ptb.make_params(W, H)
        p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
        ctx.launch(p, ptb.default_render_cfg(row_interleave_count=64, row_interleave_index=63, row_interleave_height=50, **kw))
    finally:
        for b in (d_accum, d_frame, d_hits):
            ctx.free(b)
    assert seg == full_st[0].segments
    assert np.array_equal(a.view(np.uint32), full_a.view(np.uint32)) and np.array_equal(f, full_f) and np.array_equal(h, full_h)
