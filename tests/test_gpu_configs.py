"""GPU parity on the large BASELINE configurations, at their full frame sizes, against the CPU oracle on crops.

  C4  fish + tower + five seeded icosphere stand-ins (4.6 M triangles): the AUTOMATIC 4-wide traversal
      (ptb_build_cfg.bvh_width = 0 picks it from 1 M triangles on), a material table of eight entries drawn from
      std::mt19937(material_seed) (optixSphere.cpp:553-582) of which one is emissive (cu:725-731), 1920x1080.
  C5  0.33 M-triangle model with uvs + 2048^2 albedo map under the 4096x2048 environment, 3840x2160.
  one-row / 1x1 textures and environments (the texel wrap rule on images of height 1).

The assets are generated on the box from their seeds (tools/make_assets.py); the oracle renders only the crop windows.
"""
import numpy as np
import pytest

from scenes import CAMERAS, load_config

pytestmark = pytest.mark.gpu


def _launch_full(ptb, ctx, handle, W, H, kw, camera="default", pipeline=0):
    n = W * H
    d_accum, d_frame, d_hits = ctx.alloc(n * 16), ctx.alloc(n * 4), ctx.alloc(n * 4)
    try:
        ctx.memset(d_accum, 0, n * 16)
        ctx.memset(d_hits, 0xFF, n * 4)
        p = ptb.make_params(W, H, subframe_index=0, dof=True, **CAMERAS[camera])
        p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
        ctx.launch(p, ptb.default_render_cfg(aux_primary_hit=d_hits, pipeline=pipeline, **kw))
        st = ctx.launch_stats()
        return (ctx.to_host(d_accum, (H, W, 4), np.float32), ctx.to_host(d_frame, (H, W, 4), np.uint8),
                ctx.to_host(d_hits, (H, W), np.int32), st, p)
    finally:
        for b in (d_accum, d_frame, d_hits):
            ctx.free(b)


def _check_window(oh, osc, p, kw, win, ga, gf, gh):
    x0, y0, x1, y1 = win
    ca, cf, ch, cst, rc = oh.render("oracle", osc, oh.params_from_ptb(p), oh.default_config("oracle", **kw), window=win)
    assert rc == 0
    assert np.array_equal(gh[y0:y1, x0:x1], ch[y0:y1, x0:x1]), f"primary hits differ in {win}"
    bad = (ga[y0:y1, x0:x1].view(np.uint32) != ca[y0:y1, x0:x1].view(np.uint32)).any(axis=2)
    assert bad.sum() == 0, f"{bad.sum()} of {bad.size} accum pixels differ in window {win}"
    assert np.array_equal(gf[y0:y1, x0:x1], cf[y0:y1, x0:x1]), f"frame bytes differ in {win}"
    return ch[y0:y1, x0:x1]


def test_c4_auto_wide_bvh_random_materials_parity(ptb, ctx, oh, assets):
    sc = load_config(ptb, assets, "c4", material_seed=4)
    assert sc.num_triangles >= 4_000_000 and sc.num_materials == 8
    emissive = [i for i in range(sc.num_materials) if max(sc.material(i).emission_color) > 1e-4]
    assert emissive == [4], emissive  # statue3, drawn emissive by the reference's rule (prob. 0.1) for material_seed 4
    handle, bst = ctx.accel_build(sc)  # default build cfg: bvh_width = 0 (auto)
    assert bst.bvh_width == 4, "a 4.6 M-triangle scene must pick the 4-wide traversal on its own"
    assert bst.num_triangles == sc.num_triangles and bst.max_depth < 128
    W, H = 1920, 1080
    kw = dict(spp_per_launch=4, max_depth=8)
    ga, gf, gh, st, p = _launch_full(ptb, ctx, handle, W, H, kw)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    mids = sc.material_ids()
    # window 1: the emissive statue, the tower in front of it, the fish and the floor; window 2: a diffuse statue
    h1 = _check_window(oh, osc, p, kw, (936, 660, 1064, 708), ga, gf, gh)
    seen = set(np.unique(mids[h1[h1 >= 0]]).tolist())
    assert 4 in seen and len(seen) >= 3, seen
    h2 = _check_window(oh, osc, p, kw, (620, 650, 684, 690), ga, gf, gh)
    assert (h2 >= 0).mean() > 0.9 and 3 in set(np.unique(mids[h2[h2 >= 0]]).tolist())
    # the forced 2-wide tree gives the same image (hit rule independent of the tree)
    handle2, bst2 = ctx.accel_build(sc, ptb.default_build_cfg(bvh_width=2))
    assert bst2.bvh_width == 2
    ga2, gf2, gh2, st2, _ = _launch_full(ptb, ctx, handle2, W, H, kw)
    assert st2.segments == st.segments and np.array_equal(gh, gh2) and np.array_equal(gf, gf2)
    assert np.array_equal(ga.view(np.uint32), ga2.view(np.uint32))
    sc.close()


def test_c5_4k_parity_crop(ptb, ctx, oh, assets):
    sc = load_config(ptb, assets, "c5")
    m = sc.material(0)
    assert m.has_albedo and m.albedo_w == 2048 and sc.num_triangles == 327680 + 2
    handle, bst = ctx.accel_build(sc)
    assert bst.bvh_width == 2  # below the 1 M-triangle threshold
    W, H = 3840, 2160
    kw = dict(spp_per_launch=4, max_depth=8)
    ga, gf, gh, st, p = _launch_full(ptb, ctx, handle, W, H, kw)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    h1 = _check_window(oh, osc, p, kw, (1880, 1230, 1960, 1270), ga, gf, gh)   # on the textured model
    assert (h1 < 327680).mean() > 0.9
    h2 = _check_window(oh, osc, p, kw, (2400, 1490, 2464, 1522), ga, gf, gh)   # silhouette: model, floor, reflections
    assert len(np.unique(h2)) > 1
    assert st.paths == W * H * 4
    sc.close()


def _quad_scene(ptb, y=0.0, s=4.0):
    t = np.zeros((2, 32), np.float32)
    t[0, 0:12] = [-s, y, -s, 0, -s, y, s, 0, s, y, s, 0]
    t[1, 0:12] = [-s, y, -s, 0, s, y, s, 0, s, y, -s, 0]
    t[:, 12:24] = [0, 1, 0, 0] * 3
    t[0, 24:30] = [0, 0, 0, 3, 3, 3]      # uvs run over three repeats of the map
    t[1, 24:30] = [0, 0, 3, 3, 3, 0]
    return ptb.Scene.from_triangles(t, np.zeros(2, np.uint32))


@pytest.mark.parametrize("tw,th,ew,eh", [(1, 1, 1, 1), (2, 1, 4, 1), (1, 3, 1, 2), (5, 1, 7, 1)])
def test_one_row_textures_and_environments(ptb, ctx, oh, tw, th, ew, eh):
    """Bilinear taps at x0 = y0 = -1 on images with one row or one column: must stay inside the image (kernel and oracle
    wrap the linear index the same way) -- bit-exact accum, no NaN, no fault."""
    rng = np.random.default_rng(tw * 100 + th * 10 + ew)
    sc = _quad_scene(ptb)
    tex8 = (rng.integers(0, 256, (th, tw, 4)).astype(np.float32) / np.float32(255.0)).astype(np.float32)
    texf = rng.random((th, tw, 4), dtype=np.float32)
    sc.set_materials([dict(diffuse_color=(0.5, 0.5, 0.5), roughness=0.5, albedo=tex8, roughness_map=texf, normal_map=tex8, metallic_map=texf)])
    sc.set_env_pixels((rng.random((eh, ew, 4), dtype=np.float32) * 2).astype(np.float32))
    handle, _ = ctx.accel_build(sc)
    W, H = 96, 64
    kw = dict(spp_per_launch=4, max_depth=5)
    for pipeline in (1, 3):
        ga, gf, gh, st, p = _launch_full(ptb, ctx, handle, W, H, kw, pipeline=pipeline)
        assert np.isfinite(ga).all()
        osc = oh.OracleScene.from_ptb(sc, guard=True)
        ca, cf, ch, cst, rc = oh.render("oracle", osc, oh.params_from_ptb(p), oh.default_config("oracle", **kw))
        assert rc == 0 and st.segments == cst.segments and np.array_equal(gh, ch)
        assert np.array_equal(ga.view(np.uint32), ca.view(np.uint32)) and np.array_equal(gf, cf)
    assert (gh >= 0).mean() > 0.3


def test_stale_handle_is_refused(ptb, ctx, assets):
    """a handle built before the scene was modified must not silently render the old upload."""
    sc = _quad_scene(ptb)
    sc.set_env_pixels(np.ones((2, 4, 4), np.float32))
    handle, _ = ctx.accel_build(sc)
    n = 16 * 16
    d = ctx.alloc(n * 16)
    try:
        p = ptb.make_params(16, 16)
        p.accum_buffer, p.handle = d, handle
        cfg = ptb.default_render_cfg(write_frame=0, spp_per_launch=1, max_depth=1)
        ctx.launch(p, cfg)
        sc.set_env_pixels(np.zeros((2, 4, 4), np.float32))
        with pytest.raises(ptb.PtbError):
            ctx.launch(p, cfg)
        handle2, _ = ctx.accel_build(sc)
        p.handle = handle2
        ctx.launch(p, cfg)
        ctx.synchronize()
    finally:
        ctx.free(d)


def test_bounded_pool_batches_bit_identical(ptb, ctx, assets):
    """ptb_render_cfg.max_pool_bytes: a 7-subframe launch whose pool is capped at 2 / 3 subframes' worth of path state runs as
    batches inside ptb_launch -- accumulator, frame, primary hits and every counter equal the single wavefront's."""
    sc = load_config(ptb, assets, "c2")
    handle, _ = ctx.accel_build(sc)
    W, H = 200, 120
    n = W * H
    res = []
    for cap in (0, 97 * n * 2 + 5, 97 * n * 3, 1):   # default (2 GiB: one wavefront), 2 per batch, 3 per batch, 1 per batch
        d_accum, d_frame, d_hits = ctx.alloc(n * 16), ctx.alloc(n * 4), ctx.alloc(n * 4)
        try:
            ctx.memset(d_accum, 0, n * 16); ctx.memset(d_hits, 0xFF, n * 4)
            p = ptb.make_params(W, H, subframe_index=3, dof=True, **CAMERAS["monkey_close"])
            p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
            ctx.totals(reset=True)
            ctx.launch(p, ptb.default_render_cfg(spp_per_launch=3, max_depth=6, subframes_per_launch=7, max_pool_bytes=cap, aux_primary_hit=d_hits,
                                                 count_traversal=1))
            st = ctx.launch_stats()
            tot = ctx.totals(reset=True)
            assert tot["segments"] == st.segments and tot["hits"] == st.hits and tot["launches"] == 1, (tot, st.segments)   # folded once
            res.append((ctx.to_host(d_accum, (H, W, 4), np.float32).view(np.uint32), ctx.to_host(d_frame, (H, W, 4), np.uint8),
                        ctx.to_host(d_hits, (H, W), np.int32), (st.segments, st.hits, st.misses, st.paths, st.nodes_visited, st.tris_tested)))
        finally:
            for b in (d_accum, d_frame, d_hits):
                ctx.free(b)
    for r in res[1:]:
        assert all(np.array_equal(x, y) for x, y in zip(res[0][:3], r[:3])) and res[0][3] == r[3]
    assert res[0][3][3] == n * 3 * 7 and res[0][3][0] > res[0][3][3]


@pytest.mark.parametrize("arith,batch", [(0, 1), (1, 1), (0, 3)], ids=["exact", "fast", "exact-batched"])
def test_overlapped_launches_equal_serial_launches(ptb, ctx, assets, arith, batch):
    """ptb_render_cfg.overlap_lanes: the reference's render loop (one subframe per launch, running average, tonemapped frame,
    optixSphere.cpp:1390-1437) with launches overlapping on the context's internal lanes (0 = automatic, 2, 4) against strictly
    serial launches (1): accumulation buffer, frame, per-launch statistics and context totals identical; also when the caller
    clears or reads the buffers between launches on its stream."""
    from scenes import CAMERAS
    sc = load_config(ptb, assets, "c2")
    handle, _ = ctx.accel_build(sc)
    W, H, N = 320, 200, 9
    n = W * H
    results = []
    for lanes in (1, 0, 2, 4):
        d_accum, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
        try:
            ctx.memset(d_accum, 0, n * 16)
            ctx.totals(reset=True)
            cfg = ptb.default_render_cfg(spp_per_launch=3, max_depth=6, arith_mode=arith, overlap_lanes=lanes, subframes_per_launch=batch)
            seg, mids = [], []
            for k in range(N):
                p = ptb.make_params(W, H, subframe_index=k * batch, dof=True, **CAMERAS["monkey_close"])
                p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
                ctx.launch(p, cfg)
                if k in (2, 5):
                    st = ctx.launch_stats()      # synchronises: statistics of THIS launch, whatever lane it ran on
                    seg.append((st.segments, st.hits, st.misses, st.iterations))
                if k == 4:
                    mids.append(ctx.to_host(d_frame, (H, W, 4), np.uint8).copy())   # a read in the middle of the loop
            st = ctx.launch_stats()
            seg.append((st.segments, st.hits, st.misses, st.iterations))
            results.append((ctx.to_host(d_accum, (H, W, 4), np.float32).view(np.uint32), ctx.to_host(d_frame, (H, W, 4), np.uint8), seg, mids[0],
                            ctx.totals(reset=True)))
        finally:
            ctx.free(d_accum); ctx.free(d_frame)
    ref = results[0]
    assert ref[4]["segments"] > N * n and ref[4]["launches"] == N
    for got in results[1:]:
        assert np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1]) and np.array_equal(ref[3], got[3])
        assert ref[2] == got[2] and ref[4] == got[4]
    with pytest.raises(ptb.PtbError):
        p = ptb.make_params(W, H)
        d = ctx.alloc(n * 16)
        try:
            p.accum_buffer, p.handle = d, handle
            ctx.launch(p, ptb.default_render_cfg(write_frame=0, overlap_lanes=5))
        finally:
            ctx.free(d)


def test_glass_branch_parity(ptb, ctx, oh, assets):
    """HitGroupData.transparent (optixSphere.cu:803-856): exact build bit-identical to the oracle (which is pinned to the
    reference's own code by tests/golden/ref_glass_demo.npz); the fast build renders the same picture within noise."""
    from scenes import glass_demo_scene
    sc = glass_demo_scene(ptb, assets)
    handle, _ = ctx.accel_build(sc)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    W, H = 192, 128
    kw = dict(spp_per_launch=6, max_depth=12)
    ids = sc.material_ids()
    for pipeline in (1, 2, 3, 4):
        ga, gf, gh, st, p = _launch_full(ptb, ctx, handle, W, H, kw, pipeline=pipeline)
        if pipeline == 1:
            ca, cf, ch, cst, rc = oh.render("oracle", osc, oh.params_from_ptb(p), oh.default_config("oracle", **kw))
            assert rc == 0
        assert st.segments == cst.segments and np.array_equal(gh, ch), pipeline
        assert np.array_equal(ga.view(np.uint32), ca.view(np.uint32)) and np.array_equal(gf, cf), pipeline
    assert (ids[gh[gh >= 0]] == 2).mean() > 0.03
    fa, ff, fh, fst, _ = _launch_full(ptb, ctx, handle, W, H, dict(arith_mode=ptb.PTB_ARITH_FAST, **kw))
    assert np.array_equal(fh, gh) and np.isfinite(fa).all()
    assert abs(fa[..., :3].mean() - ga[..., :3].mean()) < 0.05 * ga[..., :3].mean()
