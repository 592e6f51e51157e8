"""ptb_multi (csrc/multi.cpp): multi-GPU rendering behind the C ABI, one host process driving n devices.

On a one-GPU box the same device is listed several times (several contexts, streams and accumulators on one GPU), which
runs exactly the code path of n real devices except the NVLink hops; tools/check_multi_gpu.py runs the same checks over
real peers and its output is kept under profiles/.

  tiles    every pixel is computed whole by one device with its full-frame seed -> BIT-IDENTICAL to one ptb_launch;
  samples  per-device sums of launch means, reduced in device order and folded into the running mean as
           (old * s + sum) / (s + K): equal to the reference's lerp (optixSphere.cu:403-409) up to rounding.  Stated
           tolerance: 2e-6 relative to the pixel value + 1e-7 absolute (a handful of float roundings); frame bytes within 1.
"""
import numpy as np
import pytest

from scenes import load_config

pytestmark = pytest.mark.gpu


def _single(ptb, ctx, sc, W, H, kw, launches):
    handle, _ = ctx.accel_build(sc)
    n = W * H
    d_accum, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
    try:
        ctx.memset(d_accum, 0, n * 16)
        seg, sf = 0, 0
        for k in launches:
            p = ptb.make_params(W, H, subframe_index=sf, dof=True)
            p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
            ctx.launch(p, ptb.default_render_cfg(subframes_per_launch=k, **kw))
            seg += ctx.launch_stats().segments
            sf += k
        return ctx.to_host(d_accum, (H, W, 4), np.float32), ctx.to_host(d_frame, (H, W, 4), np.uint8), seg
    finally:
        ctx.free(d_accum); ctx.free(d_frame)


def _multi(ptb, sc, n_dev, W, H, kw, launches, split):
    m = ptb.Multi([0] * n_dev)
    try:
        m.accel_build(sc)
        root = m.root
        n = W * H
        d_accum, d_frame = root.alloc(n * 16), root.alloc(n * 4)
        root.memset(d_accum, 0, n * 16); root.memset(d_frame, 0, n * 4)
        root.synchronize()
        sf = 0
        for k in launches:
            p = ptb.make_params(W, H, subframe_index=sf, dof=True)
            p.accum_buffer, p.frame_buffer = d_accum, d_frame
            m.launch(p, ptb.default_render_cfg(subframes_per_launch=k, **kw), split)
            sf += k
        seg = m.totals()["segments"]
        a, f = root.to_host(d_accum, (H, W, 4), np.float32), root.to_host(d_frame, (H, W, 4), np.uint8)
        root.free(d_accum); root.free(d_frame)
        return a, f, seg
    finally:
        m.close()


@pytest.mark.parametrize("n_dev", [2, 3])
def test_multi_tiles_bit_identical(ptb, ctx, assets, n_dev):
    sc = load_config(ptb, assets, "c1", small=True)
    W, H = 150, 100   # 100 rows = 6 strips of 16 + 4: ragged last strip
    kw = dict(spp_per_launch=3, max_depth=5)
    launches = [2, 1, 3]
    a1, f1, s1 = _single(ptb, ctx, sc, W, H, kw, launches)
    a2, f2, s2 = _multi(ptb, sc, n_dev, W, H, kw, launches, ptb.PTB_SPLIT_TILES)
    assert s1 == s2
    assert np.array_equal(a1.view(np.uint32), a2.view(np.uint32)) and np.array_equal(f1, f2)


@pytest.mark.parametrize("n_dev", [1, 2, 3])
def test_multi_sample_split_matches_running_average(ptb, ctx, assets, n_dev):
    sc = load_config(ptb, assets, "c2")
    W, H = 160, 90
    kw = dict(spp_per_launch=4, max_depth=6)
    launches = [4, 5, 1]   # K not a multiple of the device count; a later launch continues the frame's mean
    a1, f1, s1 = _single(ptb, ctx, sc, W, H, kw, launches)
    a2, f2, s2 = _multi(ptb, sc, n_dev, W, H, kw, launches, ptb.PTB_SPLIT_SAMPLES)
    assert s1 == s2, "the same samples are rendered, whatever the device count"
    if n_dev == 1:
        assert np.array_equal(a1.view(np.uint32), a2.view(np.uint32)) and np.array_equal(f1, f2)
        return
    err = np.abs(a1[..., :3].astype(np.float64) - a2[..., :3])
    tol = 2e-6 * np.abs(a1[..., :3]) + 1e-7
    assert (err <= tol).all(), float((err / (np.abs(a1[..., :3]) + 1e-7)).max())
    assert np.all(a2[..., 3] == 1.0)
    assert np.abs(f1.astype(np.int32) - f2.astype(np.int32)).max() <= 1


def test_multi_argument_errors(ptb, assets):
    with pytest.raises(ptb.PtbError):
        ptb.Multi([0, 99])          # no such device
    m = ptb.Multi([0, 0])
    try:
        p = ptb.make_params(32, 32)
        d = m.root.alloc(32 * 32 * 16)
        p.accum_buffer = d
        with pytest.raises(ptb.PtbError):   # nothing built yet
            m.launch(p, ptb.default_render_cfg(write_frame=0))
        sc = load_config(ptb, assets, "c1", small=True)
        m.accel_build(sc)
        with pytest.raises(ptb.PtbError):   # the row partition belongs to the tile split
            m.launch(p, ptb.default_render_cfg(write_frame=0, row_begin=0, row_end=8), ptb.PTB_SPLIT_TILES)
        m.launch(p, ptb.default_render_cfg(write_frame=0, row_begin=0, row_end=8, spp_per_launch=1, max_depth=1))   # a crop under the sample split is fine
        with pytest.raises(ptb.PtbError):
            m.launch(p, ptb.default_render_cfg(write_frame=0), 7)
        m.launch(p, ptb.default_render_cfg(write_frame=0, spp_per_launch=1, max_depth=1))
        m.synchronize()
        m.root.free(d)
    finally:
        m.close()


def test_peer_flags_order_the_exchange_without_nccl(ptb, ctx):
    """ptb_peer_signal / ptb_resolve_peers_sync / ptb_peer_wait with three 'ranks' on one device, each on its own stream: the
    resolve kernels wait on the device for all signals, the root's wait sees all three slices; same result as the plain
    kernel.  Then a wait for an epoch nobody signals: it must give up (error word set) instead of hanging."""
    import torch
    from szakdolgozat_pathtracer_b200 import parallel
    rng = np.random.default_rng(11)
    n, world = 5000, 3
    accs = [(rng.random((n, 4), dtype=np.float32) * 2).astype(np.float32) for _ in range(world)]
    ptrs = [ctx.alloc(n * 16) for _ in range(world)]
    flags = [ctx.peer_flags_create() for _ in range(world)]
    d_out, d_frame, d_out2, d_frame2 = ctx.alloc(n * 16), ctx.alloc(n * 4), ctx.alloc(n * 16), ctx.alloc(n * 4)
    streams = [torch.cuda.Stream() for _ in range(world)]
    try:
        for pp, a in zip(ptrs, accs):
            ctx.to_device(pp, a)
        ctx.memset(d_out, 0, n * 16); ctx.memset(d_frame, 0, n * 4)
        ctx.synchronize()
        for epoch in (1, 2):   # twice: epochs only grow, nothing is reset in between
            for r in (2, 0, 1):   # resolves are enqueued BEFORE some of the signals they wait for
                first, cnt = parallel.pixel_slice_for_rank(r, world, n)
                if r == 2:
                    ctx.peer_signal(flags, r, 0, epoch, stream=streams[r].cuda_stream)
                    ctx.resolve_peers_sync(ptrs, r, flags[r], flags[0], epoch, d_out, d_frame, first, cnt, 1.0 / 3.0, stream=streams[r].cuda_stream)
                else:
                    ctx.resolve_peers_sync(ptrs, r, flags[r], flags[0], epoch, d_out, d_frame, first, cnt, 1.0 / 3.0, stream=streams[r].cuda_stream)
            for r in (0, 1):
                # a second stream per rank for the late signals (on the rank's own stream they would queue behind its waiting kernel)
                s2 = torch.cuda.Stream()
                ctx.peer_signal(flags, r, 0, epoch, stream=s2.cuda_stream)
            ctx.peer_wait(flags[0], 1, world, epoch, stream=streams[0].cuda_stream)
            streams[0].synchronize()
            out = ctx.to_host(d_out, (n, 4), np.float32)
            frame = ctx.to_host(d_frame, (n, 4), np.uint8)
            for r in range(world):
                first, cnt = parallel.pixel_slice_for_rank(r, world, n)
                ctx.resolve_peers(ptrs, d_out2, d_frame2, first, cnt, 1.0 / 3.0)
            ctx.synchronize()
            assert np.array_equal(out, ctx.to_host(d_out2, (n, 4), np.float32)) and np.array_equal(frame, ctx.to_host(d_frame2, (n, 4), np.uint8))
            assert not ctx.peer_flags_error(flags[0])
            ctx.memset(d_out, 0, n * 16)
            torch.cuda.synchronize()
        ctx.peer_wait(flags[1], 0, world, 99)   # nobody signals epoch 99
        assert ctx.peer_flags_error(flags[1]), "a wait without a signal must time out and set the error word"
    finally:
        torch.cuda.synchronize()
        for pp in ptrs + flags + [d_out, d_frame, d_out2, d_frame2]:
            ctx.free(pp)


def test_c5_convergence_band_sample_split_tolerance(ptb, ctx, oh, assets):
    """BASELINE config 5 in small: a band of the 4K frame of the C5 scene at 512 spp (8 subframes x 64), sample-split over 1, 2, 4
    and 8 contexts.  Stated tolerance of the multi-GPU sum order against the single-GPU running average (tools/c5_convergence.py
    runs the same check at 4096 spp over real GPUs, profiles/r2_c5_convergence.json: 2.3e-7 / 9e-7): relative RMSE <= 1e-6,
    maximum relative error <= 1e-5.  N = 1 equals the oracle bit for bit."""
    sc = load_config(ptb, assets, "c5")
    W, H, rows0, rows1 = 3840, 2160, 1240, 1248
    n = W * H
    cfg_kw = dict(spp_per_launch=64, max_depth=8, subframes_per_launch=8, row_begin=rows0, row_end=rows1, write_frame=0)
    accs = {}
    for n_dev in (1, 2, 4, 8):
        m = ptb.Multi([0] * n_dev)
        try:
            m.accel_build(sc)
            d_a = m.root.alloc(n * 16)
            m.root.memset(d_a, 0, n * 16); m.root.synchronize()
            p = ptb.make_params(W, H, subframe_index=0, dof=True)
            p.accum_buffer = d_a
            m.launch(p, ptb.default_render_cfg(**cfg_kw), ptb.PTB_SPLIT_SAMPLES)
            m.synchronize()
            accs[n_dev] = m.root.to_host(d_a, (H, W, 4), np.float32)[rows0:rows1].copy()
            m.root.free(d_a)
        finally:
            m.close()
    ref = accs[1][..., :3].astype(np.float64)
    for n_dev in (2, 4, 8):
        d = accs[n_dev][..., :3] - ref
        assert np.sqrt((d ** 2).mean()) / ref.mean() <= 1e-6, n_dev
        assert (np.abs(d) / (np.abs(ref) + 1e-6)).max() <= 1e-5, n_dev
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    win = (1900, 1242, 1932, 1246)
    ca = np.zeros((H, W, 4), np.float32)
    for sf in range(8):
        p = ptb.make_params(W, H, subframe_index=sf, dof=True)
        ca, _, _, _, rc = oh.render("oracle", osc, oh.params_from_ptb(p), oh.default_config("oracle", spp_per_launch=64, max_depth=8), accum=ca, window=win, want_hits=False)
        assert rc == 0
    g = accs[1][win[1] - rows0:win[3] - rows0, win[0]:win[2]]
    assert np.array_equal(g[..., :3].view(np.uint32), ca[win[1]:win[3], win[0]:win[2], :3].view(np.uint32))
