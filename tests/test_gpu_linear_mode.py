"""Optional mode beyond the reference: environment importance sampling through the prebuilt CDF,
under the LINEAR estimator (env_importance_sampling = 1) and its BSDF-only twin (= 2).  Not comparable with the oracle
(the reference's estimator is non-linear); validated on its own: the sampler against its analytic density, and NEE+MIS
against brute-force BSDF sampling (same expectation, lower variance)."""
import numpy as np
import pytest

from scenes import load_config

pytestmark = pytest.mark.gpu


def test_env_cdf_sampler_matches_its_density(ptb, ctx, assets):
    sc = load_config(ptb, assets, "c1", small=True)
    handle, _ = ctx.accel_build(sc)
    env = sc.env()
    h, w = env.shape[:2]
    rng = np.random.default_rng(1)
    n = 400_000
    out = ctx.test_env_sample(handle, rng.random((n, 2), dtype=np.float32))
    d, pdf = out[:, :3], out[:, 3]
    assert np.allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-4) and np.all(pdf > 0)
    # E[1/pdf] = solid angle of the sphere
    assert abs(np.mean(1.0 / pdf.astype(np.float64)) - 4 * np.pi) < 0.03 * 4 * np.pi
    # expected density: (luminance + 1 % of its mean) * sin(theta), per texel
    lum = 0.2126 * env[..., 0] + 0.7152 * env[..., 1] + 0.0722 * env[..., 2]
    st = np.sin(np.pi * (np.arange(h) + 0.5) / h)[:, None]
    f = (np.maximum(lum, 0) + 0.01 * lum.mean()) * st
    p_tex = f / f.sum()
    u = 0.5 + np.arctan2(d[:, 2], d[:, 0]) / (2 * np.pi)
    v = 0.5 - np.arcsin(np.clip(d[:, 1], -1, 1)) / np.pi
    i = np.clip((u * w).astype(int), 0, w - 1); j = np.clip((v * h).astype(int), 0, h - 1)
    # the reported pdf is the texel probability per solid angle
    want_pdf = p_tex[j, i] * w * h / (2 * np.pi ** 2 * st[j, 0])
    assert np.allclose(pdf, want_pdf, rtol=2e-3)
    # sample counts over a coarse 8 x 16 grid of the map follow the texel probabilities
    gh, gw = 8, 16
    cnt = np.zeros((gh, gw)); np.add.at(cnt, (j * gh // h, i * gw // w), 1)
    exp = p_tex.reshape(gh, h // gh, gw, w // gw).sum(axis=(1, 3)) * n
    big = exp > 200
    assert np.all(np.abs(cnt[big] - exp[big]) < 6 * np.sqrt(exp[big]))
    # the sun (0.1 % of the texels) draws a large share of the samples
    sun = lum > 50
    assert sun.mean() < 0.01 and (lum[j, i] > 50).mean() > 10 * sun.mean()


def _render(ptb, ctx, handle, W, H, mode, first_subframe, spp, batch, depth=6):
    n = W * H
    d_accum, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
    try:
        ctx.memset(d_accum, 0, n * 16)
        p = ptb.make_params(W, H, subframe_index=first_subframe, dof=True)
        p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
        ctx.launch(p, ptb.default_render_cfg(spp_per_launch=spp, max_depth=depth, subframes_per_launch=batch, accumulate_mode=1,
                                            env_importance_sampling=mode))
        st = ctx.launch_stats()
        a = ctx.to_host(d_accum, (H, W, 4), np.float32)[..., :3] / batch
    finally:
        ctx.free(d_accum); ctx.free(d_frame)
    return a, st


def test_nee_mis_agrees_with_bsdf_sampling_and_has_less_noise(ptb, ctx, assets):
    sc = load_config(ptb, assets, "c1", small=True)   # sun of radiance 200 in a 0.1 % patch of the sky: hard for BSDF sampling
    handle, _ = ctx.accel_build(sc)
    W, H = 48, 32
    nee_a, st1 = _render(ptb, ctx, handle, W, H, 1, 0, 128, 16)
    nee_b, _ = _render(ptb, ctx, handle, W, H, 1, 16, 128, 16)
    bs_a, st2 = _render(ptb, ctx, handle, W, H, 2, 0, 128, 16)
    bs_b, _ = _render(ptb, ctx, handle, W, H, 2, 16, 128, 16)
    assert np.isfinite(nee_a).all() and np.isfinite(bs_a).all() and st1.segments > 0 and st2.segments > 0
    # same expectation: frame means agree within the Monte Carlo error of the noisier estimator
    m_nee, m_bs = 0.5 * (nee_a + nee_b).mean(), 0.5 * (bs_a + bs_b).mean()
    assert abs(m_nee - m_bs) < 0.05 * m_bs, (m_nee, m_bs)
    # block means too (6 x 4 blocks of 8 x 8 pixels), where the BSDF-only image has enough samples
    blk = lambda a: a.reshape(4, 8, 6, 8, 3).mean(axis=(1, 3))
    rel = np.abs(blk(0.5 * (nee_a + nee_b)) - blk(0.5 * (bs_a + bs_b))) / np.maximum(blk(0.5 * (bs_a + bs_b)), 1e-3)
    assert np.median(rel) < 0.08, np.median(rel)
    # variance: two independent halves differ much less with light sampling
    noise_nee = np.mean((nee_a - nee_b) ** 2)
    noise_bs = np.mean((bs_a - bs_b) ** 2)
    assert noise_nee < 0.5 * noise_bs, (noise_nee, noise_bs)


def test_linear_mode_does_not_disturb_the_reference_path(ptb, ctx, oh, assets):
    """A linear-mode launch in between must leave the parity-checked integrator bit-exact."""
    sc = load_config(ptb, assets, "c1", small=True)
    handle, _ = ctx.accel_build(sc)
    _render(ptb, ctx, handle, 40, 24, 1, 0, 4, 2)
    n = 64 * 40
    d_accum, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
    try:
        ctx.memset(d_accum, 0, n * 16)
        p = ptb.make_params(64, 40, subframe_index=0, dof=True)
        p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
        ctx.launch(p, ptb.default_render_cfg(spp_per_launch=3, max_depth=5))
        ga = ctx.to_host(d_accum, (40, 64, 4), np.float32)
    finally:
        ctx.free(d_accum); ctx.free(d_frame)
    osc = oh.OracleScene.from_ptb(sc, guard=False)
    ca, _, _, _, rc = oh.render("oracle", osc, oh.params_from_ptb(p), oh.default_config("oracle", spp_per_launch=3, max_depth=5))
    assert rc == 0 and np.array_equal(ga.view(np.uint32), ca.view(np.uint32))
