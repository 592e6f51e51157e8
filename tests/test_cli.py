"""ptb_render: the reference executable's non-interactive path (optixSphere.cpp:754-791, 1443-1496) over the C ABI."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
CLI = ROOT / "szakdolgozat_pathtracer_b200" / "ptb_render"


def _run(*args):
    return subprocess.run([str(CLI), *map(str, args)], capture_output=True, text=True)


def test_usage_and_errors(ptb):
    assert CLI.exists(), "build() makes the CLI next to libptb.so"
    r = _run("--help")
    assert r.returncode == 1 and "--file | -f <filename>" in r.stderr and "--dim=<width>x<height>" in r.stderr  # printUsageAndExit exits 1
    r = _run("--bogus")
    assert r.returncode == 1 and "Unknown option '--bogus'" in r.stderr
    r = _run("--file")
    assert r.returncode == 1
    r = _run("--dim=0x7", "-f", "/tmp/x.png")
    assert r.returncode == 1 and "Invalid window dimensions" in r.stderr
    r = _run("--dim=64x64")  # no window system: --file is mandatory
    assert r.returncode == 1 and "--file" in r.stderr


def test_cli_without_gpu_fails_loudly(ptb, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = _run("-f", tmp_path / "x.png", "--scene", ROOT / "assets" / "test.obj", "--env", "missing.exr")
    assert r.returncode == 1 and "Caught exception" in r.stderr and "no CPU path" in r.stderr


@pytest.mark.gpu
def test_cli_render_matches_library(ptb, ctx, assets, tmp_path):
    from scenes import load_config
    cfg = assets.ensure("c1", small=True)
    out = tmp_path / "cli.png"
    r = _run("-f", out, "--dim=96x64", "-s", 3, "--depth", 5, "--launches", 2, "--scene", cfg["files"][0], "--env", cfg["env"], "--scale", cfg["scale"],
             "--no-gl-interop")
    assert r.returncode == 0, r.stderr
    assert "Loaded models with 14 triangles total." in r.stdout
    img = ptb.load_image_rgba8(out)
    sc = load_config(ptb, assets, "c1", small=True)
    handle, _ = ctx.accel_build(sc)
    n = 96 * 64
    d_accum, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
    try:
        for sf in range(2):
            p = ptb.make_params(96, 64, subframe_index=sf, dof=True)
            p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
            ctx.launch(p, ptb.default_render_cfg(spp_per_launch=3, max_depth=5))
        frame = ctx.to_host(d_frame, (64, 96, 4), np.uint8)
    finally:
        ctx.free(d_accum); ctx.free(d_frame)
    assert np.array_equal(img, frame[::-1])  # the file is written top row first


@pytest.mark.gpu
def test_cli_batch_accum_out_fast_and_gpus(ptb, ctx, assets, tmp_path):
    """--batch / --accum-out (ptb_save_accum_raw) / --fast / --gpus: the multi-GPU context is the CLI's only code path
    (ptb_multi with one device degenerates to ptb_launch), so the accumulator must equal the library's bit for bit."""
    import torch
    from scenes import load_config
    cfg = assets.ensure("c1", small=True)
    common = ["--dim=80x48", "-s", 2, "--depth", 4, "--launches", 2, "--batch", 2, "--scene", cfg["files"][0], "--env", cfg["env"], "--scale", cfg["scale"]]
    r = _run("-f", tmp_path / "a.png", "--accum-out", tmp_path / "a.ptba", *common)
    assert r.returncode == 0, r.stderr
    acc = ptb.load_accum_raw(tmp_path / "a.ptba")
    sc = load_config(ptb, assets, "c1", small=True)
    handle, _ = ctx.accel_build(sc)
    n = 80 * 48
    d_accum, d_frame = ctx.alloc(n * 16), ctx.alloc(n * 4)
    try:
        ctx.memset(d_accum, 0, n * 16)
        for sf in (0, 2):
            p = ptb.make_params(80, 48, subframe_index=sf, dof=True)
            p.accum_buffer, p.frame_buffer, p.handle = d_accum, d_frame, handle
            ctx.launch(p, ptb.default_render_cfg(spp_per_launch=2, max_depth=4, subframes_per_launch=2))
        want = ctx.to_host(d_accum, (48, 80, 4), np.float32)
    finally:
        ctx.free(d_accum); ctx.free(d_frame)
    assert np.array_equal(acc.view(np.uint32), want.view(np.uint32))
    r = _run("-f", tmp_path / "f.png", "--accum-out", tmp_path / "f.ptba", "--fast", *common)
    assert r.returncode == 0, r.stderr
    accf = ptb.load_accum_raw(tmp_path / "f.ptba")
    assert np.abs(accf[..., :3] - acc[..., :3]).mean() < 2e-2 * acc[..., :3].mean()
    if torch.cuda.device_count() >= 2:
        r = _run("-f", tmp_path / "t.png", "--accum-out", tmp_path / "t.ptba", "--gpus", 2, "--split", "tiles", *common)
        assert r.returncode == 0 and "2 GPUs, tile split" in r.stdout, r.stderr
        assert np.array_equal(ptb.load_accum_raw(tmp_path / "t.ptba").view(np.uint32), want.view(np.uint32))
        r = _run("-f", tmp_path / "s.png", "--accum-out", tmp_path / "s.ptba", "--gpus", 2, *common)
        assert r.returncode == 0 and "sample split" in r.stdout, r.stderr
        accs = ptb.load_accum_raw(tmp_path / "s.ptba")
        assert np.abs(accs[..., :3] - want[..., :3]).max() <= 2e-6 * np.abs(want[..., :3]).max() + 1e-7
    else:
        r = _run("-f", tmp_path / "t.png", "--gpus", 2, *common)
        assert r.returncode == 1 and "Caught exception" in r.stderr   # no second device on this box
