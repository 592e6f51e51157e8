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
