#!/usr/bin/env python3
"""Seeded synthetic stand-ins for the 16 blobs the reference checkout lacks
(/root/reference/.MISSING_LARGE_BLOBS: env1..5.exr, statue1..4.obj, lion.obj,
model.obj, model_albedo.png, suitcase_albedo.png, test_{albedo,normal,roughness}.png).

Everything is generated with numpy.random.default_rng(seed) and written into
assets/_gen/<config>/ (git-ignored).  The real meshes/maps under assets/ are
linked next to the synthetic files so that the reference's file-name
convention <stem>_{albedo,roughness,normal,metallic}.png (optixSphere.cpp:522-546)
finds them.  Configurations follow SURVEY.md section 8d / BASELINE.md:

  c1  test.obj + synthetic test_{albedo,normal,roughness}.png, env1 1024x512
  c2  monkey.obj + real monkey_albedo.png, env2 2048x1024          (bench workload)
  c3  suitcase.obj + real metallic/normal/roughness + synthetic albedo, env3 4096x2048
  c4  fish.obj + tower.obj + synthetic statue1-4/lion icospheres, env4 2048x1024
  c5  synthetic model.obj (0.5 M tris, uvs) + model_albedo.png, env5 4096x2048
"""
from __future__ import annotations

import os
import shutil
import struct
import sys
import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ASSETS = ROOT / "assets"
GEN = ASSETS / "_gen"


# ---- OpenEXR scanline writer (FLOAT or HALF, NONE / ZIPS / ZIP) -------------------
def write_exr(path, img: np.ndarray, compression: str = "zip", half: bool = False, channels: str = "RGB"):
    """img: float32 [h, w, len(channels)], row 0 = top scanline."""
    img = np.asarray(img, np.float32)
    h, w, nc = img.shape
    assert nc == len(channels)
    comp = {"none": 0, "zips": 2, "zip": 3}[compression]
    lines_per_block = 16 if comp == 3 else 1
    order = sorted(range(nc), key=lambda i: channels[i])  # channels are stored alphabetically
    ptype = 1 if half else 2

    def attr(name, typ, data):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(data)) + data

    chlist = b"".join(channels[i].encode() + b"\0" + struct.pack("<iBBBBii", ptype, 0, 0, 0, 0, 1, 1) for i in order) + b"\0"
    box = struct.pack("<iiii", 0, 0, w - 1, h - 1)
    header = struct.pack("<ii", 20000630, 2)
    header += attr("channels", "chlist", chlist)
    header += attr("compression", "compression", bytes([comp]))
    header += attr("dataWindow", "box2i", box) + attr("displayWindow", "box2i", box)
    header += attr("lineOrder", "lineOrder", b"\0")
    header += attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
    header += attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
    header += attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
    header += b"\0"
    nblocks = (h + lines_per_block - 1) // lines_per_block
    planes = img[:, :, order].astype(np.float16 if half else np.float32)  # [h, w, nc] in file channel order
    chunks = []
    for b in range(nblocks):
        y0 = b * lines_per_block
        rows = planes[y0:y0 + lines_per_block]            # [nl, w, nc]
        raw = np.ascontiguousarray(rows.transpose(0, 2, 1)).tobytes()  # per line: channel-major
        data = raw
        if comp:
            a = np.frombuffer(raw, np.uint8)
            t = np.concatenate([a[0::2], a[1::2]])
            d = t.copy()
            d[1:] = (t[1:].astype(np.int32) - t[:-1].astype(np.int32) + 128) & 255
            z = zlib.compress(d.tobytes(), 6)
            data = z if len(z) < len(raw) else raw
        chunks.append(struct.pack("<ii", y0, len(data)) + data)
    offset = len(header) + 8 * nblocks
    table = b""
    for c in chunks:
        table += struct.pack("<Q", offset)
        offset += len(c)
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as f:
        f.write(header + table + b"".join(chunks))


def write_png(path, img: np.ndarray):
    """img: uint8 [h, w] (gray) or [h, w, 3|4]; plain zlib PNG writer (no PIL dependency)."""
    img = np.ascontiguousarray(img, np.uint8)
    if img.ndim == 2:
        ctype, row = 0, img
    else:
        ctype, row = {3: 2, 4: 6}[img.shape[2]], img.reshape(img.shape[0], -1)
    h, w = img.shape[:2]
    raw = np.concatenate([np.zeros((h, 1), np.uint8), row], axis=1).tobytes()

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)

    Path(path).parent.mkdir(parents=True, exist_ok=True)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


# ---- procedural content -------------------------------------------------------------
def _smooth_noise(rng, h, w, cells):
    g = rng.random((cells + 1, 2 * cells + 1)).astype(np.float32)
    ys = np.linspace(0, cells, h, endpoint=False, dtype=np.float32)
    xs = np.linspace(0, 2 * cells, w, endpoint=False, dtype=np.float32)
    y0, x0 = ys.astype(int), xs.astype(int)
    fy, fx = (ys - y0)[:, None], (xs - x0)[None, :]
    fy, fx = fy * fy * (3 - 2 * fy), fx * fx * (3 - 2 * fx)
    g[:, -1] = g[:, 0]  # wrap in longitude
    a = g[y0][:, x0] * (1 - fx) + g[y0][:, x0 + 1] * fx
    b = g[y0 + 1][:, x0] * (1 - fx) + g[y0 + 1][:, x0 + 1] * fx
    return a * (1 - fy) + b * fy


def make_env(seed: int, w: int, h: int) -> np.ndarray:
    """Equirect sky: the reference's own fallback sky (optixSphere.cu:552-557: (0.4,0.4,0.6) plus a
    (200,175,125) sun around normalize(0,2,3)) modulated by low-frequency noise and a horizon gradient."""
    rng = np.random.default_rng(seed)
    v = (np.arange(h, dtype=np.float32) + 0.5) / h
    u = (np.arange(w, dtype=np.float32) + 0.5) / w
    theta = (0.5 - v) * np.pi             # asin(y)
    phi = (u - 0.5) * 2 * np.pi           # atan2(z, x)
    y = np.sin(theta)[:, None] * np.ones((1, w), np.float32)
    x = np.cos(theta)[:, None] * np.cos(phi)[None, :]
    z = np.cos(theta)[:, None] * np.sin(phi)[None, :]
    sun = np.array([0.0, 2.0, 3.0], np.float32)
    sun /= np.linalg.norm(sun)
    cosang = x * sun[0] + y * sun[1] + z * sun[2]
    sky = np.array([0.4, 0.4, 0.6], np.float32)[None, None, :] * (0.6 + 0.8 * np.clip(y, 0, 1))[:, :, None]
    ground = np.array([0.25, 0.22, 0.2], np.float32)[None, None, :] * np.ones((h, w, 1), np.float32)
    img = np.where((y >= 0)[:, :, None], sky, ground)
    img = img * (0.75 + 0.5 * _smooth_noise(rng, h, w, 8))[:, :, None]
    img = np.where((cosang > 0.99)[:, :, None], np.array([200.0, 175.0, 125.0], np.float32)[None, None, :], img)
    return img.astype(np.float32)


def make_albedo(seed: int, n: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n]
    cells = 16
    checker = (((yy * cells) // n + (xx * cells) // n) % 2).astype(np.float32)
    palette = rng.random((cells, cells, 3)).astype(np.float32) * 0.6 + 0.3
    base = palette[(yy * cells) // n, (xx * cells) // n]
    img = base * (0.55 + 0.45 * checker[:, :, None])
    return (np.clip(img, 0, 1) * 255 + 0.5).astype(np.uint8)


def make_normal(seed: int, n: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float32) / n
    fx, fy = rng.integers(3, 9, 2)
    hx = 0.35 * np.cos(2 * np.pi * fx * xx) * np.sin(2 * np.pi * fy * yy)
    hy = 0.35 * np.sin(2 * np.pi * fx * xx) * np.cos(2 * np.pi * fy * yy)
    nrm = np.stack([-hx, -hy, np.ones_like(hx)], -1)
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    return (np.clip(nrm * 0.5 + 0.5, 0, 1) * 255 + 0.5).astype(np.uint8)


def make_roughness(seed: int, n: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:n, 0:n].astype(np.float32) / n
    g = 0.08 + 0.8 * (0.5 * xx + 0.5 * yy) + 0.1 * rng.random((n, n)).astype(np.float32)
    return (np.clip(g, 0, 1) * 255 + 0.5).astype(np.uint8)


def icosphere(subdiv: int):
    t = (1 + 5 ** 0.5) / 2
    v = np.array([[-1, t, 0], [1, t, 0], [-1, -t, 0], [1, -t, 0], [0, -1, t], [0, 1, t], [0, -1, -t], [0, 1, -t],
                  [t, 0, -1], [t, 0, 1], [-t, 0, -1], [-t, 0, 1]], np.float64)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    f = np.array([[0, 11, 5], [0, 5, 1], [0, 1, 7], [0, 7, 10], [0, 10, 11], [1, 5, 9], [5, 11, 4], [11, 10, 2], [10, 7, 6],
                  [7, 1, 8], [3, 9, 4], [3, 4, 2], [3, 2, 6], [3, 6, 8], [3, 8, 9], [4, 9, 5], [2, 4, 11], [6, 2, 10],
                  [8, 6, 7], [9, 8, 1]], np.int64)
    for _ in range(subdiv):
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        e.sort(axis=1)
        ue, inv = np.unique(e, axis=0, return_inverse=True)
        mid = v[ue[:, 0]] + v[ue[:, 1]]
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        base = len(v)
        v = np.concatenate([v, mid])
        n = len(f)
        m01, m12, m20 = base + inv[:n], base + inv[n:2 * n], base + inv[2 * n:]
        f = np.concatenate([np.stack([f[:, 0], m01, m20], 1), np.stack([f[:, 1], m12, m01], 1),
                            np.stack([f[:, 2], m20, m12], 1), np.stack([m01, m12, m20], 1)])
    return v, f


def write_blob_obj(path, subdiv: int, seed: int, radius=1.0, center=(0, 0, 0), with_uv=False):
    """Displaced icosphere (20 * 4^subdiv triangles) with vertex normals."""
    rng = np.random.default_rng(seed)
    v, f = icosphere(subdiv)
    k = rng.normal(size=(6, 3))
    ph = rng.random(6) * 6.28
    disp = sum(0.06 * np.sin(v @ (k[i] * (2 + i)) + ph[i]) for i in range(6))
    p = v * (1.0 + disp)[:, None]
    # vertex normals from face normals
    fn = np.cross(p[f[:, 1]] - p[f[:, 0]], p[f[:, 2]] - p[f[:, 0]])
    vn = np.zeros_like(p)
    for c in range(3):
        np.add.at(vn, f[:, c], fn)
    vn /= np.maximum(np.linalg.norm(vn, axis=1, keepdims=True), 1e-20)
    p = p * radius + np.asarray(center, np.float64)[None, :]
    Path(path).parent.mkdir(parents=True, exist_ok=True)
    def rows(fh, fmt, arr, block=65536):
        # one C-level % over a whole block of rows: ~6x faster than np.savetxt, byte-identical output
        ncol = arr.shape[1]
        for i in range(0, len(arr), block):
            chunk = arr[i:i + block]
            fh.write(((fmt + "\n") * len(chunk)) % tuple(chunk.ravel().tolist()))
        assert fmt.count("%") == ncol

    with open(path, "w") as fh:
        fh.write(f"# synthetic stand-in, seed {seed}, {len(f)} triangles\n")
        rows(fh, "v %.6f %.6f %.6f", p)
        rows(fh, "vn %.4f %.4f %.4f", vn)
        if with_uv:
            uv = np.stack([0.5 + np.arctan2(v[:, 2], v[:, 0]) / (2 * np.pi), 0.5 + np.arcsin(np.clip(v[:, 1], -1, 1)) / np.pi], 1)
            rows(fh, "vt %.6f %.6f", uv)
            idx = f + 1
            rows(fh, "f %d/%d/%d %d/%d/%d %d/%d/%d", np.stack([idx[:, 0]] * 3 + [idx[:, 1]] * 3 + [idx[:, 2]] * 3, 1))
        else:
            idx = f + 1
            rows(fh, "f %d//%d %d//%d %d//%d", np.stack([idx[:, 0]] * 2 + [idx[:, 1]] * 2 + [idx[:, 2]] * 2, 1))
    return len(f)


def _link(src: Path, dst: Path):
    dst.parent.mkdir(parents=True, exist_ok=True)
    if dst.exists() or dst.is_symlink():
        dst.unlink()
    shutil.copyfile(src, dst)


def ensure(config: str, small: bool = False) -> dict:
    """Create (once) the files of one configuration; returns {'files': [...obj], 'env': path, 'scale': s}."""
    d = GEN / (config + ("_small" if small else ""))
    stamp = d / ".done"
    cfgs = {
        "c1": dict(objs=["test.obj"], scale=0.05, env=(1, 1024, 512)),
        "c2": dict(objs=["monkey.obj"], scale=1.0, env=(2, 2048, 1024)),
        "c3": dict(objs=["suitcase.obj"], scale=0.05, env=(3, 4096, 2048)),
        "c4": dict(objs=["fish.obj", "tower.obj"], scale=1.0, env=(4, 2048, 1024)),
        "c5": dict(objs=[], scale=1.0, env=(5, 4096, 2048)),
    }
    c = cfgs[config]
    es, ew, eh = c["env"]
    if small:
        ew, eh = ew // 8, eh // 8
    out = dict(files=[str(d / o) for o in c["objs"]], env=str(d / f"env{es}.exr"), scale=c["scale"])
    if config == "c4":
        out["files"] += [str(d / f"statue{i}.obj") for i in range(1, 5)] + [str(d / "lion.obj")]
    if config == "c5":
        out["files"] = [str(d / "model.obj")]
    if stamp.exists():
        return out
    # concurrent callers (two gloo ranks, pytest-xdist workers) must not read half-written files:
    # one generates under an exclusive lock, the others wait and then find the stamp
    import fcntl
    GEN.mkdir(parents=True, exist_ok=True)
    with open(GEN / ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not stamp.exists():
            _generate(config, d, c, es, ew, eh, small)
            stamp.write_text("ok\n")
    return out


def _generate(config, d, c, es, ew, eh, small):
    d.mkdir(parents=True, exist_ok=True)
    write_exr(d / f"env{es}.exr", make_env(es, ew, eh), compression="zip")
    tex_n = 128 if small else None
    if config == "c1":
        _link(ASSETS / "test.obj", d / "test.obj")
        n = tex_n or 1024
        write_png(d / "test_albedo.png", make_albedo(11, n))
        write_png(d / "test_normal.png", make_normal(12, n))
        write_png(d / "test_roughness.png", make_roughness(13, n))
    elif config == "c2":
        _link(ASSETS / "monkey.obj", d / "monkey.obj")
        _link(ASSETS / "monkey_albedo.png", d / "monkey_albedo.png")
    elif config == "c3":
        for f in ("suitcase.obj", "suitcase_metallic.png", "suitcase_normal.png", "suitcase_roughness.png"):
            _link(ASSETS / f, d / f)
        write_png(d / "suitcase_albedo.png", make_albedo(31, tex_n or 2048))
    elif config == "c4":
        _link(ASSETS / "fish.obj", d / "fish.obj")
        _link(ASSETS / "tower.obj", d / "tower.obj")
        # statue1..4, lion: 0.33 M / 0.33 M / 1.3 M / 1.3 M / 1.3 M triangles (20*4^7, 20*4^8); small: 20*4^3
        subs = [3] * 5 if small else [7, 7, 8, 8, 8]
        names = ["statue1", "statue2", "statue3", "statue4", "lion"]
        for i, (nm, sd) in enumerate(zip(names, subs)):
            write_blob_obj(d / f"{nm}.obj", sd, 41 + i, radius=0.8, center=(-4.0 + 2.0 * i, 0.8, -1.5))
    elif config == "c5":
        write_blob_obj(d / "model.obj", 3 if small else 7, 51, radius=1.2, center=(0, 1.2, 0), with_uv=True)
        write_png(d / "model_albedo.png", make_albedo(52, tex_n or 2048))


if __name__ == "__main__":
    for name in (sys.argv[1:] or ["c1", "c2"]):
        small = name.endswith("_small")
        print(name, ensure(name.replace("_small", ""), small))
