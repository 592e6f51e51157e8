"""Scene-load row of the hot path (SURVEY.md section 8 a13/a15/a16 + f2): OBJ reader vs the reference's vendored
tiny_obj_loader.h (golden hashes written by oracle/_ref/ref_probe), createSceneGeometry rules, image decoders, camera."""
import hashlib
import json
import subprocess
import tempfile
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
OBJS = {"test": 12, "monkey": 15744, "suitcase": 2204, "fish": 8168, "tower": 4802}


@pytest.mark.parametrize("name", list(OBJS))
def test_obj_reader_matches_tinyobj_golden(ptb, name):
    g = json.loads((ROOT / "tests" / "golden" / f"obj_{name}.json").read_text())
    rec = ptb.obj_read(ROOT / "assets" / f"{name}.obj")
    assert rec.shape[0] == g["face_vertices"] == 3 * OBJS[name]
    assert hashlib.sha256(rec.tobytes()).hexdigest() == g["sha256_face_vertex_stream"]


@pytest.mark.parametrize("chunk", [64, 1000, 100000])
def test_obj_reader_chunked_parse_matches_golden(ptb, tmp_path, monkeypatch, chunk):
    """Large files are cut at line ends and parsed by several threads; forcing tiny chunks on the real meshes (and on a file
    with relative indices, CRLF line ends and polygons) must give the same face-vertex stream as one sequential pass."""
    p = tmp_path / "edge.obj"
    p.write_bytes(b"# c\r\nv 0 0 0\r\nv 1e0 0 0\nv 1 1.5E+0 0\nv 0 1 -2.5e-1\nvn 0 0 2\nvt 0.25 0.75\n"
                  b"f 1 2 3 4\nf -4//1 -3//1 -2//1\nv 2 2 2\rf -1 -2 -3 -4 -5\r\nf 1/1 2/1 3/1\n\n")
    whole = {n: ptb.obj_read(ROOT / "assets" / f"{n}.obj") for n in ("monkey", "tower")}
    whole["edge"] = ptb.obj_read(p)
    monkeypatch.setenv("PTB_OBJ_CHUNK_BYTES", str(chunk))
    for n, ref in whole.items():
        got = ptb.obj_read(p if n == "edge" else ROOT / "assets" / f"{n}.obj")
        assert np.array_equal(got, ref), (n, chunk)
    g = json.loads((ROOT / "tests" / "golden" / "obj_monkey.json").read_text())
    assert hashlib.sha256(whole["monkey"].tobytes()).hexdigest() == g["sha256_face_vertex_stream"]


def test_obj_reader_matches_live_tinyobj(ptb, oh):
    if not oh.REF_PROBE.exists():
        pytest.skip("oracle/_ref/ref_probe not built")
    with tempfile.NamedTemporaryFile(suffix=".bin") as tf:
        subprocess.run([str(oh.REF_PROBE), "obj", str(ROOT / "assets" / "tower.obj"), tf.name], check=True, capture_output=True)
        raw = np.fromfile(tf.name, np.uint32).reshape(-1, 10)
    assert np.array_equal(raw, ptb.obj_read(ROOT / "assets" / "tower.obj"))


def test_obj_edge_cases(ptb, tmp_path):
    """Ragged input: negative (relative) indices, v//vn and v-only triples, polygons (fan), CRLF, exponents, empty file."""
    p = tmp_path / "edge.obj"
    p.write_bytes(b"# c\r\nv 0 0 0\r\nv 1e0 0 0\nv 1 1.5E+0 0\nv 0 1 -2.5e-1\nvn 0 0 2\nvt 0.25 0.75\n"
                  b"f 1 2 3 4\nf -4//1 -3//1 -2//1\nf 1/1 2/1 3/1\n\n")
    rec = ptb.obj_read(p)
    f = rec[:, :8].view(np.float32)
    flags = rec[:, 8:].view(np.int32)
    assert rec.shape[0] == 12  # quad -> 2 triangles (0,1,2),(0,2,3); + 2 triangles
    assert np.allclose(f[0, :3], 0) and np.allclose(f[4, :3], (1, 1.5, 0)) and np.allclose(f[5, :3], (0, 1, -0.25))
    assert list(flags[:6].ravel()) == [0] * 12
    assert np.array_equal(f[6:9, :3], f[[0, 1, 2], :3]) and np.all(flags[6:9, 0] == 1) and np.all(f[6:9, 5] == 2.0)
    assert np.all(flags[9:12, 1] == 1) and np.all(f[9:12, 6] == 0.25)
    (tmp_path / "empty.obj").write_text("")
    assert ptb.obj_read(tmp_path / "empty.obj").shape[0] == 0
    with pytest.raises(ptb.PtbError) as e:
        ptb.obj_read(tmp_path / "missing.obj")
    assert e.value.code == ptb.PTB_ERR_IO
    (tmp_path / "bad.obj").write_text("v 0 0 0\nf 1 2 3\n")
    with pytest.raises(ptb.PtbError):
        ptb.obj_read(tmp_path / "bad.obj")


def test_create_scene_geometry_rules(ptb, assets):
    """optixSphere.cpp:400-649: scale, normalised normals, one material per file, floor at the lowest vertex."""
    cfg = assets.ensure("c1", small=True)
    sc = ptb.Scene.load_obj(cfg["files"], scale=0.05, material_seed=1)
    t = sc.triangles()
    assert t.shape == (14, 32) and sc.num_materials == 2
    raw = ptb.obj_read(cfg["files"][0])[:, :8].view(np.float32)
    v = t[:12, 0:12].reshape(36, 4)
    assert np.array_equal(v[:, :3], raw[:, :3] * np.float32(0.05)) and np.all(v[:, 3] == 0)
    n = t[:12, 12:24].reshape(36, 4)[:, :3]
    assert np.allclose(np.linalg.norm(n, axis=1), 1, atol=1e-6)
    uv = t[:12, 24:30].reshape(36, 2)
    assert np.array_equal(uv, raw[:, 6:8])
    # floor: two triangles, half-size 200 at y = min vertex height, appended last with its own material
    miny = v[:, 1].min()
    floor = t[12:, 0:12].reshape(6, 4)
    assert np.all(floor[:, 1] == miny) and set(np.abs(floor[:, [0, 2]]).ravel()) == {200.0}
    assert np.array_equal(floor[:3, :3], [[-200, miny, -200], [-200, miny, 200], [200, miny, -200]])
    assert np.array_equal(floor[3:, :3], [[200, miny, -200], [-200, miny, 200], [200, miny, 200]])
    assert list(sc.material_ids()) == [0] * 12 + [1] * 2
    m0, m1 = sc.material(0), sc.material(1)
    # textured file -> neutral fallback material (optixSphere.cpp:555-571)
    assert m0.has_albedo and m0.has_normal and m0.has_roughness and not m0.has_metallic
    assert list(m0.diffuse_color) == [0.5] * 3 and abs(m0.roughness - 0.4) < 1e-7 and not m0.metallic
    assert np.allclose(list(m1.diffuse_color), 0.2) and abs(m1.roughness - 0.1) < 1e-7 and list(m1.emission_color) == [0, 0, 0]
    tex = sc.texture(0, 0)
    assert tex.shape == (128, 128, 4) and tex.max() <= 1.0 and np.all(tex[..., 3] == 1.0)


def test_untextured_material_is_seeded(ptb):
    """optixSphere.cpp:572-582 with an explicit seed instead of std::random_device."""
    f = [str(ROOT / "assets" / "fish.obj"), str(ROOT / "assets" / "tower.obj")]
    a, b, c = (ptb.Scene.load_obj(f, 1.0, s) for s in (4, 4, 5))
    ma, mb, mc = a.material(0), b.material(0), c.material(0)
    assert list(ma.diffuse_color) == list(mb.diffuse_color) and ma.roughness == mb.roughness
    assert list(ma.diffuse_color) != list(mc.diffuse_color)
    assert list(ma.specular) == list(ma.diffuse_color) and not ma.has_albedo
    assert list(a.material(1).diffuse_color) != list(ma.diffuse_color)
    assert a.num_materials == 3 and a.num_triangles == 8168 + 4802 + 2
    e = np.array(ma.emission_color)
    assert np.all(e == 0) or np.allclose(e, 100 * np.array(ma.diffuse_color))


def test_png_decoder_matches_pil(ptb, tmp_path):
    from PIL import Image
    for name in ("monkey_albedo.png", "suitcase_metallic.png"):
        ours = ptb.load_image_rgba8(ROOT / "assets" / name)
        ref = np.asarray(Image.open(ROOT / "assets" / name).convert("RGBA"))
        assert np.array_equal(ours, ref), name
    rng = np.random.default_rng(1)
    cases = {"L": rng.integers(0, 256, (37, 53), dtype=np.uint8), "RGB": rng.integers(0, 256, (19, 31, 3), dtype=np.uint8),
             "RGBA": rng.integers(0, 256, (8, 9, 4), dtype=np.uint8), "LA": rng.integers(0, 256, (5, 7, 2), dtype=np.uint8)}
    for mode, arr in cases.items():
        p = tmp_path / f"{mode}.png"
        Image.fromarray(arr, mode).save(p)
        assert np.array_equal(ptb.load_image_rgba8(p), np.asarray(Image.open(p).convert("RGBA"))), mode
    pal = Image.fromarray(cases["RGB"], "RGB").convert("P", palette=Image.ADAPTIVE, colors=17)
    pal.save(tmp_path / "pal.png")
    assert np.array_equal(ptb.load_image_rgba8(tmp_path / "pal.png"), np.asarray(pal.convert("RGBA")))
    Image.fromarray(cases["RGB"], "RGB").save(tmp_path / "inter.png", interlace=True) if False else None
    g16 = (rng.integers(0, 65536, (6, 11), dtype=np.uint16))
    Image.fromarray(g16, "I;16").save(tmp_path / "g16.png")
    assert np.array_equal(ptb.load_image_rgba8(tmp_path / "g16.png")[..., 0], (g16 >> 8).astype(np.uint8))  # stb: high byte
    bit1 = Image.fromarray((rng.random((9, 13)) > 0.5)).convert("1")
    bit1.save(tmp_path / "b1.png")
    assert np.array_equal(ptb.load_image_rgba8(tmp_path / "b1.png")[..., 0], np.asarray(bit1.convert("L")))
    (tmp_path / "junk.png").write_bytes(b"not a png")
    with pytest.raises(ptb.PtbError):
        ptb.load_image_rgba8(tmp_path / "junk.png")


def test_png_writer_round_trip(ptb, tmp_path):
    rng = np.random.default_rng(2)
    img = rng.integers(0, 256, (21, 34, 4), dtype=np.uint8)
    ptb.save_image(tmp_path / "o.png", img, flip_y=False)
    assert np.array_equal(ptb.load_image_rgba8(tmp_path / "o.png"), img)
    ptb.save_image(tmp_path / "f.png", img, flip_y=True)  # row 0 of a frame buffer is the bottom row (optixSphere.cu:332,400)
    assert np.array_equal(ptb.load_image_rgba8(tmp_path / "f.png"), img[::-1])
    ptb.save_image(tmp_path / "o.ppm", img, flip_y=False)
    raw = (tmp_path / "o.ppm").read_bytes()
    assert raw.startswith(b"P6\n34 21\n255\n") and np.array_equal(np.frombuffer(raw[-21 * 34 * 3:], np.uint8).reshape(21, 34, 3), img[..., :3])
    with pytest.raises(ptb.PtbError):
        ptb.save_image(tmp_path / "o.tiff", img)


@pytest.mark.parametrize("compression,half", [("none", False), ("zips", False), ("zip", False), ("zip", True)])
def test_exr_reader(ptb, assets, tmp_path, compression, half):
    rng = np.random.default_rng(3)
    img = (rng.random((37, 50, 3)) * 40).astype(np.float32)
    img[0, 0] = (200.0, 175.0, 125.0)
    p = tmp_path / "t.exr"
    assets.write_exr(p, img, compression=compression, half=half)
    got = ptb.load_image_float4(p)
    want = img.astype(np.float16).astype(np.float32) if half else img
    assert got.shape == (37, 50, 4) and np.array_equal(got[..., :3], want) and np.all(got[..., 3] == 1.0)
    rgba = np.concatenate([img, rng.random((37, 50, 1)).astype(np.float32)], -1)
    assets.write_exr(p, rgba, compression=compression, half=False, channels="RGBA")
    assert np.array_equal(ptb.load_image_float4(p), rgba)


def test_exr_reader_against_opencv(ptb, assets, tmp_path):
    import os
    os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
    import cv2
    img = assets.make_env(9, 64, 32)
    p = tmp_path / "cv.exr"
    if not cv2.imwrite(str(p), img[..., ::-1].copy()):
        pytest.skip("this OpenCV build cannot write EXR")
    got = ptb.load_image_float4(p)
    assert np.array_equal(got[..., :3], img)
    ours = tmp_path / "ours.exr"
    assets.write_exr(ours, img)
    back = cv2.imread(str(ours), cv2.IMREAD_UNCHANGED)
    assert back is not None and np.array_equal(back[..., :3][..., ::-1], img)


def test_camera_uvw(ptb, oh):
    """sutil::Camera::UVWFrame with the reference camera (optixSphere.cpp:104-107): product == oracle, and analytic."""
    import ctypes as C
    U, V, W = ptb.camera_uvw((0, 2, 6), (0, 0, 0), (0, 1, 0), 50.0, 1920 / 1080)
    oU, oV, oW = ((C.c_float * 3)() for _ in range(3))
    oh.load("oracle").orc_camera_uvw((C.c_float * 3)(0, 2, 6), (C.c_float * 3)(0, 0, 0), (C.c_float * 3)(0, 1, 0),
                                     C.c_float(50.0), C.c_float(1920 / 1080), oU, oV, oW)
    assert list(U) == list(oU) and list(V) == list(oV) and list(W) == list(oW)
    assert np.allclose(W, (0, -2, -6)) and abs(np.dot(U, V)) < 1e-5 and abs(np.dot(U, W)) < 1e-5
    wl = np.linalg.norm(W)
    assert np.isclose(np.linalg.norm(V), wl * np.tan(np.radians(25.0)), rtol=1e-6)
    assert np.isclose(np.linalg.norm(U) / np.linalg.norm(V), 1920 / 1080, rtol=1e-6)
    p = ptb.make_params(640, 480)
    assert (p.eye.x, p.eye.y, p.eye.z) == (0.0, 2.0, 6.0) and p.origin_x == 320 and p.dof is True


def test_raw_scene_and_material_table(ptb):
    """ptb_scene_create + ptb_scene_set_materials: HitGroupData semantics (optixSphere.cpp:1196-1261)."""
    tri = np.zeros((1, 32), np.float32)
    tri[0, 0:12] = [0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0]
    tri[0, 12:24] = [0, 0, 1, 0] * 3
    sc = ptb.Scene.from_triangles(tri, np.zeros(1, np.uint32))
    tex8 = (np.arange(4 * 6 * 4, dtype=np.float32).reshape(4, 6, 4) % 256) / np.float32(255.0)
    texf = np.full((2, 2, 4), 0.123456, np.float32)
    sc.set_materials([dict(diffuse_color=(0.1, 0.2, 0.3), roughness=0.7, metallic=True, emission_color=(1, 2, 3),
                           albedo=tex8, roughness_map=texf)])
    m = sc.material(0)
    assert np.allclose(list(m.diffuse_color), (0.1, 0.2, 0.3)) and m.metallic == 1 and m.albedo_w == 6 and m.albedo_h == 4
    assert np.array_equal(sc.texture(0, 0), tex8)      # exactly byte/255 -> kept as RGBA8, widened identically
    assert np.array_equal(sc.texture(0, 1), texf)      # anything else stays float4
    sc.set_materials([dict(transparent=True)])         # HitGroupData.transparent (optixSphere.cpp:1215): the glass branch
    assert sc.material(0).transparent == 1
    sc2 = ptb.Scene.from_triangles(tri, np.array([2], np.uint32))
    with pytest.raises(ptb.PtbError):
        sc2.set_materials([dict()])                    # material id beyond the table


def test_demo_scene(ptb):
    """createSceneGeometry(loadFromFile=false), optixSphere.cpp:650-751: 2 + 3*16*32*2 triangles, 4 fixed materials."""
    sc = ptb.Scene.demo()
    assert sc.num_triangles == 2 + 3 * 16 * 32 * 2 and sc.num_materials == 4
    t = sc.triangles()
    ids = sc.material_ids()
    assert list(ids[:2]) == [0, 0] and set(ids[2:2 + 1024]) == {1} and set(ids[-1024:]) == {3}
    v = t[:, 0:12].reshape(-1, 4)
    assert np.all(v[:, 3] == 1.0)  # the demo scene stores w = 1 (files store 0)
    centers = np.array([[-3, 1, 0], [0, 1, 0], [3, 1, 0]], np.float32)
    for k in range(3):
        vv = t[2 + k * 1024: 2 + (k + 1) * 1024, 0:12].reshape(-1, 4)[:, :3]
        assert np.allclose(np.linalg.norm(vv - centers[k], axis=1), 1.0, atol=1e-5)
    m = sc.material(1)
    assert list(m.diffuse_color) == [1.0, 0.0, 0.0] and m.roughness == 0.0 and sc.material(0).roughness == np.float32(0.8)
