"""Test-side access to the CPU oracle (oracle/liboracle.so) and to the host-compiled
reference (oracle/_ref/libref_pt.so).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"
ORACLE_LIB = ORACLE_DIR / "liboracle.so"
REF_LIB = ORACLE_DIR / "_ref" / "libref_pt.so"
REF_PROBE = ORACLE_DIR / "_ref" / "ref_probe"
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))


class OrcTexture(C.Structure):
    _fields_ = [("rgba", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("has", C.c_int32), ("_pad", C.c_int32)]


class OrcMaterial(C.Structure):
    _fields_ = [("albedo", OrcTexture), ("roughness", OrcTexture), ("normal", OrcTexture), ("metallic", OrcTexture),
                ("emission_color", C.c_float * 3), ("diffuse_color", C.c_float * 3), ("specular", C.c_float * 3),
                ("roughness_value", C.c_float), ("metallic_flag", C.c_int32), ("transparent_flag", C.c_int32)]


class OrcScene(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("normals", C.c_void_p), ("texcoords", C.c_void_p), ("mat_ids", C.c_void_p),
                ("num_tris", C.c_uint32), ("num_mats", C.c_int32), ("mats", C.c_void_p), ("env_rgba", C.c_void_p),
                ("env_w", C.c_int32), ("env_h", C.c_int32)]


class OrcParams(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("subframe_index", C.c_int32), ("dof", C.c_int32),
                ("eye", C.c_float * 3), ("U", C.c_float * 3), ("V", C.c_float * 3), ("W", C.c_float * 3)]


class OrcConfig(C.Structure):
    _fields_ = [("spp_per_launch", C.c_int32), ("max_depth", C.c_int32), ("tmin", C.c_float), ("tmax", C.c_float),
                ("dof_blur", C.c_float), ("focus_dist", C.c_float), ("nmap_strength", C.c_float), ("exposure", C.c_float),
                ("gamma", C.c_float), ("contrast", C.c_float), ("sat_cuda", C.c_int32), ("use_bvh", C.c_int32),
                ("threads", C.c_int32), ("accumulate_sum", C.c_int32)]


class OrcStats(C.Structure):
    _fields_ = [("segments", C.c_uint64), ("paths", C.c_uint64), ("hits", C.c_uint64), ("misses", C.c_uint64),
                ("seconds", C.c_double), ("threads", C.c_int32), ("_pad", C.c_int32)]


def build_oracle():
    """make -C oracle: liboracle.so always; oracle/_ref only where /root/reference exists."""
    r = subprocess.run(["make", "-C", str(ORACLE_DIR)], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])


_libs = {}


def load(which="oracle") -> C.CDLL:
    path = ORACLE_LIB if which == "oracle" else REF_LIB
    if which not in _libs:
        if not path.exists():
            build_oracle()
        L = C.CDLL(str(path))
        L.orc_rng_next.restype = C.c_uint32
        L.orc_atan2.restype = C.c_float
        L.orc_asin.restype = C.c_float
        L.orc_atan2.argtypes = [C.c_float, C.c_float]
        L.orc_asin.argtypes = [C.c_float]
        L.orc_pow.restype = C.c_float
        L.orc_pow.argtypes = [C.c_float, C.c_float]
        L.orc_impl_name.restype = C.c_char_p
        L.orc_closest_hit.restype = C.c_int32
        _libs[which] = L
    return _libs[which]


def have_ref() -> bool:
    return REF_LIB.exists()


def guard_texture(t: np.ndarray) -> np.ndarray:
    """Return a view of t whose memory is preceded by a copy of its last row.  The reference reads texel
    index -1 / row -1 (optixSphere.cu:509-510, 579-580), which is out of bounds; with this guard the
    host-compiled reference reads exactly what oracle rule R3 (negative linear index wraps by +w*h) reads."""
    t = np.ascontiguousarray(t, np.float32)
    h, w, _ = t.shape
    buf = np.empty((2 * h, w, 4), np.float32)
    buf[:h] = t
    buf[h:] = t
    return buf[h:]


class OracleScene:
    """Host arrays in the oracle's layout, built from a ptb Scene (same loader output the GPU sees)."""

    def __init__(self, tris32: np.ndarray, mat_ids: np.ndarray, mats: list, env: np.ndarray, guard=True):
        t = np.ascontiguousarray(tris32, np.float32).reshape(-1, 32)
        n = t.shape[0]
        self.vertices = np.ascontiguousarray(t[:, 0:12].reshape(n * 3, 4))
        self.normals = np.ascontiguousarray(t[:, 12:24].reshape(n * 3, 4))
        self.texcoords = np.ascontiguousarray(t[:, 24:30].reshape(n * 3, 2))
        self.mat_ids = np.ascontiguousarray(mat_ids, np.uint32)
        self.env = guard_texture(env) if guard else np.ascontiguousarray(env, np.float32)
        self._tex = []
        self.mats = (OrcMaterial * len(mats))()
        for i, m in enumerate(mats):
            om = self.mats[i]
            for key, fld in (("albedo", om.albedo), ("roughness_map", om.roughness), ("normal_map", om.normal), ("metallic_map", om.metallic)):
                tx = m.get(key)
                if tx is not None:
                    g = guard_texture(tx) if guard else np.ascontiguousarray(tx, np.float32)
                    self._tex.append(g)
                    fld.rgba, fld.w, fld.h, fld.has = g.ctypes.data, g.shape[1], g.shape[0], 1
            om.emission_color[:] = [float(x) for x in m.get("emission_color", (0, 0, 0))]
            om.diffuse_color[:] = [float(x) for x in m.get("diffuse_color", (0.5, 0.5, 0.5))]
            om.specular[:] = [float(x) for x in m.get("specular", (0.5, 0.5, 0.5))]
            om.roughness_value = float(m.get("roughness", 0.4))
            om.metallic_flag = int(bool(m.get("metallic", False)))
            om.transparent_flag = int(bool(m.get("transparent", False)))
        self.c = OrcScene()
        self.c.vertices, self.c.normals = self.vertices.ctypes.data, self.normals.ctypes.data
        self.c.texcoords, self.c.mat_ids = self.texcoords.ctypes.data, self.mat_ids.ctypes.data
        self.c.num_tris, self.c.num_mats = n, len(mats)
        self.c.mats = C.addressof(self.mats)
        self.c.env_rgba, self.c.env_w, self.c.env_h = self.env.ctypes.data, self.env.shape[1], self.env.shape[0]

    @classmethod
    def from_ptb(cls, scene, guard=True):
        mats = []
        for i in range(scene.num_materials):
            mi = scene.material(i)
            m = dict(emission_color=list(mi.emission_color), diffuse_color=list(mi.diffuse_color), specular=list(mi.specular),
                     roughness=mi.roughness, metallic=bool(mi.metallic), transparent=bool(mi.transparent))
            for kind, key, has in ((0, "albedo", mi.has_albedo), (1, "roughness_map", mi.has_roughness),
                                   (2, "normal_map", mi.has_normal), (3, "metallic_map", mi.has_metallic)):
                if has:
                    m[key] = scene.texture(i, kind)
            mats.append(m)
        return cls(scene.triangles(), scene.material_ids(), mats, scene.env(), guard=guard)


def default_config(which="oracle", **kw) -> OrcConfig:
    cfg = OrcConfig()
    load(which).orc_default_config(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def params_from_ptb(p) -> OrcParams:
    o = OrcParams()
    o.width, o.height, o.subframe_index, o.dof = p.image_width, p.image_height, p.subframe_index, int(p.dof)
    for name in ("eye", "U", "V", "W"):
        v = getattr(p, name)
        getattr(o, name)[:] = [v.x, v.y, v.z]
    return o


def render(which, scene: OracleScene, params: OrcParams, cfg: OrcConfig, accum=None, window=None, want_hits=True):
    """Returns (accum float32 [H,W,4], frame uint8 [H,W,4], primary_hit int32 [H,W], stats, rc)."""
    W, H = params.width, params.height
    if accum is None:
        accum = np.zeros((H, W, 4), np.float32)
    frame = np.zeros((H, W, 4), np.uint8)
    hits = np.full((H, W), -2, np.int32) if want_hits else None
    st = OrcStats()
    x0, y0, x1, y1 = window or (0, 0, W, H)
    rc = load(which).orc_render(C.byref(scene.c), C.byref(params), C.byref(cfg), accum.ctypes.data_as(C.c_void_p),
                                frame.ctypes.data_as(C.c_void_p), hits.ctypes.data_as(C.c_void_p) if want_hits else None,
                                C.byref(st), C.c_int32(x0), C.c_int32(y0), C.c_int32(x1), C.c_int32(y1))
    return accum, frame, hits, st, rc


def closest_hit(which, scene: OracleScene, org, direction, tmin=0.01, tmax=1e16, use_bvh=1):
    o = (C.c_float * 3)(*map(float, org)); d = (C.c_float * 3)(*map(float, direction))
    t, b1, b2 = C.c_float(), C.c_float(), C.c_float()
    prim = load(which).orc_closest_hit(C.byref(scene.c), o, d, C.c_float(tmin), C.c_float(tmax), C.c_int32(use_bvh),
                                       C.byref(t), C.byref(b1), C.byref(b2))
    return prim, t.value, b1.value, b2.value
