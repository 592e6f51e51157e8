"""ctypes binding of libptb.so, the C-ABI library declared in include/ptb.h.

This is the Python-side mirror of the reference's render-loop surface
(/root/reference/optixSphere.cpp: createSceneGeometry 400-649, GAS build
860-968, Params 1293-1308, optixLaunch 1403-1418, CUDAOutputBuffer 1284,
saveImage 1483-1489).  It only marshals arguments; all work happens in the
library.  There is no Python or CPU fallback: importing works anywhere, but every
compute call raises PtbError when libptb.so or a CUDA device is missing.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
_ROOT = _PKG.parent
LIB_PATH = Path(os.environ["PTB_LIB"]) if os.environ.get("PTB_LIB") else _PKG / "libptb.so"  # PTB_LIB: A/B builds of the same library (experiments)

PTB_OK, PTB_ERR_INVALID, PTB_ERR_IO, PTB_ERR_CUDA, PTB_ERR_UNSUPPORTED, PTB_ERR_NO_DEVICE = range(6)
PTB_ARITH_EXACT, PTB_ARITH_FAST = 0, 1  # ptb_render_cfg.arith_mode


class PtbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ptb error {code}: {msg}")
        self.code = code


# ---- struct mirrors (layout pinned by tests/test_abi_layout.py) -----------------
class Float2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]


class Float4(C.Structure):
    _pack_ = 16
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]


class TriangleData(C.Structure):
    _fields_ = [(n, C.c_float * 4) for n in ("v0", "v1", "v2", "n0", "n1", "n2")] + [
        (n, C.c_float * 2) for n in ("uv0", "uv1", "uv2")
    ] + [("_tail_pad", C.c_float * 2)]  # float4 members make the struct 16-byte aligned: sizeof == 128


class Params(C.Structure):
    _fields_ = [
        ("image_width", C.c_uint), ("image_height", C.c_uint), ("origin_x", C.c_int), ("origin_y", C.c_int),
        ("subframe_index", C.c_int), ("frame_buffer", C.c_void_p), ("accum_buffer", C.c_void_p), ("dof", C.c_bool),
        ("eye", Float3), ("U", Float3), ("V", Float3), ("W", Float3),
        ("triangles", C.c_void_p), ("num_triangles", C.c_uint), ("handle", C.c_ulonglong),
    ]


class HitGroupData(C.Structure):
    _fields_ = [
        ("albedo_texture_data", C.c_void_p), ("tex_width", C.c_int), ("tex_height", C.c_int), ("has_texture", C.c_bool),
        ("roughness_texture_data", C.c_void_p), ("roughness_width", C.c_int), ("roughness_height", C.c_int), ("has_roughness_map", C.c_bool),
        ("normal_texture_data", C.c_void_p), ("normal_width", C.c_int), ("normal_height", C.c_int), ("has_normal_map", C.c_bool),
        ("metallic_texture_data", C.c_void_p), ("metallic_width", C.c_int), ("metallic_height", C.c_int), ("has_metallic_map", C.c_bool),
        ("texcoords", C.c_void_p), ("vertices", C.c_void_p), ("normals", C.c_void_p),
        ("emission_color", Float3), ("diffuse_color", Float3), ("specular", Float3),
        ("roughness", C.c_float), ("metallic", C.c_bool), ("transparent", C.c_bool),
    ]


class RenderCfg(C.Structure):
    _fields_ = [
        ("spp_per_launch", C.c_int32), ("max_depth", C.c_int32), ("tmin", C.c_float), ("tmax", C.c_float),
        ("dof_blur", C.c_float), ("focus_dist", C.c_float), ("nmap_strength", C.c_float),
        ("exposure", C.c_float), ("gamma", C.c_float), ("contrast", C.c_float),
        ("accumulate_mode", C.c_int32), ("write_frame", C.c_int32), ("env_importance_sampling", C.c_int32),
        ("count_traversal", C.c_int32), ("profile_stages", C.c_int32), ("subframes_per_launch", C.c_int32), ("pipeline", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32), ("row_interleave_count", C.c_int32), ("row_interleave_index", C.c_int32),
        ("row_interleave_height", C.c_int32), ("aux_primary_hit", C.c_void_p), ("chunk_slots_per_thread", C.c_int32), ("arith_mode", C.c_int32), ("max_pool_bytes", C.c_int64),
        ("overlap_lanes", C.c_int32), ("reserved0", C.c_int32),
    ]


class LaunchStats(C.Structure):
    _fields_ = [
        ("segments", C.c_uint64), ("paths", C.c_uint64), ("hits", C.c_uint64), ("misses", C.c_uint64),
        ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("iterations", C.c_uint32), ("kernel_launches", C.c_uint32),
    ]


class BuildCfg(C.Structure):
    _fields_ = [("max_leaf_size", C.c_int32), ("sah_refine", C.c_int32), ("sah_bins", C.c_int32), ("treelet_size", C.c_int32),
                ("morton_bits", C.c_int32), ("bvh_width", C.c_int32)]


class BuildStats(C.Structure):
    _fields_ = [
        ("num_triangles", C.c_uint32), ("num_nodes", C.c_uint32), ("num_leaves", C.c_uint32), ("max_depth", C.c_uint32),
        ("sah_cost", C.c_float), ("build_ms", C.c_float), ("bvh_bytes", C.c_uint64), ("bvh_width", C.c_uint32), ("sah_cost_mesh", C.c_float),
        ("num_nodes8", C.c_uint32),
    ]


class MaterialInfo(C.Structure):
    _fields_ = [
        ("emission_color", C.c_float * 3), ("diffuse_color", C.c_float * 3), ("specular", C.c_float * 3),
        ("roughness", C.c_float), ("metallic", C.c_int32), ("transparent", C.c_int32),
        ("has_albedo", C.c_int32), ("albedo_w", C.c_int32), ("albedo_h", C.c_int32),
        ("has_roughness", C.c_int32), ("roughness_w", C.c_int32), ("roughness_h", C.c_int32),
        ("has_normal", C.c_int32), ("normal_w", C.c_int32), ("normal_h", C.c_int32),
        ("has_metallic", C.c_int32), ("metallic_w", C.c_int32), ("metallic_h", C.c_int32),
    ]


# Every symbol include/ptb.h declares (tests check that the library exports all of them).
EXPORTS = [
    "ptb_last_error", "ptb_version", "ptb_context_create", "ptb_context_destroy", "ptb_context_synchronize",
    "ptb_scene_load_obj", "ptb_scene_create", "ptb_scene_create_demo", "ptb_scene_set_materials", "ptb_scene_set_env_file",
    "ptb_scene_set_env_pixels", "ptb_scene_destroy", "ptb_scene_num_triangles", "ptb_scene_num_materials",
    "ptb_scene_copy_triangles", "ptb_scene_copy_material_ids", "ptb_scene_get_material", "ptb_scene_copy_texture",
    "ptb_scene_env_size", "ptb_scene_copy_env", "ptb_default_build_cfg", "ptb_accel_build", "ptb_accel_read", "ptb_accel_read8",
    "ptb_camera_uvw", "ptb_params_default_camera", "ptb_default_render_cfg", "ptb_launch", "ptb_launch_get_stats", "ptb_launch_get_stage_ms", "ptb_context_get_totals",
    "ptb_resolve", "ptb_resolve_peers", "ptb_resolve_peers_accumulate", "ptb_resolve_peers_sync", "ptb_peer_flags_create", "ptb_peer_signal", "ptb_peer_wait", "ptb_peer_flags_error", "ptb_multi_create", "ptb_multi_destroy", "ptb_multi_device_count", "ptb_multi_context", "ptb_multi_stream", "ptb_multi_accel_build", "ptb_multi_launch", "ptb_multi_synchronize", "ptb_multi_get_totals", "ptb_ipc_export", "ptb_ipc_open", "ptb_ipc_close", "ptb_trace_rays", "ptb_output_create", "ptb_output_resize", "ptb_output_map", "ptb_output_unmap",
    "ptb_output_host_ptr", "ptb_output_width", "ptb_output_height", "ptb_output_destroy", "ptb_device_alloc",
    "ptb_device_free", "ptb_device_memset", "ptb_copy_to_device", "ptb_copy_to_host", "ptb_copy_to_host_async", "ptb_image_load_rgba8",
    "ptb_image_load_float4", "ptb_save_image", "ptb_save_accum_raw", "ptb_load_accum_raw", "ptb_free", "ptb_obj_read", "ptb_microbench_read", "ptb_test_env_sample", "ptb_test_device_math",
]

_lib = None


def build(verbose: bool = False) -> Path:
    """Compile libptb.so in-tree with nvcc for sm_100a (szakdolgozat_pathtracer_b200/csrc/Makefile)."""
    r = subprocess.run(["make", "-C", str(_PKG / "csrc")], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libptb.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    """Load libptb.so.  Fails loudly when it has not been built: there is no fallback."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise PtbError(-1, f"{LIB_PATH} is missing; run `python -c 'import __graft_entry__ as g; g.build()'` (nvcc, sm_100a)")
        L = C.CDLL(str(LIB_PATH))
        L.ptb_last_error.restype = C.c_char_p
        L.ptb_version.restype = C.c_char_p
        L.ptb_scene_num_triangles.restype = C.c_uint32
        L.ptb_output_map.restype = C.c_void_p
        L.ptb_output_host_ptr.restype = C.c_void_p
        L.ptb_output_width.restype = C.c_uint32
        L.ptb_output_height.restype = C.c_uint32
        for name in ("ptb_context_destroy", "ptb_scene_destroy", "ptb_output_destroy", "ptb_output_unmap", "ptb_free",
                     "ptb_camera_uvw", "ptb_params_default_camera", "ptb_default_render_cfg", "ptb_default_build_cfg"):
            getattr(L, name).restype = None
        L.ptb_context_destroy.argtypes = [C.c_void_p]
        L.ptb_multi_destroy.restype = None
        L.ptb_multi_destroy.argtypes = [C.c_void_p]
        L.ptb_scene_destroy.argtypes = [C.c_void_p]
        L.ptb_output_destroy.argtypes = [C.c_void_p]
        L.ptb_free.argtypes = [C.c_void_p]
        L.ptb_scene_num_triangles.argtypes = [C.c_void_p]
        L.ptb_scene_num_materials.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != PTB_OK:
        raise PtbError(rc, lib().ptb_last_error().decode(errors="replace"))


def _fptr(a):
    return a.ctypes.data_as(C.c_void_p)


def default_render_cfg(**kw) -> RenderCfg:
    cfg = RenderCfg()
    lib().ptb_default_render_cfg(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def default_build_cfg(**kw) -> BuildCfg:
    cfg = BuildCfg()
    lib().ptb_default_build_cfg(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def camera_uvw(eye, lookat, up, fovy_deg, aspect):
    """sutil::Camera::UVWFrame (optixSphere.cpp:238-247)."""
    e, l, u = (np.asarray(v, np.float32) for v in (eye, lookat, up))
    U, V, W = (np.zeros(3, np.float32) for _ in range(3))
    lib().ptb_camera_uvw(_fptr(e), _fptr(l), _fptr(u), C.c_float(fovy_deg), C.c_float(aspect), _fptr(U), _fptr(V), _fptr(W))
    return U, V, W


def make_params(width, height, subframe_index=0, dof=True, eye=(0.0, 2.0, 6.0), lookat=(0.0, 0.0, 0.0),
                up=(0.0, 1.0, 0.0), fovy_deg=50.0) -> Params:
    """Params as main() fills it (optixSphere.cpp:1293-1308) with the reference camera (102-120)."""
    p = Params()
    p.image_width, p.image_height = width, height
    p.origin_x, p.origin_y = width // 2, height // 2
    p.subframe_index = subframe_index
    p.dof = bool(dof)
    U, V, W = camera_uvw(eye, lookat, up, fovy_deg, width / height)
    p.eye = Float3(*[float(x) for x in eye])
    p.U, p.V, p.W = Float3(*map(float, U)), Float3(*map(float, V)), Float3(*map(float, W))
    return p


class Scene:
    """createSceneGeometry + materials + environment (host side; needs no GPU)."""

    def __init__(self, handle):
        self._h = C.c_void_p(handle)
        self._keep = []

    @classmethod
    def load_obj(cls, files, scale=1.0, material_seed=0):
        arr = (C.c_char_p * len(files))(*[os.fsencode(str(f)) for f in files])
        h = C.c_void_p()
        _check(lib().ptb_scene_load_obj(arr, len(files), C.c_float(scale), C.c_uint32(material_seed), C.byref(h)))
        return cls(h.value)

    @classmethod
    def demo(cls):
        """The reference's procedural scene (optixSphere.cpp:650-751): ground quad + three UV spheres."""
        h = C.c_void_p()
        _check(lib().ptb_scene_create_demo(C.byref(h)))
        return cls(h.value)

    @classmethod
    def from_triangles(cls, tris: np.ndarray, mat_ids: np.ndarray | None = None):
        """tris: float32 [N, 32] laid out as TriangleData (optixSphere.h:2-7)."""
        t = np.ascontiguousarray(tris, np.float32).reshape(-1, 32)
        m = None if mat_ids is None else np.ascontiguousarray(mat_ids, np.uint32)
        h = C.c_void_p()
        _check(lib().ptb_scene_create(_fptr(t), C.c_uint32(t.shape[0]), None if m is None else _fptr(m), C.byref(h)))
        return cls(h.value)

    def set_materials(self, mats):
        """mats: list of dicts with emission_color/diffuse_color/specular/roughness/metallic and optional
        albedo/roughness_map/normal_map/metallic_map float32 [h, w, 4] arrays (the reference's float4 textures)."""
        arr = (HitGroupData * len(mats))()
        keep = []
        for i, m in enumerate(mats):
            h = arr[i]
            h.emission_color = Float3(*map(float, m.get("emission_color", (0, 0, 0))))
            h.diffuse_color = Float3(*map(float, m.get("diffuse_color", (0.5, 0.5, 0.5))))
            h.specular = Float3(*map(float, m.get("specular", (0.5, 0.5, 0.5))))
            h.roughness = float(m.get("roughness", 0.4))
            h.metallic = bool(m.get("metallic", False))
            h.transparent = bool(m.get("transparent", False))
            for key, ptr, wn, hn, flag in (("albedo", "albedo_texture_data", "tex_width", "tex_height", "has_texture"),
                                           ("roughness_map", "roughness_texture_data", "roughness_width", "roughness_height", "has_roughness_map"),
                                           ("normal_map", "normal_texture_data", "normal_width", "normal_height", "has_normal_map"),
                                           ("metallic_map", "metallic_texture_data", "metallic_width", "metallic_height", "has_metallic_map")):
                t = m.get(key)
                if t is not None:
                    t = np.ascontiguousarray(t, np.float32)
                    keep.append(t)
                    setattr(h, ptr, t.ctypes.data)
                    setattr(h, wn, t.shape[1]); setattr(h, hn, t.shape[0]); setattr(h, flag, True)
        _check(lib().ptb_scene_set_materials(self._h, arr, len(mats)))

    def set_env_file(self, path):
        _check(lib().ptb_scene_set_env_file(self._h, os.fsencode(str(path))))

    def set_env_pixels(self, rgba: np.ndarray):
        a = np.ascontiguousarray(rgba, np.float32)
        _check(lib().ptb_scene_set_env_pixels(self._h, _fptr(a), a.shape[1], a.shape[0]))

    # host read-back
    @property
    def num_triangles(self):
        return int(lib().ptb_scene_num_triangles(self._h))

    @property
    def num_materials(self):
        return int(lib().ptb_scene_num_materials(self._h))

    def triangles(self) -> np.ndarray:
        out = np.zeros((self.num_triangles, 32), np.float32)
        _check(lib().ptb_scene_copy_triangles(self._h, _fptr(out), C.c_uint32(out.shape[0])))
        return out

    def material_ids(self) -> np.ndarray:
        out = np.zeros(self.num_triangles, np.uint32)
        _check(lib().ptb_scene_copy_material_ids(self._h, _fptr(out), C.c_uint32(out.shape[0])))
        return out

    def material(self, i) -> MaterialInfo:
        mi = MaterialInfo()
        _check(lib().ptb_scene_get_material(self._h, i, C.byref(mi)))
        return mi

    def texture(self, material, kind) -> np.ndarray:
        mi = self.material(material)
        w, h = [(mi.albedo_w, mi.albedo_h), (mi.roughness_w, mi.roughness_h), (mi.normal_w, mi.normal_h), (mi.metallic_w, mi.metallic_h)][kind]
        out = np.zeros((h, w, 4), np.float32)
        _check(lib().ptb_scene_copy_texture(self._h, material, kind, _fptr(out), C.c_size_t(out.size)))
        return out

    def env(self) -> np.ndarray:
        w, h = C.c_int(), C.c_int()
        _check(lib().ptb_scene_env_size(self._h, C.byref(w), C.byref(h)))
        out = np.zeros((h.value, w.value, 4), np.float32)
        _check(lib().ptb_scene_copy_env(self._h, _fptr(out), C.c_size_t(out.size)))
        return out

    def close(self):
        if self._h:
            lib().ptb_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """optixDeviceContextCreate .. optixLaunch (optixSphere.cpp:798-812, 1403-1418) on one GPU."""

    def __init__(self, device=0):
        h = C.c_void_p()
        _check(lib().ptb_context_create(device, C.byref(h)))
        self._h = h
        self.device = device

    def accel_build(self, scene: Scene, cfg: BuildCfg | None = None, stream=0):
        handle, st = C.c_ulonglong(), BuildStats()
        _check(lib().ptb_accel_build(self._h, scene._h, C.byref(cfg) if cfg is not None else None, C.c_void_p(stream),
                                     C.byref(handle), C.byref(st)))
        return handle.value, st

    def accel_read(self, handle):
        nn, nt = C.c_uint32(), C.c_uint32()
        _check(lib().ptb_accel_read(self._h, C.c_ulonglong(handle), None, 0, None, 0, C.byref(nn), C.byref(nt)))
        nodes = np.zeros((nn.value, 16), np.float32)
        tris = np.zeros((nt.value, 12), np.float32)
        _check(lib().ptb_accel_read(self._h, C.c_ulonglong(handle), _fptr(nodes), nn, _fptr(tris), nt, None, None))
        return nodes, tris

    def accel_read8(self, handle):
        """The 8-wide quantised tree: (nodes8 as uint32 [n, 20], tris8 as float32 [n_tris, 12]); empty arrays without one."""
        nn, nt = C.c_uint32(), C.c_uint32()
        _check(lib().ptb_accel_read8(self._h, C.c_ulonglong(handle), None, 0, None, 0, C.byref(nn), C.byref(nt)))
        nodes = np.zeros((nn.value, 20), np.uint32)
        tris = np.zeros((nt.value, 12), np.float32)
        if nn.value:
            _check(lib().ptb_accel_read8(self._h, C.c_ulonglong(handle), nodes.ctypes.data_as(C.POINTER(C.c_uint32)), nn, _fptr(tris), nt, None, None))
        return nodes, tris

    def launch(self, params: Params, cfg: RenderCfg | None = None, stream=0):
        _check(lib().ptb_launch(self._h, C.byref(params), C.byref(cfg) if cfg is not None else None, C.c_void_p(stream)))

    def launch_stats(self) -> LaunchStats:
        st = LaunchStats()
        _check(lib().ptb_launch_get_stats(self._h, C.byref(st)))
        return st

    def totals(self, reset=False) -> dict:
        out = (C.c_uint64 * 4)()
        _check(lib().ptb_context_get_totals(self._h, out, int(bool(reset))))
        return dict(zip(("segments", "hits", "misses", "launches"), [int(x) for x in out]))

    def stage_ms(self) -> dict:
        out = (C.c_float * 6)()
        _check(lib().ptb_launch_get_stage_ms(self._h, out))
        return dict(zip(("raygen", "trace", "shade", "miss", "resolve", "total"), [float(x) for x in out]))

    def resolve(self, accum_ptr, accum_out_ptr, frame_ptr, n_pixels, scale, cfg: RenderCfg | None = None, stream=0):
        _check(lib().ptb_resolve(self._h, C.c_void_p(accum_ptr), C.c_void_p(accum_out_ptr), C.c_void_p(frame_ptr), C.c_uint32(n_pixels),
                                 C.c_float(scale), C.byref(cfg) if cfg is not None else None, C.c_void_p(stream)))

    def resolve_peers(self, accum_ptrs, accum_out_ptr, frame_ptr, first_pixel, n_pixels, scale, cfg: RenderCfg | None = None, stream=0):
        """Fused reduce -> tonemap -> gather over local/peer accumulators (one kernel, no NCCL on the data path)."""
        arr = (C.c_void_p * len(accum_ptrs))(*accum_ptrs)
        _check(lib().ptb_resolve_peers(self._h, arr, len(accum_ptrs), C.c_void_p(accum_out_ptr), C.c_void_p(frame_ptr), C.c_uint32(first_pixel),
                                       C.c_uint32(n_pixels), C.c_float(scale), C.byref(cfg) if cfg is not None else None, C.c_void_p(stream)))

    def peer_flags_create(self) -> int:
        p = C.c_void_p()
        _check(lib().ptb_peer_flags_create(self._h, C.byref(p)))
        return p.value

    def peer_signal(self, flag_blocks, my_rank, kind, epoch, stream=0):
        arr = (C.c_void_p * len(flag_blocks))(*flag_blocks)
        _check(lib().ptb_peer_signal(self._h, arr, len(flag_blocks), int(my_rank), int(kind), C.c_uint32(epoch), C.c_void_p(stream)))

    def peer_wait(self, my_flags, kind, n_ranks, epoch, stream=0):
        _check(lib().ptb_peer_wait(self._h, C.c_void_p(my_flags), int(kind), int(n_ranks), C.c_uint32(epoch), C.c_void_p(stream)))

    def peer_flags_error(self, flags, stream=0) -> bool:
        e = C.c_int()
        _check(lib().ptb_peer_flags_error(self._h, C.c_void_p(flags), C.c_void_p(stream), C.byref(e)))
        return bool(e.value)

    def resolve_peers_sync(self, accum_ptrs, my_rank, my_flags, root_flags, epoch, accum_out_ptr, frame_ptr, first_pixel, n_pixels, scale,
                           cfg: RenderCfg | None = None, stream=0, prev_accum=0, prev_weight=0.0):
        """ptb_resolve_peers_sync: waits on the device for all ranks' arrive signals, reduces + tonemaps this rank's slice, signals done."""
        arr = (C.c_void_p * len(accum_ptrs))(*accum_ptrs)
        _check(lib().ptb_resolve_peers_sync(self._h, arr, len(accum_ptrs), int(my_rank), C.c_void_p(my_flags), C.c_void_p(root_flags), C.c_uint32(epoch),
                                            C.c_void_p(prev_accum), C.c_float(prev_weight), C.c_void_p(accum_out_ptr), C.c_void_p(frame_ptr),
                                            C.c_uint32(first_pixel), C.c_uint32(n_pixels), C.c_float(scale), C.byref(cfg) if cfg is not None else None,
                                            C.c_void_p(stream)))

    def ipc_export(self, ptr) -> bytes:
        h = (C.c_ubyte * 64)()
        _check(lib().ptb_ipc_export(self._h, C.c_void_p(ptr), h))
        return bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        h = (C.c_ubyte * 64)(*handle)
        p = C.c_void_p()
        _check(lib().ptb_ipc_open(self._h, h, C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        _check(lib().ptb_ipc_close(self._h, C.c_void_p(ptr)))

    def synchronize(self, stream=0):
        _check(lib().ptb_context_synchronize(self._h, C.c_void_p(stream)))

    # device memory
    def alloc(self, nbytes) -> int:
        p = C.c_void_p()
        _check(lib().ptb_device_alloc(self._h, C.c_size_t(nbytes), C.byref(p)))
        return p.value

    def free(self, ptr):
        _check(lib().ptb_device_free(self._h, C.c_void_p(ptr)))

    def memset(self, ptr, value, nbytes, stream=0):
        _check(lib().ptb_device_memset(self._h, C.c_void_p(ptr), value, C.c_size_t(nbytes), C.c_void_p(stream)))

    def to_device(self, ptr, arr: np.ndarray, stream=0):
        a = np.ascontiguousarray(arr)
        _check(lib().ptb_copy_to_device(self._h, C.c_void_p(ptr), _fptr(a), C.c_size_t(a.nbytes), C.c_void_p(stream)))
        self.synchronize(stream)

    def to_host(self, ptr, shape, dtype, stream=0) -> np.ndarray:
        out = np.zeros(shape, dtype)
        _check(lib().ptb_copy_to_host(self._h, _fptr(out), C.c_void_p(ptr), C.c_size_t(out.nbytes), C.c_void_p(stream)))
        return out

    def trace_rays(self, handle, origins, dirs, tmin=0.01, tmax=1e16):
        o = np.ascontiguousarray(origins, np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(dirs, np.float32).reshape(-1, 3)
        n = o.shape[0]
        bufs = [self.alloc(max(n, 1) * 12), self.alloc(max(n, 1) * 12)] + [self.alloc(max(n, 1) * 4) for _ in range(4)]
        try:
            self.to_device(bufs[0], o); self.to_device(bufs[1], d)
            _check(lib().ptb_trace_rays(self._h, C.c_ulonglong(handle), C.c_void_p(bufs[0]), C.c_void_p(bufs[1]), C.c_uint32(n),
                                        C.c_float(tmin), C.c_float(tmax), C.c_void_p(bufs[2]), C.c_void_p(bufs[3]),
                                        C.c_void_p(bufs[4]), C.c_void_p(bufs[5]), None))
            prim = self.to_host(bufs[2], n, np.int32)
            t = self.to_host(bufs[3], n, np.float32)
            b1 = self.to_host(bufs[4], n, np.float32)
            b2 = self.to_host(bufs[5], n, np.float32)
        finally:
            for b in bufs:
                self.free(b)
        return prim, t, b1, b2

    def microbench_read(self, nbytes, iters=20) -> float:
        """Read bandwidth in GB/s over a buffer of nbytes (<= 32 MiB: L2-resident; >> 126 MiB: HBM)."""
        out = C.c_double()
        _check(lib().ptb_microbench_read(self._h, C.c_size_t(nbytes), iters, C.byref(out)))
        return out.value

    def test_env_sample(self, handle, xi: np.ndarray) -> np.ndarray:
        """Samples of the scene's environment CDF: xi [n, 2] uniforms -> [n, 4] (direction, solid-angle pdf)."""
        x = np.ascontiguousarray(xi, np.float32).reshape(-1, 2)
        out = np.zeros((x.shape[0], 4), np.float32)
        _check(lib().ptb_test_env_sample(self._h, C.c_ulonglong(handle), _fptr(x), C.c_uint32(x.shape[0]), _fptr(out)))
        return out

    def test_device_math(self, op, inp: np.ndarray, out_stride) -> np.ndarray:
        a = np.ascontiguousarray(inp, np.float32)
        a2 = a.reshape(a.shape[0], -1)
        out = np.zeros((a2.shape[0], out_stride), np.float32)
        _check(lib().ptb_test_device_math(self._h, op, _fptr(a2), a2.shape[1], _fptr(out), out_stride, C.c_uint32(a2.shape[0])))
        return out

    def close(self):
        if self._h:
            lib().ptb_context_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class OutputBuffer:
    """sutil::CUDAOutputBuffer<uchar4> (optixSphere.cpp:1284, 1401, 1419, 1484-1486)."""

    def __init__(self, ctx: Context, width, height):
        h = C.c_void_p()
        _check(lib().ptb_output_create(ctx._h, C.c_uint32(width), C.c_uint32(height), C.byref(h)))
        self._h, self.width, self.height = h, width, height

    def map(self) -> int:
        return lib().ptb_output_map(self._h)

    def unmap(self, stream=0):
        lib().ptb_output_unmap(self._h, C.c_void_p(stream))

    def host(self) -> np.ndarray:
        p = lib().ptb_output_host_ptr(self._h)
        if not p:
            raise PtbError(PTB_ERR_CUDA, lib().ptb_last_error().decode())
        buf = (C.c_uint8 * (self.width * self.height * 4)).from_address(p)
        return np.frombuffer(buf, np.uint8).reshape(self.height, self.width, 4).copy()

    def close(self):
        if self._h:
            lib().ptb_output_destroy(self._h)
            self._h = None


def load_image_rgba8(path) -> np.ndarray:
    px, w, h = C.POINTER(C.c_uint8)(), C.c_int(), C.c_int()
    _check(lib().ptb_image_load_rgba8(os.fsencode(str(path)), C.byref(px), C.byref(w), C.byref(h)))
    out = np.ctypeslib.as_array(px, shape=(h.value, w.value, 4)).copy()
    lib().ptb_free(px)
    return out


def load_image_float4(path) -> np.ndarray:
    px, w, h = C.POINTER(C.c_float)(), C.c_int(), C.c_int()
    _check(lib().ptb_image_load_float4(os.fsencode(str(path)), C.byref(px), C.byref(w), C.byref(h)))
    out = np.ctypeslib.as_array(px, shape=(h.value, w.value, 4)).copy()
    lib().ptb_free(px)
    return out


def obj_read(path) -> np.ndarray:
    """Raw face-vertex stream of the OBJ reader: uint32 words [n, 10] (8 float32 + 2 int32 flags per face vertex)."""
    rec, n = C.c_void_p(), C.c_uint64()
    _check(lib().ptb_obj_read(os.fsencode(str(path)), C.byref(rec), C.byref(n)))
    buf = (C.c_uint32 * (n.value * 10)).from_address(rec.value)
    out = np.frombuffer(buf, np.uint32).reshape(-1, 10).copy()
    lib().ptb_free(rec)
    return out


PTB_SPLIT_SAMPLES, PTB_SPLIT_TILES = 0, 1


class _BorrowedContext(Context):
    """A context owned by a Multi (never destroyed from Python)."""

    def __init__(self, handle, device):
        self._h, self.device = handle, device

    def close(self):
        self._h = None

    def __del__(self):
        pass


class Multi:
    """ptb_multi: one host process drives n devices (include/ptb.h; csrc/multi.cpp).  contexts[0] is the root: the
    accumulator / frame buffers of Params live there."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(lib().ptb_multi_create(devs, len(devices), C.byref(h)))
        self._h = h
        L = lib()
        L.ptb_multi_context.restype = C.c_void_p
        L.ptb_multi_context.argtypes = [C.c_void_p, C.c_int]
        self.contexts = [_BorrowedContext(C.c_void_p(L.ptb_multi_context(h, i)), d) for i, d in enumerate(devices)]
        self.root = self.contexts[0]

    def accel_build(self, scene: Scene, cfg: BuildCfg | None = None) -> BuildStats:
        st = BuildStats()
        _check(lib().ptb_multi_accel_build(self._h, scene._h, C.byref(cfg) if cfg is not None else None, C.byref(st)))
        return st

    def launch(self, params: Params, cfg: RenderCfg | None = None, split=PTB_SPLIT_SAMPLES):
        _check(lib().ptb_multi_launch(self._h, C.byref(params), C.byref(cfg) if cfg is not None else None, int(split)))

    def synchronize(self):
        _check(lib().ptb_multi_synchronize(self._h))

    def totals(self, reset=False) -> dict:
        out = (C.c_uint64 * 4)()
        _check(lib().ptb_multi_get_totals(self._h, out, int(bool(reset))))
        return dict(zip(("segments", "hits", "misses", "launches"), [int(x) for x in out]))

    def close(self):
        if self._h:
            for c in self.contexts:
                c.close()
            lib().ptb_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def save_image(path, rgba: np.ndarray, flip_y=True):
    a = np.ascontiguousarray(rgba, np.uint8)
    _check(lib().ptb_save_image(os.fsencode(str(path)), _fptr(a), a.shape[1], a.shape[0], int(bool(flip_y))))


def save_accum_raw(path, accum: np.ndarray):
    """Host copy of the float4 accumulation buffer -> raw 'PTBA' file (ptb_save_accum_raw)."""
    a = np.ascontiguousarray(accum, np.float32)
    assert a.ndim == 3 and a.shape[2] == 4
    _check(lib().ptb_save_accum_raw(os.fsencode(str(path)), _fptr(a), a.shape[1], a.shape[0]))


def load_accum_raw(path) -> np.ndarray:
    px, w, h = C.c_void_p(), C.c_int(), C.c_int()
    _check(lib().ptb_load_accum_raw(os.fsencode(str(path)), C.byref(px), C.byref(w), C.byref(h)))
    try:
        return np.ctypeslib.as_array(C.cast(px, C.POINTER(C.c_float)), shape=(h.value, w.value, 4)).copy()
    finally:
        lib().ptb_free(px)
