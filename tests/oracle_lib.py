"""ctypes binding of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
The product package (raytracingincuda_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_LIB = None

SLOT_DTYPE = np.dtype([("c", "<f4", 3), ("r", "<f4"), ("type", "<i4"), ("albedo", "<f4", 3),
                       ("fuzz", "<f4"), ("ri", "<f4")])
SLOT64_DTYPE = np.dtype([("c", "<f8", 3), ("r", "<f8"), ("type", "<i4"), ("pad", "<i4"),
                         ("albedo", "<f8", 3), ("fuzz", "<f8"), ("ri", "<f8")])
assert SLOT_DTYPE.itemsize == 40 and SLOT64_DTYPE.itemsize == 80


class Camera(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("scale", C.c_float), ("center", C.c_float * 3), ("pixel00", C.c_float * 3),
                ("du", C.c_float * 3), ("dv", C.c_float * 3), ("defocus_angle", C.c_float),
                ("disk_u", C.c_float * 3), ("disk_v", C.c_float * 3)]


class Camera64(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("scale", C.c_double), ("center", C.c_double * 3), ("pixel00", C.c_double * 3),
                ("du", C.c_double * 3), ("dv", C.c_double * 3), ("defocus_angle", C.c_double),
                ("disk_u", C.c_double * 3), ("disk_v", C.c_double * 3)]


class GlibcRand(C.Structure):
    _fields_ = [("r", C.c_int32 * 34), ("f", C.c_int), ("b", C.c_int)]


def build():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        src_newer = (not os.path.exists(path)) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(path)
            for f in ("rt_oracle.c", "rt_oracle_impl.inc", "rt_oracle.h"))
        if src_newer:
            build()
        L = C.CDLL(path)
        L.orc_scene.restype = C.c_int
        L.orc_scene.argtypes = [C.c_int, C.c_void_p]
        L.orc_scene_scaled.restype = C.c_int
        L.orc_scene_scaled.argtypes = [C.c_int, C.c_void_p]
        L.orc_scene64.restype = C.c_int
        L.orc_scene64.argtypes = [C.c_int, C.c_void_p]
        L.orc_rand.restype = C.c_int
        L.orc_uniform.restype = C.c_float
        L.orc_uniform.argtypes = [C.c_uint32]
        L.orc_num_chunks.restype = C.c_int
        L.orc_quantise.restype = C.c_int
        L.orc_quantise.argtypes = [C.c_float]
        L.orc_hit_world.restype = C.c_int
        L.orc_render.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p]
        L.orc_render64.argtypes = L.orc_render.argtypes
        L.orc_primary.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_primary64.argtypes = L.orc_primary.argtypes
        L.orc_sample.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p]
        L.orc_sample64.argtypes = L.orc_sample.argtypes
        L.orc_fix.restype = C.c_int64
        L.orc_fix.argtypes = [C.c_double]
        L.orc_accumulate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p]
        L.orc_accumulate64.argtypes = L.orc_accumulate.argtypes
        L.orc_finalize.argtypes = [C.c_void_p, C.c_uint64, C.c_float, C.c_void_p]
        L.orc_finalize64.argtypes = [C.c_void_p, C.c_uint64, C.c_double, C.c_void_p]
        _LIB = L
    return _LIB


def scene(scene_id, double=False):
    L = lib()
    fn, dt = (L.orc_scene64, SLOT64_DTYPE) if double else (L.orc_scene, SLOT_DTYPE)
    n = fn(scene_id, None)
    out = np.zeros(n, dtype=dt)
    fn(scene_id, out.ctypes.data)
    return out


def scene_scaled(half):
    L = lib()
    n = L.orc_scene_scaled(half, None)
    out = np.zeros(n, dtype=SLOT_DTYPE)
    L.orc_scene_scaled(half, out.ctypes.data)
    return out


def camera(width, height, spp=10, max_depth=25, double=False):
    L = lib()
    cam = Camera64() if double else Camera()
    (L.orc_camera_init64 if double else L.orc_camera_init)(C.byref(cam), width, height, spp, max_depth)
    return cam


def primary(slots, cam):
    L = lib()
    double = isinstance(cam, Camera64)
    n = cam.width * cam.height
    ids = np.empty(n, dtype=np.int32)
    t = np.empty(n, dtype=np.float64 if double else np.float32)
    (L.orc_primary64 if double else L.orc_primary)(slots.ctypes.data, len(slots), C.byref(cam),
                                                   ids.ctypes.data, t.ctypes.data)
    return ids.reshape(cam.height, cam.width), t.reshape(cam.height, cam.width)


def render(slots, cam, seed=1227, row0=0, row1=None):
    """Gamma-encoded image rows [row0,row1) and the number of hit_world calls (ray segments)."""
    L = lib()
    double = isinstance(cam, Camera64)
    row1 = cam.height if row1 is None else row1
    out = np.empty((row1 - row0, cam.width, 3), dtype=np.float64 if double else np.float32)
    seg = C.c_uint64(0)
    (L.orc_render64 if double else L.orc_render)(slots.ctypes.data, len(slots), C.byref(cam), seed,
                                                 row0, row1, out.ctypes.data, C.byref(seg))
    return out, seg.value


def sample(slots, cam, i, j, s, seed=1227):
    L = lib()
    double = isinstance(cam, Camera64)
    rgb = np.zeros(3, dtype=np.float64 if double else np.float32)
    (L.orc_sample64 if double else L.orc_sample)(slots.ctypes.data, len(slots), C.byref(cam), seed,
                                                 i, j, s, rgb.ctypes.data, None)
    return rgb


def accumulate(slots, cam, s0=0, s1=None, seed=1227, row0=0, row1=None, acc=None):
    """Integer accumulators (rows, width, 3) int64 of samples [s0,s1); adds to `acc` when given."""
    L = lib()
    double = isinstance(cam, Camera64)
    row1 = cam.height if row1 is None else row1
    s1 = cam.spp if s1 is None else s1
    if acc is None:
        acc = np.zeros((row1 - row0, cam.width, 3), dtype=np.int64)
    (L.orc_accumulate64 if double else L.orc_accumulate)(slots.ctypes.data, len(slots), C.byref(cam), seed, row0, row1,
                                                         s0, s1, acc.ctypes.data, None)
    return acc


def finalize(acc, cam):
    """Gamma-encoded frame of integer accumulators."""
    L = lib()
    double = isinstance(cam, Camera64)
    acc = np.ascontiguousarray(acc, dtype=np.int64)
    out = np.empty(acc.shape, dtype=np.float64 if double else np.float32)
    (L.orc_finalize64 if double else L.orc_finalize)(acc.ctypes.data, acc.size // 3, cam.scale, out.ctypes.data)
    return out


def pixel(slots, cam, i, j, seed=1227):
    """One gamma-encoded pixel of the canonical image (all cam.spp samples)."""
    acc = np.zeros(3, dtype=np.int64)
    for s in range(cam.spp):
        rgb = sample(slots, cam, i, j, s, seed)
        for k in range(3):
            acc[k] += lib().orc_fix(float(rgb[k]))
    return finalize(acc.reshape(1, 1, 3), cam)[0, 0]


def num_chunks(width, height, spp):
    return lib().orc_num_chunks(width, height, spp)


def quantise(img):
    """GF main.cu:366-377 on a float image -> uint8 code values."""
    x = np.asarray(img, dtype=np.float32)
    x = np.where(x < np.float32(0.0), np.float32(0.0), x)
    x = np.where(x > np.float32(0.999), np.float32(0.999), x)
    return (np.float32(256) * x).astype(np.int32).astype(np.uint8)
