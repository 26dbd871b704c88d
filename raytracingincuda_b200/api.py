"""ctypes binding of librt_b200.so (include/rt_b200.h) -- host-side mirror of the reference's
`main.cu` flow: build scene -> upload -> render -> write PPM (GF main.cu:142-379).

No CPU fallback: if the shared library is missing this module raises at import time.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# RT_B200_LIB selects another build of the same library (kernel tuning variants); never a fallback
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(_HERE, "librt_b200.so")

SPLIT_NONE, SPLIT_ROWS, SPLIT_SPP = 0, 1, 2
LAMBERTIAN, METAL, DIELECTRIC = 0, 1, 2

SLOT_DTYPE = np.dtype([("c", "<f4", 3), ("r", "<f4"), ("type", "<i4"), ("albedo", "<f4", 3),
                       ("fuzz", "<f4"), ("ri", "<f4")])
SLOT64_DTYPE = np.dtype([("c", "<f8", 3), ("r", "<f8"), ("type", "<i4"), ("pad", "<i4"),
                         ("albedo", "<f8", 3), ("fuzz", "<f8"), ("ri", "<f8")])
Slot, Slot64 = SLOT_DTYPE, SLOT64_DTYPE


class RtError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib().rt_error_string(code).decode()
        if code > 0:
            msg = f"CUDA error {code}"
        super().__init__(f"{where}: {msg} (rc={code})")


class Camera(C.Structure):
    """Device-visible fields of the reference `camera` (GF camera.h:10-31)."""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("scale", C.c_float), ("center", C.c_float * 3), ("pixel00", C.c_float * 3),
                ("du", C.c_float * 3), ("dv", C.c_float * 3), ("defocus_angle", C.c_float),
                ("disk_u", C.c_float * 3), ("disk_v", C.c_float * 3)]


class Camera64(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("scale", C.c_double), ("center", C.c_double * 3), ("pixel00", C.c_double * 3),
                ("du", C.c_double * 3), ("dv", C.c_double * 3), ("defocus_angle", C.c_double),
                ("disk_u", C.c_double * 3), ("disk_v", C.c_double * 3)]


class Opts(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("split", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
                ("tile_rows", C.c_int32), ("accel", C.c_int32), ("threads", C.c_int32), ("kernel", C.c_int32),
                ("place_rows", C.c_int32), ("primary_bins", C.c_int32), ("reserved", C.c_int32 * 5)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("segments", C.c_uint64), ("sphere_tests", C.c_uint64),
                ("node_visits", C.c_uint64), ("render_ms", C.c_float), ("trace_ms", C.c_float),
                ("launches", C.c_int32), ("chunks", C.c_int32), ("grid", C.c_int32), ("block", C.c_int32),
                ("regs", C.c_int32), ("smem_bytes", C.c_int32), ("binned_segments", C.c_uint64),
                ("filter_tests", C.c_uint64), ("bvh_build_ms", C.c_float), ("grid_build_ms", C.c_float),
                ("accel_used", C.c_int32), ("reserved", C.c_int32)]


# every symbol include/rt_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "rt_abi_version": (C.c_int, []),
    "rt_kernel_build_id": (C.c_char_p, []),
    "rt_scene_generate": (C.c_int, [C.c_int, _P, C.c_int]),
    "rt_scene_generate64": (C.c_int, [C.c_int, _P, C.c_int]),
    "rt_scene_generate_scaled": (C.c_int, [C.c_int, _P, C.c_int]),
    "rt_camera_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "rt_camera_init64": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "rt_opts_default": (None, [_P]),
    "rt_num_chunks": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "rt_partition_rows": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_int]),
    "rt_partition_samples": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P]),
    "rt_ppm_write": (C.c_int, [C.c_char_p, _P, C.c_int, C.c_int]),
    "rt_ppm_write64": (C.c_int, [C.c_char_p, _P, C.c_int, C.c_int]),
    "rt_ppm_quantise": (C.c_int, [_P, C.c_size_t, _P]),
    "rt_error_string": (C.c_char_p, [C.c_int]),
    "rt_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "rt_destroy": (C.c_int, [_P]),
    "rt_set_stream": (C.c_int, [_P, _P]),
    "rt_upload_scene": (C.c_int, [_P, _P, C.c_int]),
    "rt_upload_scene64": (C.c_int, [_P, _P, C.c_int]),
    "rt_render": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_float)]),
    "rt_render64": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_float)]),
    "rt_render_partials": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_float)]),
    "rt_finalize": (C.c_int, [_P, _P, _P, _P, C.POINTER(C.c_float)]),
    "rt_finalize_sum": (C.c_int, [_P, _P, _P, C.c_int, _P, C.POINTER(C.c_float)]),
    "rt_primary_hits": (C.c_int, [_P, _P, _P, _P]),
    "rt_primary_hits64": (C.c_int, [_P, _P, _P, _P]),
    "rt_primary_hits_accel": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "rt_primary_hits_accel64": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "rt_filter_audit": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, _P]),
    "rt_scene_read_text": (C.c_int, [C.c_char_p, _P, C.c_int]),
    "rt_scene_write_text": (C.c_int, [C.c_char_p, _P, C.c_int]),
    "rt_get_stats": (C.c_int, [_P, _P]),
    "rt_debug_checks": (C.c_int, [_P, _P, _P, _P, C.c_int]),
    "rt_frame_alloc": (C.c_int, [_P, C.c_size_t, C.POINTER(_P)]),
    "rt_frame_free": (C.c_int, [_P, _P]),
    "rt_frame_read": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "rt_frame_write": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "rt_enable_peer_access": (C.c_int, [_P, C.c_int]),
    "rt_copy_peer": (C.c_int, [_P, _P, _P, C.c_int, C.c_size_t]),
}

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C raytracingincuda_b200/csrc` "
                "(or __graft_entry__.build()). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)          # AttributeError if the ABI and the header drifted apart
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _ck(rc, where):
    if rc != 0:
        raise RtError(rc, where)


# ------------------------------------------------------------------ host side ---------------
def scene(scene_id, double=False):
    """The reference's scene (GF/GD main.cu:142-298) as a structured array of slots."""
    L = lib()
    fn, dt = (L.rt_scene_generate64, SLOT64_DTYPE) if double else (L.rt_scene_generate, SLOT_DTYPE)
    n = fn(scene_id, None, 0)
    out = np.zeros(n, dtype=dt)
    fn(scene_id, out.ctypes.data, n)
    return out


def scene_scaled(half):
    L = lib()
    n = L.rt_scene_generate_scaled(half, None, 0)
    if n < 0:
        raise RtError(n, "rt_scene_generate_scaled")
    out = np.zeros(n, dtype=SLOT_DTYPE)
    L.rt_scene_generate_scaled(half, out.ctypes.data, n)
    return out


def load_scene(path):
    """rt_scene_read_text: slots of a text scene file (one sphere per line)."""
    n = lib().rt_scene_read_text(os.fsencode(path), None, 0)
    if n < 0:
        raise RtError(n, "rt_scene_read_text")
    slots = np.zeros(n, dtype=SLOT_DTYPE)
    lib().rt_scene_read_text(os.fsencode(path), slots.ctypes.data, n)
    return slots


def save_scene(path, slots):
    slots = np.ascontiguousarray(slots, dtype=SLOT_DTYPE)
    _ck(lib().rt_scene_write_text(os.fsencode(path), slots.ctypes.data, len(slots)), "rt_scene_write_text")


def camera(width, height, spp=10, max_depth=25, double=False):
    """camera::initialize() (GF camera.h:33-68) for the reference's fixed view."""
    cam = Camera64() if double else Camera()
    fn = lib().rt_camera_init64 if double else lib().rt_camera_init
    _ck(fn(C.byref(cam), width, height, spp, max_depth), "rt_camera_init")
    return cam


def num_chunks(width, height, spp):
    return lib().rt_num_chunks(width, height, spp)


def partition_rows(height, tile_rows, rank, world):
    n = lib().rt_partition_rows(height, tile_rows, rank, world, None, 0)
    if n < 0:
        raise RtError(n, "rt_partition_rows")
    rows = np.zeros(n, dtype=np.int32)
    lib().rt_partition_rows(height, tile_rows, rank, world, rows.ctypes.data, n)
    return rows


def partition_samples(spp, rank, world):
    """Samples [s0, s1) of every pixel that `rank` of `world` renders under SPLIT_SPP."""
    s0, s1 = C.c_int32(), C.c_int32()
    _ck(lib().rt_partition_samples(spp, rank, world, C.byref(s0), C.byref(s1)), "rt_partition_samples")
    return s0.value, s1.value


def ppm_write(path, rgb):
    rgb = np.ascontiguousarray(rgb)
    h, w, _ = rgb.shape
    if rgb.dtype == np.float64:
        _ck(lib().rt_ppm_write64(os.fsencode(path), rgb.ctypes.data, w, h), "rt_ppm_write64")
    else:
        rgb = rgb.astype(np.float32, copy=False)
        _ck(lib().rt_ppm_write(os.fsencode(path), rgb.ctypes.data, w, h), "rt_ppm_write")


def ppm_quantise(rgb):
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    out = np.empty(rgb.shape, dtype=np.uint8)
    _ck(lib().rt_ppm_quantise(rgb.ctypes.data, rgb.size, out.ctypes.data), "rt_ppm_quantise")
    return out


ACCEL_LINEAR, ACCEL_LBVH, ACCEL_AUTO, ACCEL_GRID = 0, 1, 2, 4
ACCEL_NAMES = {ACCEL_LINEAR: "linear", ACCEL_LBVH: "lbvh", ACCEL_AUTO: "auto", ACCEL_GRID: "grid"}
KERNEL_MEGA, KERNEL_WAVEFRONT = 0, 1
PBINS_AUTO, PBINS_OFF, PBINS_ON = 0, 1, 2


def make_opts(seed=1227, split=SPLIT_NONE, rank=0, world=1, tile_rows=1, threads=8, accel=ACCEL_AUTO,
              kernel=KERNEL_MEGA, primary_bins=PBINS_AUTO):
    """rt_opts; the defaults are rt_opts_default's (accel AUTO: the fastest structure, same image bit for bit)."""
    o = Opts()
    lib().rt_opts_default(C.byref(o))
    o.seed, o.split, o.rank, o.world, o.tile_rows, o.threads = seed, split, rank, world, tile_rows, threads
    o.accel, o.kernel, o.primary_bins = accel, kernel, primary_bins
    return o


def _ptr(buf):
    """Raw address of a numpy array, a torch tensor (host or device) or an int."""
    if isinstance(buf, int):
        return buf
    if isinstance(buf, np.ndarray):
        return buf.ctypes.data
    return buf.data_ptr()


# ------------------------------------------------------------------ device side -------------
class Renderer:
    """One rt_ctx: the replacement for the reference's device setup + kernel launches
    (GF main.cu:81-95, 300-341)."""

    def __init__(self, device=0):
        self._ctx = _P()
        _ck(lib().rt_create(device, C.byref(self._ctx)), "rt_create")
        self.device = device
        self.double = False

    def close(self):
        if self._ctx:
            lib().rt_destroy(self._ctx)
            self._ctx = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def enable_peer_access(self, peer_device):
        _ck(lib().rt_enable_peer_access(self._ctx, peer_device), "rt_enable_peer_access")

    def set_stream(self, cuda_stream):
        _ck(lib().rt_set_stream(self._ctx, _P(cuda_stream)), "rt_set_stream")

    def upload_scene(self, slots):
        slots = np.ascontiguousarray(slots)
        if slots.dtype == SLOT64_DTYPE:
            self.double = True
            _ck(lib().rt_upload_scene64(self._ctx, slots.ctypes.data, len(slots)), "rt_upload_scene64")
        elif slots.dtype == SLOT_DTYPE:
            self.double = False
            _ck(lib().rt_upload_scene(self._ctx, slots.ctypes.data, len(slots)), "rt_upload_scene")
        else:
            raise TypeError("slots must have dtype SLOT_DTYPE or SLOT64_DTYPE")

    def render(self, cam, opts=None, out=None):
        """Gamma-encoded frame (rows of this partition, W, 3).  `out` may be a numpy array, a
        torch tensor (host pinned or device) or None (a new numpy array)."""
        opts = opts or make_opts()
        double = isinstance(cam, Camera64)
        rows = cam.height
        if opts.split == SPLIT_ROWS and not opts.place_rows:
            rows = lib().rt_partition_rows(cam.height, opts.tile_rows, opts.rank, opts.world, None, 0)
        if out is None:
            out = np.empty((rows, cam.width, 3), dtype=np.float64 if double else np.float32)
        ms = C.c_float(0)
        fn = lib().rt_render64 if double else lib().rt_render
        _ck(fn(self._ctx, C.byref(cam), C.byref(opts), _ptr(out), C.byref(ms)), "rt_render")
        self.last_ms = ms.value
        return out

    def render_partials(self, cam, opts, acc_dev):
        """rt_render_partials: overwrite the DEVICE int64 buffer `acc_dev` (height*width*3) with this rank's fixed-point
        radiance sums (opts.split SPLIT_SPP: its share of the samples; SPLIT_NONE: all of them)."""
        ms = C.c_float(0)
        _ck(lib().rt_render_partials(self._ctx, C.byref(cam), C.byref(opts), _ptr(acc_dev), C.byref(ms)),
            "rt_render_partials")
        self.last_ms = ms.value
        return ms.value

    def finalize(self, cam, acc_dev, out=None):
        """rt_finalize / rt_finalize_sum: one accumulation buffer or a list of them (device pointers, possibly on peer
        GPUs) -> gamma-encoded frame."""
        if out is None:
            out = np.empty((cam.height, cam.width, 3), dtype=np.float32)
        ms = C.c_float(0)
        if isinstance(acc_dev, (list, tuple)):
            ptrs = (C.c_void_p * len(acc_dev))(*[_ptr(a) for a in acc_dev])
            _ck(lib().rt_finalize_sum(self._ctx, C.byref(cam), ptrs, len(acc_dev), _ptr(out), C.byref(ms)), "rt_finalize_sum")
        else:
            _ck(lib().rt_finalize(self._ctx, C.byref(cam), _ptr(acc_dev), _ptr(out), C.byref(ms)), "rt_finalize")
        self.last_finalize_ms = ms.value
        return out

    def primary_hits(self, cam, accel=ACCEL_LINEAR):
        double = isinstance(cam, Camera64)
        ids = np.empty((cam.height, cam.width), dtype=np.int32)
        t = np.empty((cam.height, cam.width), dtype=np.float64 if double else np.float32)
        if accel != ACCEL_LINEAR:
            fn = lib().rt_primary_hits_accel64 if double else lib().rt_primary_hits_accel
            _ck(fn(self._ctx, C.byref(cam), accel, ids.ctypes.data, t.ctypes.data), "rt_primary_hits_accel")
            return ids, t
        fn = lib().rt_primary_hits64 if double else lib().rt_primary_hits
        _ck(fn(self._ctx, C.byref(cam), ids.ctypes.data, t.ctypes.data), "rt_primary_hits")
        return ids, t

    def filter_audit(self, cam, n_rays, seed=1):
        """rt_filter_audit: dict(pairs, exact_pass, filter_pass, missed, skipped)."""
        out = (C.c_uint64 * 5)()
        _ck(lib().rt_filter_audit(self._ctx, C.byref(cam), seed, n_rays, out), "rt_filter_audit")
        return dict(zip(("pairs", "exact_pass", "filter_pass", "missed", "skipped"), [int(v) for v in out]))

    def debug_checks(self, selftest=False):
        """rt_debug_checks: (enabled, first_code, failures) of the checked build's bounds assertions; resets them."""
        en, code, n = C.c_int32(0), C.c_uint32(0), C.c_uint32(0)
        _ck(lib().rt_debug_checks(self._ctx, C.byref(en), C.byref(code), C.byref(n), 1 if selftest else 0), "rt_debug_checks")
        return bool(en.value), code.value, n.value

    def stats(self):
        s = Stats()
        _ck(lib().rt_get_stats(self._ctx, C.byref(s)), "rt_get_stats")
        return s
