"""plade_b200 — B200-native PLADE registration hot path.

Thin ctypes mirror of the C ABI in ``include/plade_b200.h`` (which in turn mirrors the reference's
``registration()`` overloads, /root/reference/code/PLADE/plade.h:44-96).  There is no Python or CPU
implementation behind these calls: if ``libplade_b200.so`` is missing, or no CUDA device is present,
they raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplade_b200.so")

_c_float_p = ctypes.POINTER(ctypes.c_float)
_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_uint_p = ctypes.POINTER(ctypes.c_uint)
_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_ubyte_p = ctypes.POINTER(ctypes.c_ubyte)
ALLREDUCE_FN = ctypes.CFUNCTYPE(None, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_void_p)

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "plade_ctx_create": (ctypes.c_void_p, [ctypes.c_int]),
    "plade_ctx_destroy": (None, [ctypes.c_void_p]),
    "plade_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "plade_create_error": (ctypes.c_char_p, []),
    "plade_device_count": (ctypes.c_int, []),
    "plade_set_param": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_double]),
    "plade_set_shard": (None, [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ALLREDUCE_FN, ctypes.c_void_p]),
    "plade_nccl_unique_id": (ctypes.c_int, [ctypes.c_char_p]),
    "plade_shard_init_nccl": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_int]),
    "plade_shard_init_nccl_all": (ctypes.c_int, [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int]),
    "plade_shard_finalize": (None, [ctypes.c_void_p]),
    "plade_verify_sharded": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, _c_float_p, _c_float_p, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                            _c_int_p, _c_uint_p, _c_float_p]),
    "plade_ply_read": (ctypes.c_longlong, [ctypes.c_char_p, _c_float_p, ctypes.c_size_t]),
    "plade_last_report": (ctypes.c_char_p, [ctypes.c_void_p]),
    "plade_dump_planes_vg": (ctypes.c_int, [_c_float_p, ctypes.c_size_t, _c_int_p, _c_int_p, _c_float_p, ctypes.c_int, ctypes.c_char_p]),
    "plade_launch_count": (ctypes.c_longlong, [ctypes.c_void_p]),
    "plade_stage_times": (ctypes.c_int, [ctypes.c_void_p, _c_double_p, ctypes.c_int]),
    "plade_kernel_times": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, _c_double_p]),
    "plade_timer_start": (ctypes.c_int, [ctypes.c_void_p]),
    "plade_timer_stop_ms": (ctypes.c_float, [ctypes.c_void_p]),
    "plade_register_files": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, _c_float_p]),
    "plade_register_clouds": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t, _c_float_p, ctypes.c_size_t, _c_float_p]),
    "plade_register_with_planes": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t, _c_float_p, ctypes.c_size_t,
                                                  _c_int_p, _c_int_p, _c_float_p, ctypes.c_int,
                                                  _c_int_p, _c_int_p, _c_float_p, ctypes.c_int, _c_float_p]),
    "plade_register_min_support": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t, _c_float_p, ctypes.c_size_t,
                                                  ctypes.c_int, ctypes.c_int, _c_float_p]),
    "plade_register_batch": (ctypes.c_int, [_c_int_p, ctypes.c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_char_p),
                                            ctypes.c_int, _c_float_p, _c_int_p]),
    "plade_batch_release": (None, []),
    "plade_cloud_upload": (ctypes.c_void_p, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t]),
    "plade_cloud_free": (None, [ctypes.c_void_p, ctypes.c_void_p]),
    "plade_cloud_size": (ctypes.c_size_t, [ctypes.c_void_p]),
    "plade_register_resident": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, _c_float_p]),
    "plade_register_resident_with_planes": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                           _c_int_p, _c_int_p, _c_float_p, ctypes.c_int,
                                                           _c_int_p, _c_int_p, _c_float_p, ctypes.c_int, _c_float_p]),
    "plade_extract_planes": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t, ctypes.c_int]),
    "plade_detect_planes": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t, ctypes.c_int]),
    "plade_planes_size": (ctypes.c_int, [ctypes.c_void_p, _c_int_p, ctypes.POINTER(ctypes.c_longlong)]),
    "plade_planes_get": (ctypes.c_int, [ctypes.c_void_p, _c_int_p, _c_int_p, _c_float_p]),
    "plade_score_planes": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t, _c_int_p, _c_float_p, ctypes.c_int,
                                          ctypes.c_float, ctypes.c_float, _c_uint_p, _c_ubyte_p]),
    "plade_largest_component": (ctypes.c_int, [ctypes.c_void_p, _c_ubyte_p, ctypes.c_int, ctypes.c_int, _c_ubyte_p]),
    "plade_refine_candidate": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t, _c_int_p, _c_float_p, _c_float_p, ctypes.c_int,
                                              _c_float_p, _c_float_p, _c_ubyte_p, ctypes.POINTER(ctypes.c_longlong), _c_int_p, _c_double_p]),
    "plade_plane_parameters": (None, [_c_float_p, _c_float_p, _c_float_p, ctypes.c_size_t, _c_float_p, _c_float_p]),
    "plade_plane_ls_fit": (ctypes.c_int, [_c_float_p, ctypes.c_size_t, _c_float_p, _c_float_p]),
    "plade_bitmap_layout": (None, [ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, _c_float_p, ctypes.c_size_t,
                                   ctypes.POINTER(ctypes.c_longlong), _c_int_p]),
    "plade_average_spacing": (ctypes.c_float, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t]),
    "plade_voxel_downsample": (ctypes.c_longlong, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, _c_float_p]),
    "plade_bounding_box": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t, _c_float_p, _c_double_p, _c_float_p]),
    "plade_nearest_points_two_lines": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_int, _c_float_p, _c_double_p]),
    "plade_penetration_filter": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_int, _c_float_p, _c_float_p, _c_float_p, _c_int_p,
                                                _c_float_p, ctypes.c_int, _c_float_p, _c_float_p, _c_float_p, _c_int_p,
                                                _c_float_p, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_void_p]),
    "plade_match_descriptors": (ctypes.c_longlong, [ctypes.c_void_p, _c_float_p, ctypes.c_int, _c_float_p, ctypes.c_int, ctypes.c_float, _c_int_p]),
    "plade_match_results": (ctypes.c_int, [ctypes.c_void_p, _c_int_p, _c_double_p]),
    "plade_transforms_from_matches": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_int, _c_float_p, _c_float_p]),
    "plade_cluster_transforms": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, _c_float_p, ctypes.c_int, ctypes.c_float, ctypes.c_float, _c_int_p]),
    "plade_verify_hypotheses": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t, _c_float_p, ctypes.c_size_t,
                                               _c_float_p, _c_float_p, _c_float_p, ctypes.c_int, ctypes.c_float, ctypes.c_float, _c_uint_p]),
    "plade_verify_upload": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, ctypes.c_size_t, _c_float_p, ctypes.c_size_t, ctypes.c_float]),
    "plade_verify_resident": (ctypes.c_int, [ctypes.c_void_p, _c_float_p, _c_float_p, _c_float_p, ctypes.c_int, ctypes.c_float,
                                             ctypes.c_float, _c_uint_p, _c_float_p]),
    "plade_set_debug": (None, [ctypes.c_void_p, ctypes.c_int]),
    "plade_debug_blob": (ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)]),
}

_lib = None


def load_library():
    """dlopen the product library (no GPU needed for this) and bind every exported symbol."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("plade_b200: %s is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(t)


STAGE_NAMES = ["upload", "planes", "spacing", "downsample", "lines", "descriptors", "match", "hypotheses",
               "penetration", "verify", "total", "verify_kernel_ms", "verify_h", "verify_ns", "verify_nt"]


def dump_planes_vg(xyzn, planes, path):
    """the reference's save_vg: ASCII vertex-group file of a cloud and its planes"""
    lib = load_library()
    a = _f32(xyzn).reshape(-1, 6)
    if not lib.plade_dump_planes_vg(_p(a, _c_float_p), len(a), _p(planes.offsets, _c_int_p), _p(planes.indices, _c_int_p), _p(planes.params, _c_float_p),
                                    len(planes), os.fsencode(path)):
        raise RuntimeError("plade_dump_planes_vg failed: %s" % path)


def ply_read(path):
    """the library's PLY reader: (n, 6) float32, or None when the reference would report a load failure"""
    lib = load_library()
    n = lib.plade_ply_read(os.fsencode(path), None, 0)
    if n < 0:
        return None
    out = np.zeros((n, 6), np.float32)
    if lib.plade_ply_read(os.fsencode(path), _p(out, _c_float_p), n) != n:
        return None
    return out


def device_count():
    """visible CUDA devices (0 when there is none)"""
    return int(load_library().plade_device_count())


def nccl_unique_id():
    """ncclGetUniqueId through the library (rank 0 calls it and ships the 128 bytes to the other ranks)"""
    lib = load_library()
    buf = ctypes.create_string_buffer(128)
    if not lib.plade_nccl_unique_id(buf):
        raise RuntimeError("plade_nccl_unique_id: %s" % lib.plade_create_error().decode())
    return buf.raw


def shard_init_nccl_all(contexts):
    """one process, one context per GPU: ncclCommInitAll over the contexts' devices (rank i = contexts[i])"""
    lib = load_library()
    arr = (ctypes.c_void_p * len(contexts))(*[c.h for c in contexts])
    if not lib.plade_shard_init_nccl_all(arr, len(contexts)):
        raise RuntimeError("plade_shard_init_nccl_all: %s" % lib.plade_create_error().decode())


def plane_parameters(normal, position, xyz):
    """host-side restatement shared with the kernels (planefit.h): (uv[n,2], u[3], v[3]); needs no device"""
    lib = load_library()
    nrm, pos, a = _f32(normal), _f32(position), _f32(xyz).reshape(-1, 3)
    uv, fr = np.zeros((len(a), 2), np.float32), np.zeros(6, np.float32)
    lib.plade_plane_parameters(_p(nrm, _c_float_p), _p(pos, _c_float_p), _p(a, _c_float_p), len(a), _p(uv, _c_float_p), _p(fr, _c_float_p))
    return uv, fr[:3].copy(), fr[3:].copy()


def plane_ls_fit(xyz):
    """host-side restatement of Plane::LeastSquaresFit as the kernels evaluate it: (ok, normal, position)"""
    lib = load_library()
    a = _f32(xyz).reshape(-1, 3)
    n, p = np.zeros(3, np.float32), np.zeros(3, np.float32)
    ok = lib.plade_plane_ls_fit(_p(a, _c_float_p), len(a), _p(n, _c_float_p), _p(p, _c_float_p))
    return bool(ok), n, p


def bitmap_layout(uv, bitmap_eps):
    """BitmapExtent / InBitmap as the kernels evaluate them: ((uextent, vextent), pixel index per point)"""
    lib = load_library()
    a = _f32(uv).reshape(-1, 2)
    ext = (ctypes.c_longlong * 2)()
    pix = np.zeros(len(a), np.int32)
    lib.plade_bitmap_layout(float(a[:, 0].min()), float(a[:, 0].max()), float(a[:, 1].min()), float(a[:, 1].max()), float(bitmap_eps),
                            _p(a, _c_float_p), len(a), ext, _p(pix, _c_int_p))
    return (int(ext[0]), int(ext[1])), pix


class Planes:
    """CSR plane set: offsets[np+1], indices, params[np,4] = (nx, ny, nz, d)."""

    def __init__(self, offsets, indices, params):
        self.offsets = _i32(offsets)
        self.indices = _i32(indices)
        self.params = _f32(params).reshape(-1, 4)

    def __len__(self):
        return len(self.offsets) - 1

    def sizes(self):
        return np.diff(self.offsets)


class Context:
    """One registration context = one CUDA device + stream + scratch (single-threaded)."""

    def __init__(self, device=-1):
        self.lib = load_library()
        self.h = self.lib.plade_ctx_create(device)
        if not self.h:
            raise RuntimeError("plade_b200: cannot create a context: %s" % self.lib.plade_create_error().decode())
        self._cb = None

    def close(self):
        if self.h:
            self.lib.plade_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- configuration
    def set_param(self, name, value):
        if not self.lib.plade_set_param(self.h, name.encode(), float(value)):
            raise KeyError(name)

    # ---- NCCL-sharded verification: the collective lives in the C++ library
    def shard_init_nccl(self, unique_id, rank, world):
        """join the NCCL communicator identified by `unique_id` (128 bytes from nccl_unique_id() on rank 0, sent out of band)"""
        if not self.lib.plade_shard_init_nccl(self.h, bytes(unique_id), int(rank), int(world)):
            raise RuntimeError(self.last_error())

    def shard_finalize(self):
        self.lib.plade_shard_finalize(self.h)

    def verify_sharded(self, R, T, centers, ball_radius, inlier_dist):
        """config 4: all H hypotheses given on every rank, this rank verifies h % world == rank, one ncclAllReduce(u64, max)
        agrees on the winner: (best index, best inlier count, device ms incl. the collective)"""
        R9, T3, C3 = _f32(R).reshape(-1, 9), _f32(T).reshape(-1, 3), _f32(centers).reshape(-1, 3)
        bi, bc, ms = ctypes.c_int(-1), ctypes.c_uint(0), ctypes.c_float(0)
        if not self.lib.plade_verify_sharded(self.h, _p(R9, _c_float_p), _p(T3, _c_float_p), _p(C3, _c_float_p), len(R9), float(ball_radius),
                                             float(inlier_dist), ctypes.byref(bi), ctypes.byref(bc), ctypes.byref(ms)):
            raise RuntimeError(self.last_error())
        return bi.value, bc.value, ms.value

    def last_report(self):
        """dict describing the last registration (planes, hypotheses, winner, inliers, overlap ratio, score)"""
        import json
        t = self.lib.plade_last_report(self.h).decode()
        return json.loads(t) if t else {}

    def set_debug(self, on=True):
        self.lib.plade_set_debug(self.h, 1 if on else 0)

    def set_shard(self, rank, world, allreduce_max=None):
        """allreduce_max: python callable taking and returning an int (u64)."""
        if allreduce_max is None:
            self._cb = ctypes.cast(None, ALLREDUCE_FN)
        else:
            def _tramp(ptr, _user):
                ptr[0] = ctypes.c_ulonglong(int(allreduce_max(int(ptr[0])))).value
            self._cb = ALLREDUCE_FN(_tramp)
        self.lib.plade_set_shard(self.h, rank, world, self._cb, None)

    def last_error(self):
        return self.lib.plade_last_error(self.h).decode()

    def launch_count(self):
        return int(self.lib.plade_launch_count(self.h))

    def stage_times(self):
        out = np.zeros(15, dtype=np.float64)
        self.lib.plade_stage_times(self.h, _p(out, _c_double_p), 15)
        return dict(zip(STAGE_NAMES, out.tolist()))

    def kernel_times(self, kernel):
        """{ms, launches, algorithmic_bytes} of one kernel family during the last registration (CUDA events)."""
        out = np.zeros(3, dtype=np.float64)
        if not self.lib.plade_kernel_times(self.h, kernel.encode(), _p(out, _c_double_p)):
            raise KeyError(kernel)
        return {"ms": float(out[0]), "launches": int(out[1]), "algorithmic_bytes": float(out[2])}

    def timer_start(self):
        self.lib.plade_timer_start(self.h)

    def timer_stop_ms(self):
        return float(self.lib.plade_timer_stop_ms(self.h))

    def blob(self, name, dtype):
        n = ctypes.c_size_t(0)
        p = self.lib.plade_debug_blob(self.h, name.encode(), ctypes.byref(n))
        if not p or n.value == 0:
            return np.zeros(0, dtype=dtype)
        buf = (ctypes.c_char * n.value).from_address(p)
        return np.frombuffer(bytes(buf), dtype=dtype).copy()

    # -- registration (target first, as in the reference)
    def register_files(self, target_ply, source_ply):
        out = np.zeros(16, dtype=np.float32)
        ok = self.lib.plade_register_files(self.h, os.fsencode(target_ply), os.fsencode(source_ply), _p(out, _c_float_p))
        return bool(ok), out.reshape(4, 4)

    def register_clouds(self, tgt_xyzn, src_xyzn):
        t, s = _f32(tgt_xyzn).reshape(-1, 6), _f32(src_xyzn).reshape(-1, 6)
        out = np.zeros(16, dtype=np.float32)
        ok = self.lib.plade_register_clouds(self.h, _p(t, _c_float_p), len(t), _p(s, _c_float_p), len(s), _p(out, _c_float_p))
        return bool(ok), out.reshape(4, 4)

    def register_min_support(self, tgt_xyzn, src_xyzn, ms_t, ms_s):
        t, s = _f32(tgt_xyzn).reshape(-1, 6), _f32(src_xyzn).reshape(-1, 6)
        out = np.zeros(16, dtype=np.float32)
        ok = self.lib.plade_register_min_support(self.h, _p(t, _c_float_p), len(t), _p(s, _c_float_p), len(s), int(ms_t), int(ms_s),
                                                 _p(out, _c_float_p))
        return bool(ok), out.reshape(4, 4)

    def register_with_planes(self, tgt_xyzn, src_xyzn, tp, sp):
        t, s = _f32(tgt_xyzn).reshape(-1, 6), _f32(src_xyzn).reshape(-1, 6)
        out = np.zeros(16, dtype=np.float32)
        ok = self.lib.plade_register_with_planes(
            self.h, _p(t, _c_float_p), len(t), _p(s, _c_float_p), len(s),
            _p(tp.offsets, _c_int_p), _p(tp.indices, _c_int_p), _p(tp.params, _c_float_p), len(tp),
            _p(sp.offsets, _c_int_p), _p(sp.indices, _c_int_p), _p(sp.params, _c_float_p), len(sp), _p(out, _c_float_p))
        return bool(ok), out.reshape(4, 4)

    def upload(self, xyzn):
        a = _f32(xyzn).reshape(-1, 6)
        h = self.lib.plade_cloud_upload(self.h, _p(a, _c_float_p), len(a))
        if not h:
            raise RuntimeError(self.last_error())
        return h

    def free_cloud(self, h):
        self.lib.plade_cloud_free(self.h, h)

    def register_resident(self, tgt_handle, src_handle):
        out = np.zeros(16, dtype=np.float32)
        ok = self.lib.plade_register_resident(self.h, tgt_handle, src_handle, _p(out, _c_float_p))
        return bool(ok), out.reshape(4, 4)

    def register_resident_with_planes(self, tgt_handle, src_handle, tp, sp):
        out = np.zeros(16, dtype=np.float32)
        ok = self.lib.plade_register_resident_with_planes(
            self.h, tgt_handle, src_handle,
            _p(tp.offsets, _c_int_p), _p(tp.indices, _c_int_p), _p(tp.params, _c_float_p), len(tp),
            _p(sp.offsets, _c_int_p), _p(sp.indices, _c_int_p), _p(sp.params, _c_float_p), len(sp), _p(out, _c_float_p))
        return bool(ok), out.reshape(4, 4)

    # -- stages
    def _planes(self, n):
        if n < 0:
            raise RuntimeError(self.last_error())
        np_, ni = ctypes.c_int(0), ctypes.c_longlong(0)
        self.lib.plade_planes_size(self.h, ctypes.byref(np_), ctypes.byref(ni))
        off = np.zeros(np_.value + 1, dtype=np.int32)
        idx = np.zeros(max(ni.value, 1), dtype=np.int32)
        par = np.zeros((max(np_.value, 1), 4), dtype=np.float32)
        self.lib.plade_planes_get(self.h, _p(off, _c_int_p), _p(idx, _c_int_p), _p(par, _c_float_p))
        return Planes(off, idx[:ni.value], par[:np_.value])

    def extract_planes(self, xyzn, init_min_support=10000):
        a = _f32(xyzn).reshape(-1, 6)
        return self._planes(self.lib.plade_extract_planes(self.h, _p(a, _c_float_p), len(a), int(init_min_support)))

    def detect_planes(self, xyzn, min_support):
        a = _f32(xyzn).reshape(-1, 6)
        return self._planes(self.lib.plade_detect_planes(self.h, _p(a, _c_float_p), len(a), int(min_support)))

    def largest_component(self, bitmap):
        """closing + largest 8-connected component of a (ve, ue) uint8 bitmap, as the RANSAC acceptance test runs it"""
        b = np.ascontiguousarray(bitmap, dtype=np.uint8)
        mask = np.zeros_like(b)
        if not self.lib.plade_largest_component(self.h, _p(b, _c_ubyte_p), b.shape[1], b.shape[0], _p(mask, _c_ubyte_p)):
            raise RuntimeError(self.last_error())
        return mask

    def refine_candidate(self, xyzn, normal, position, min_support, assigned=None):
        """acceptance chain of one candidate plane on the device: (size, normal, position, member mask, evaluations, weighted score)"""
        a = _f32(xyzn).reshape(-1, 6)
        nrm, pos = _f32(normal), _f32(position)
        asg = _i32(assigned) if assigned is not None else None
        on, op = np.zeros(3, np.float32), np.zeros(3, np.float32)
        mask = np.zeros(len(a), np.uint8)
        size, ev, sc = ctypes.c_longlong(0), ctypes.c_int(0), ctypes.c_double(0)
        if not self.lib.plade_refine_candidate(self.h, _p(a, _c_float_p), len(a), _p(asg, _c_int_p) if asg is not None else None, _p(nrm, _c_float_p),
                                               _p(pos, _c_float_p), int(min_support), _p(on, _c_float_p), _p(op, _c_float_p), _p(mask, _c_ubyte_p),
                                               ctypes.byref(size), ctypes.byref(ev), ctypes.byref(sc)):
            raise RuntimeError(self.last_error())
        return size.value, on, op, mask, ev.value, sc.value

    def score_planes(self, xyzn, planes4, eps, normal_thresh, assigned=None, want_mask=False):
        a = _f32(xyzn).reshape(-1, 6)
        pl = _f32(planes4).reshape(-1, 4)
        counts = np.zeros(len(pl), dtype=np.uint32)
        mask = np.zeros(len(a), dtype=np.uint8) if want_mask else None
        asg = _i32(assigned) if assigned is not None else None
        ok = self.lib.plade_score_planes(self.h, _p(a, _c_float_p), len(a), _p(asg, _c_int_p) if asg is not None else None,
                                         _p(pl, _c_float_p), len(pl), float(eps), float(normal_thresh), _p(counts, _c_uint_p),
                                         _p(mask, _c_ubyte_p) if mask is not None else None)
        if not ok:
            raise RuntimeError(self.last_error())
        return (counts, mask) if want_mask else counts

    def average_spacing(self, xyzn):
        a = _f32(xyzn).reshape(-1, 6)
        return float(self.lib.plade_average_spacing(self.h, _p(a, _c_float_p), len(a)))

    def voxel_downsample(self, pts, leaf):
        a = _f32(pts)
        stride = a.shape[1]
        out = np.zeros((len(a), 3), dtype=np.float32)
        n = self.lib.plade_voxel_downsample(self.h, _p(a, _c_float_p), len(a), stride, float(leaf), _p(out, _c_float_p))
        if n < 0:
            raise RuntimeError("voxel_downsample failed: " + self.last_error())
        return out[:n].copy()

    def bounding_box(self, xyz):
        a = _f32(xyz).reshape(-1, 3)
        c = np.zeros(3, dtype=np.float32)
        whd = np.zeros(3, dtype=np.float64)
        corners = np.zeros((8, 3), dtype=np.float32)
        rc = self.lib.plade_bounding_box(self.h, _p(a, _c_float_p), len(a), _p(c, _c_float_p), _p(whd, _c_double_p), _p(corners, _c_float_p))
        return rc, c, whd, corners

    def nearest_points_two_lines(self, lines12):
        """lines12: (n, 12) = v1 p1 v2 p2 -> (points (n, 2, 3), length (n,)), PLADE/util.cpp:1167-1229."""
        a = _f32(lines12).reshape(-1, 12)
        pts = np.zeros((len(a), 2, 3), dtype=np.float32)
        length = np.zeros(len(a), dtype=np.float64)
        rc = self.lib.plade_nearest_points_two_lines(self.h, _p(a, _c_float_p), len(a), _p(pts, _c_float_p), _p(length, _c_double_p))
        if rc != 0:
            raise RuntimeError(self.last_error())
        return pts, length

    def penetration_filter(self, src, tgt, hyp12, length_threshold, angle_threshold):
        """K4d (PLADE/util.cpp:466-511).  src / tgt: dicts with planes (P,4), corners4 (P,4,3), center (P,3),
        pts (n,3), off (P+1,); hyp12 (H,12) = R row-major, T.  Returns flags (H,) uint8, 1 = dropped."""
        def side(d):
            return (_f32(d["planes"]).reshape(-1, 4), _f32(d["corners4"]).reshape(-1, 12), _f32(d["center"]).reshape(-1, 3),
                    _f32(d["pts"]).reshape(-1, 3), np.ascontiguousarray(d["off"], dtype=np.int32))
        sp, sc, sce, spt, so = side(src)
        tp, tc, tce, tpt, to = side(tgt)
        h = _f32(hyp12).reshape(-1, 12)
        flags = np.zeros(len(h), dtype=np.uint8)
        rc = self.lib.plade_penetration_filter(self.h, _p(sp, _c_float_p), len(sp), _p(sc, _c_float_p), _p(sce, _c_float_p), _p(spt, _c_float_p), _p(so, _c_int_p),
                                               _p(tp, _c_float_p), len(tp), _p(tc, _c_float_p), _p(tce, _c_float_p), _p(tpt, _c_float_p), _p(to, _c_int_p),
                                               _p(h, _c_float_p), len(h), float(length_threshold), float(angle_threshold), flags.ctypes.data_as(ctypes.c_void_p))
        if rc != 0:
            raise RuntimeError(self.last_error())
        return flags

    def match_descriptors(self, db8, q8, radius=0.04):
        db, q = _f32(db8).reshape(-1, 8), _f32(q8).reshape(-1, 8)
        off = np.zeros(len(q) + 1, dtype=np.int32)
        m = self.lib.plade_match_descriptors(self.h, _p(db, _c_float_p), len(db), _p(q, _c_float_p), len(q), float(radius), _p(off, _c_int_p))
        if m < 0:
            raise RuntimeError(self.last_error())
        idx = np.zeros(max(m, 1), dtype=np.int32)
        d2 = np.zeros(max(m, 1), dtype=np.float64)
        self.lib.plade_match_results(self.h, _p(idx, _c_int_p), _p(d2, _c_double_p))
        return off, idx[:m], d2[:m]

    def transforms_from_matches(self, in18):
        a = _f32(in18).reshape(-1, 18)
        R = np.zeros((len(a), 9), dtype=np.float32)
        T = np.zeros((len(a), 3), dtype=np.float32)
        if not self.lib.plade_transforms_from_matches(self.h, _p(a, _c_float_p), len(a), _p(R, _c_float_p), _p(T, _c_float_p)):
            raise RuntimeError(self.last_error())
        return R.reshape(-1, 3, 3), T

    def cluster_transforms(self, R, T, dist_thresh, ang_thresh):
        R9, T3 = _f32(R).reshape(-1, 9), _f32(T).reshape(-1, 3)
        lab = np.zeros(len(R9), dtype=np.int32)
        if not self.lib.plade_cluster_transforms(self.h, _p(R9, _c_float_p), _p(T3, _c_float_p), len(R9), float(dist_thresh),
                                                 float(ang_thresh), _p(lab, _c_int_p)):
            raise RuntimeError(self.last_error())
        return lab

    def verify_hypotheses(self, src_ds, tgt_ds, R, T, centers, ball_radius, inlier_dist):
        s, t = _f32(src_ds).reshape(-1, 3), _f32(tgt_ds).reshape(-1, 3)
        R9, T3, C3 = _f32(R).reshape(-1, 9), _f32(T).reshape(-1, 3), _f32(centers).reshape(-1, 3)
        counts = np.zeros(len(R9), dtype=np.uint32)
        ok = self.lib.plade_verify_hypotheses(self.h, _p(s, _c_float_p), len(s), _p(t, _c_float_p), len(t), _p(R9, _c_float_p),
                                              _p(T3, _c_float_p), _p(C3, _c_float_p), len(R9), float(ball_radius), float(inlier_dist),
                                              _p(counts, _c_uint_p))
        if not ok:
            raise RuntimeError(self.last_error())
        return counts

    def verify_upload(self, src_ds, tgt_ds, inlier_dist):
        s, t = _f32(src_ds).reshape(-1, 3), _f32(tgt_ds).reshape(-1, 3)
        if not self.lib.plade_verify_upload(self.h, _p(s, _c_float_p), len(s), _p(t, _c_float_p), len(t), float(inlier_dist)):
            raise RuntimeError(self.last_error())

    def verify_resident(self, R, T, centers, ball_radius, inlier_dist):
        R9, T3, C3 = _f32(R).reshape(-1, 9), _f32(T).reshape(-1, 3), _f32(centers).reshape(-1, 3)
        counts = np.zeros(len(R9), dtype=np.uint32)
        ms = ctypes.c_float(0)
        ok = self.lib.plade_verify_resident(self.h, _p(R9, _c_float_p), _p(T3, _c_float_p), _p(C3, _c_float_p), len(R9),
                                            float(ball_radius), float(inlier_dist), _p(counts, _c_uint_p), ctypes.byref(ms))
        if not ok:
            raise RuntimeError(self.last_error())
        return counts, float(ms.value)


def register_batch(pairs, devices=None):
    """Batch mode of the reference CLI (PLADE/main.cpp:97-159) over the GPUs of one box: ``pairs`` is a list of
    (target_ply, source_ply); ``devices`` a list of CUDA ordinals (default: device 0).  Returns (ok[n], T[n,4,4])."""
    lib = load_library()
    devices = _i32([0] if devices is None else devices)
    n = len(pairs)
    tf = (ctypes.c_char_p * max(n, 1))(*[os.fsencode(p[0]) for p in pairs])
    sf = (ctypes.c_char_p * max(n, 1))(*[os.fsencode(p[1]) for p in pairs])
    out = np.zeros((n, 16), dtype=np.float32)
    ok = np.zeros(n, dtype=np.int32)
    rc = lib.plade_register_batch(_p(devices, _c_int_p), len(devices), tf, sf, n, _p(out, _c_float_p), _p(ok, _c_int_p))
    if rc < 0:
        raise RuntimeError("plade_b200: no usable CUDA device (%s); there is no CPU fallback" % lib.plade_create_error().decode())
    return ok.astype(bool), out.reshape(n, 4, 4)


def registration(target, source, ctx=None):
    """Python spelling of ``bool registration(T, target, source)``: file names or (n,6) arrays; returns (ok, 4x4)."""
    own = ctx is None
    ctx = ctx or Context()
    try:
        if isinstance(target, (str, bytes, os.PathLike)):
            return ctx.register_files(target, source)
        return ctx.register_clouds(target, source)
    finally:
        if own:
            ctx.close()
