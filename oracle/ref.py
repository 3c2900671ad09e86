"""ORACLE / TEST INFRASTRUCTURE ONLY.

ctypes access to the two CPU checkers:
  * ``Restate``  — oracle/libplade_oracle.so, the plain-C restatement (oracle/restate.c), kind "port";
  * ``Ref``      — oracle/_ref/libplade_ref.so, the reference's own sources compiled by oracle/Makefile
                   (kind "reference").  It travels to the GPU box as a prebuilt file; /root/reference
                   itself is never read at run time.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference arm may import this.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
RESTATE_PATH = os.path.join(_HERE, "libplade_oracle.so")
REF_PATH = os.path.join(_HERE, "_ref", "libplade_ref.so")
# the same sources at -O3 -march=x86-64-v3 (make -C oracle fast): the TIMING build of the CPU baseline, never used for parity
REF_FAST_PATH = os.path.join(_HERE, "_ref_fast", "libplade_ref.so")

_fp = ctypes.POINTER(ctypes.c_float)
_ip = ctypes.POINTER(ctypes.c_int)
_up = ctypes.POINTER(ctypes.c_uint32)
_dp = ctypes.POINTER(ctypes.c_double)
_bp = ctypes.POINTER(ctypes.c_ubyte)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t):
    return a.ctypes.data_as(t)


def build_restate():
    subprocess.check_call(["make", "-s", "-C", _HERE, "restate"])


def have_ref():
    return os.path.exists(REF_PATH)


def have_ref_fast():
    return os.path.exists(REF_FAST_PATH)


class Restate:
    def __init__(self):
        if not os.path.exists(RESTATE_PATH):
            build_restate()
        self.lib = ctypes.CDLL(RESTATE_PATH)
        L = self.lib
        L.oracle_verify_counts.argtypes = [_fp, ctypes.c_size_t, _fp, ctypes.c_size_t, _fp, _fp, _fp, ctypes.c_int, ctypes.c_float, ctypes.c_float, _up]
        L.oracle_voxel_downsample.restype = ctypes.c_longlong
        L.oracle_voxel_downsample.argtypes = [_fp, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, _fp]
        L.oracle_knn_sqdist.argtypes = [_fp, ctypes.c_size_t, ctypes.c_int, _ip, ctypes.c_int, ctypes.c_int, _fp]
        L.oracle_average_spacing.restype = ctypes.c_float
        L.oracle_average_spacing.argtypes = [_fp, ctypes.c_size_t, ctypes.c_int]
        L.oracle_match_descriptors.restype = ctypes.c_longlong
        L.oracle_match_descriptors.argtypes = [_fp, ctypes.c_int, _fp, ctypes.c_int, ctypes.c_float, _ip, _ip, _dp]
        L.oracle_score_planes.argtypes = [_fp, ctypes.c_size_t, _ip, _fp, ctypes.c_int, ctypes.c_float, ctypes.c_float, _up, _bp]
        L.oracle_cluster_transforms.argtypes = [_fp, _fp, ctypes.c_int, ctypes.c_float, ctypes.c_float, _ip]

    def verify_counts(self, src, tgt, R, T, centers, ball_radius, inlier_dist):
        s, t = _f32(src).reshape(-1, 3), _f32(tgt).reshape(-1, 3)
        R9, T3, C3 = _f32(R).reshape(-1, 9), _f32(T).reshape(-1, 3), _f32(centers).reshape(-1, 3)
        out = np.zeros(len(R9), dtype=np.uint32)
        self.lib.oracle_verify_counts(_p(s, _fp), len(s), _p(t, _fp), len(t), _p(R9, _fp), _p(T3, _fp), _p(C3, _fp), len(R9),
                                      float(ball_radius), float(inlier_dist), _p(out, _up))
        return out

    def voxel_downsample(self, pts, leaf):
        a = _f32(pts)
        out = np.zeros((len(a), 3), dtype=np.float32)
        n = self.lib.oracle_voxel_downsample(_p(a, _fp), len(a), a.shape[1], float(leaf), _p(out, _fp))
        if n < 0:
            raise RuntimeError("oracle_voxel_downsample: %d" % n)
        return out[:n].copy()

    def knn_sqdist(self, pts, qidx, k):
        a, q = _f32(pts), _i32(qidx)
        out = np.zeros((len(q), k), dtype=np.float32)
        self.lib.oracle_knn_sqdist(_p(a, _fp), len(a), a.shape[1], _p(q, _ip), len(q), k, _p(out, _fp))
        return out

    def average_spacing(self, pts):
        a = _f32(pts)
        return float(self.lib.oracle_average_spacing(_p(a, _fp), len(a), a.shape[1]))

    def match_descriptors(self, db8, q8, radius=0.04):
        db, q = _f32(db8).reshape(-1, 8), _f32(q8).reshape(-1, 8)
        off = np.zeros(len(q) + 1, dtype=np.int32)
        m = self.lib.oracle_match_descriptors(_p(db, _fp), len(db), _p(q, _fp), len(q), float(radius), _p(off, _ip), None, None)
        idx = np.zeros(max(m, 1), dtype=np.int32)
        d2 = np.zeros(max(m, 1), dtype=np.float64)
        self.lib.oracle_match_descriptors(_p(db, _fp), len(db), _p(q, _fp), len(q), float(radius), _p(off, _ip), _p(idx, _ip), _p(d2, _dp))
        return off, idx[:m], d2[:m]

    def score_planes(self, xyzn, planes4, eps, normal_thresh, assigned=None, want_mask=False):
        a, pl = _f32(xyzn).reshape(-1, 6), _f32(planes4).reshape(-1, 4)
        counts = np.zeros(len(pl), dtype=np.uint32)
        mask = np.zeros(len(a), dtype=np.uint8) if want_mask else None
        asg = _i32(assigned) if assigned is not None else None
        self.lib.oracle_score_planes(_p(a, _fp), len(a), _p(asg, _ip) if asg is not None else None, _p(pl, _fp), len(pl),
                                     float(eps), float(normal_thresh), _p(counts, _up), _p(mask, _bp) if mask is not None else None)
        return (counts, mask) if want_mask else counts

    def cluster_transforms(self, R, T, dist_thresh, ang_thresh):
        R9, T3 = _f32(R).reshape(-1, 9), _f32(T).reshape(-1, 3)
        lab = np.zeros(len(R9), dtype=np.int32)
        self.lib.oracle_cluster_transforms(_p(R9, _fp), _p(T3, _fp), len(R9), float(dist_thresh), float(ang_thresh), _p(lab, _ip))
        return lab


class Ref:
    """The reference's own code (compiled from /root/reference by oracle/Makefile)."""

    def __init__(self, quiet=True, fast=False):
        if not have_ref():
            raise RuntimeError("oracle/_ref/libplade_ref.so not built (needs /root/reference; `make -C oracle ref`)")
        self.fast = bool(fast and have_ref_fast())
        self.lib = ctypes.CDLL(REF_FAST_PATH if self.fast else REF_PATH)
        self.quiet = quiet
        L = self.lib
        L.ref_set_seed.argtypes = [ctypes.c_long]
        L.ref_blob.restype = ctypes.c_void_p
        L.ref_blob.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_size_t)]
        L.ref_registration_files.argtypes = [ctypes.c_char_p, ctypes.c_char_p, _fp]
        L.ref_registration_clouds.argtypes = [_fp, ctypes.c_size_t, _fp, ctypes.c_size_t, _fp]
        L.ref_registration_planes.argtypes = [_fp, ctypes.c_size_t, _fp, ctypes.c_size_t, _ip, _ip, _fp, ctypes.c_int,
                                              _ip, _ip, _fp, ctypes.c_int, _fp, ctypes.c_int]
        L.ref_extract.argtypes = [_fp, ctypes.c_size_t, ctypes.c_int, ctypes.c_char_p]
        L.ref_detect.argtypes = [_fp, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_char_p]
        L.ref_average_spacing.restype = ctypes.c_float
        L.ref_average_spacing.argtypes = [_fp, ctypes.c_size_t]
        L.ref_voxel_downsample.argtypes = [_fp, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, ctypes.c_char_p]
        L.ref_bounding_box.argtypes = [_fp, ctypes.c_size_t, _fp, _dp, _fp]
        L.ref_plane_intersection.argtypes = [_fp, _fp, _fp, _fp]
        L.ref_nearest_points_two_lines.argtypes = [_fp, _fp, _fp, _fp, _fp, _fp, _dp]
        L.ref_line_line_intersection.argtypes = [_fp, _fp, _fp, _fp, _fp]
        L.ref_match_descriptors.argtypes = [_fp, ctypes.c_int, _fp, ctypes.c_int, ctypes.c_double, ctypes.c_char_p]
        L.ref_pair_descriptor.argtypes = [_fp] * 9
        L.ref_transform_from_two_vecs.argtypes = [_fp, ctypes.c_int, _fp, _fp]
        L.ref_cluster_transformations.argtypes = [_fp, _fp, ctypes.c_int, ctypes.c_float, ctypes.c_float, _ip]
        L.ref_compute_overlap.argtypes = [_fp, ctypes.c_size_t, _fp, ctypes.c_size_t, _fp, _fp, _fp, ctypes.c_int, ctypes.c_float, ctypes.c_float, _fp, _ip]
        # RANSAC building blocks and the PLY reader
        L.ref_plane_parameters.restype = None
        L.ref_plane_parameters.argtypes = [_fp, _fp, _fp, ctypes.c_size_t, _fp, _fp]
        L.ref_connected_component.restype = ctypes.c_longlong
        L.ref_connected_component.argtypes = [_fp, ctypes.c_size_t, _fp, _fp, _ip, ctypes.c_size_t, ctypes.c_float, ctypes.c_int, _ip]
        L.ref_plane_ls_fit.argtypes = [_fp, ctypes.c_size_t, _ip, ctypes.c_size_t, _fp, _fp]
        L.ref_weighted_score.restype = ctypes.c_float
        L.ref_weighted_score.argtypes = [_fp, ctypes.c_size_t, _fp, _fp, _ip, ctypes.c_size_t, ctypes.c_float, ctypes.c_float]
        L.ref_refine_candidate.restype = ctypes.c_longlong
        L.ref_refine_candidate.argtypes = [_fp, ctypes.c_size_t, _ip, _fp, _fp, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_uint,
                                           _fp, _fp, _ip, _ip]
        L.ref_load_ply.restype = ctypes.c_longlong
        L.ref_load_ply.argtypes = [ctypes.c_char_p, ctypes.c_char_p]

    class _Quiet:
        def __init__(self, on):
            self.on = on

        def __enter__(self):
            if self.on:
                import sys
                sys.stdout.flush()
                self.saved = os.dup(1)
                self.null = os.open(os.devnull, os.O_WRONLY)
                os.dup2(self.null, 1)

        def __exit__(self, *a):
            if self.on:
                self.lib_flush()
                os.dup2(self.saved, 1)
                os.close(self.null)
                os.close(self.saved)

        @staticmethod
        def lib_flush():
            try:
                ctypes.CDLL(None).fflush(None)
            except Exception:
                pass

    def q(self):
        return Ref._Quiet(self.quiet)

    def set_seed(self, seed):
        self.lib.ref_set_seed(int(seed))

    def blob(self, name, dtype):
        n = ctypes.c_size_t(0)
        p = self.lib.ref_blob(name.encode(), ctypes.byref(n))
        if not p or n.value == 0:
            return np.zeros(0, dtype=dtype)
        return np.frombuffer(bytes((ctypes.c_char * n.value).from_address(p)), dtype=dtype).copy()

    def registration_files(self, tgt, src):
        out = np.zeros(16, dtype=np.float32)
        with self.q():
            ok = self.lib.ref_registration_files(os.fsencode(tgt), os.fsencode(src), _p(out, _fp))
        return bool(ok), out.reshape(4, 4)

    def registration_clouds(self, tgt, src):
        t, s = _f32(tgt).reshape(-1, 6), _f32(src).reshape(-1, 6)
        out = np.zeros(16, dtype=np.float32)
        with self.q():
            ok = self.lib.ref_registration_clouds(_p(t, _fp), len(t), _p(s, _fp), len(s), _p(out, _fp))
        return bool(ok), out.reshape(4, 4)

    def registration_planes(self, tgt, src, tp, sp, dump=False):
        """tp/sp: (offsets, indices, params[np,4])."""
        t, s = _f32(tgt).reshape(-1, 6), _f32(src).reshape(-1, 6)
        to, ti, tpar = _i32(tp[0]), _i32(tp[1]), _f32(tp[2]).reshape(-1, 4)
        so, si, spar = _i32(sp[0]), _i32(sp[1]), _f32(sp[2]).reshape(-1, 4)
        out = np.zeros(16, dtype=np.float32)
        with self.q():
            ok = self.lib.ref_registration_planes(_p(t, _fp), len(t), _p(s, _fp), len(s), _p(to, _ip), _p(ti, _ip), _p(tpar, _fp), len(tpar),
                                                  _p(so, _ip), _p(si, _ip), _p(spar, _fp), len(spar), _p(out, _fp), 1 if dump else 0)
        return bool(ok), out.reshape(4, 4)

    def _planes(self, prefix):
        return (self.blob(prefix + "plane_offsets", np.int32), self.blob(prefix + "plane_indices", np.int32),
                self.blob(prefix + "plane_params", np.float32).reshape(-1, 4))

    def extract(self, xyzn, init_min_support=10000, prefix="x_"):
        a = _f32(xyzn).reshape(-1, 6)
        with self.q():
            self.lib.ref_extract(_p(a, _fp), len(a), int(init_min_support), prefix.encode())
        return self._planes(prefix)

    def detect(self, xyzn, min_support, prefix="d_"):
        a = _f32(xyzn).reshape(-1, 6)
        with self.q():
            self.lib.ref_detect(_p(a, _fp), len(a), int(min_support), 0.005, 0.02, 0.8, 0.001, prefix.encode())
        return self._planes(prefix)

    def average_spacing(self, xyzn):
        a = _f32(xyzn).reshape(-1, 6)
        return float(self.lib.ref_average_spacing(_p(a, _fp), len(a)))

    def voxel_downsample(self, pts, leaf):
        a = _f32(pts)
        n = self.lib.ref_voxel_downsample(_p(a, _fp), len(a), a.shape[1], float(leaf), b"vox")
        if n < 0:
            raise RuntimeError("ref_voxel_downsample failed")
        return self.blob("vox", np.float32).reshape(-1, 3)

    def bounding_box(self, xyz):
        a = _f32(xyz).reshape(-1, 3)
        c = np.zeros(3, dtype=np.float32)
        whd = np.zeros(3, dtype=np.float64)
        corners = np.zeros((8, 3), dtype=np.float32)
        rc = self.lib.ref_bounding_box(_p(a, _fp), len(a), _p(c, _fp), _p(whd, _dp), _p(corners, _fp))
        return rc, c, whd, corners

    def plane_intersection(self, p1, p2):
        a, b = _f32(p1), _f32(p2)
        v, p = np.zeros(3, np.float32), np.zeros(3, np.float32)
        rc = self.lib.ref_plane_intersection(_p(a, _fp), _p(b, _fp), _p(v, _fp), _p(p, _fp))
        return rc, v, p

    def self_adjoint_eig3(self, A):
        a = _f32(A).reshape(9)
        w, V = np.zeros(3, np.float32), np.zeros(9, np.float32)
        self.lib.ref_self_adjoint_eig3(_p(a, _fp), _p(w, _fp), _p(V, _fp))
        return w, V.reshape(3, 3)

    def line_line_intersection(self, v1, p1, v2, p2):
        a, b, c, d = _f32(v1), _f32(p1), _f32(v2), _f32(p2)
        o = np.zeros(3, np.float32)
        rc = self.lib.ref_line_line_intersection(_p(a, _fp), _p(b, _fp), _p(c, _fp), _p(d, _fp), _p(o, _fp))
        return rc, o

    def penetration_filter(self, src, tgt, hyp12, length_threshold, angle_threshold):
        """src / tgt: dicts with planes (P,4), corners4 (P,4,3), center (P,3), pts (n,3), off (P+1,).  Returns flags (H,)."""
        def side(d):
            return (_f32(d["planes"]).reshape(-1, 4), _f32(d["corners4"]).reshape(-1, 12), _f32(d["center"]).reshape(-1, 3),
                    _f32(d["pts"]).reshape(-1, 3), np.ascontiguousarray(d["off"], dtype=np.int32))
        sp, sc, sce, spt, so = side(src)
        tp, tc, tce, tpt, to = side(tgt)
        h = _f32(hyp12).reshape(-1, 12)
        flags = np.zeros(len(h), np.uint8)
        self.lib.ref_penetration_filter.restype = None
        self.lib.ref_penetration_filter.argtypes = None
        self.lib.ref_penetration_filter(_p(sp, _fp), ctypes.c_int(len(sp)), _p(sc, _fp), _p(sce, _fp), _p(spt, _fp), _p(so, _ip),
                                        _p(tp, _fp), ctypes.c_int(len(tp)), _p(tc, _fp), _p(tce, _fp), _p(tpt, _fp), _p(to, _ip),
                                        _p(h, _fp), ctypes.c_int(len(h)), ctypes.c_float(length_threshold), ctypes.c_float(angle_threshold),
                                        flags.ctypes.data_as(ctypes.c_void_p))
        return flags

    def nearest_points_two_lines(self, v1, p1, v2, p2):
        a, b, c, d = _f32(v1), _f32(p1), _f32(v2), _f32(p2)
        q1, q2 = np.zeros(3, np.float32), np.zeros(3, np.float32)
        ln = ctypes.c_double(0)
        rc = self.lib.ref_nearest_points_two_lines(_p(a, _fp), _p(b, _fp), _p(c, _fp), _p(d, _fp), _p(q1, _fp), _p(q2, _fp), ctypes.byref(ln))
        return rc, q1, q2, ln.value

    def match_descriptors(self, db8, q8, radius=0.04):
        db, q = _f32(db8).reshape(-1, 8), _f32(q8).reshape(-1, 8)
        self.lib.ref_match_descriptors(_p(db, _fp), len(db), _p(q, _fp), len(q), float(radius), b"m")
        return self.blob("m_offsets", np.int32), self.blob("m_idx", np.int32), self.blob("m_dist", np.float32)

    def transform_from_two_vecs(self, in18):
        a = _f32(in18).reshape(-1, 18)
        R = np.zeros((len(a), 9), np.float32)
        T = np.zeros((len(a), 3), np.float32)
        self.lib.ref_transform_from_two_vecs(_p(a, _fp), len(a), _p(R, _fp), _p(T, _fp))
        return R.reshape(-1, 3, 3), T

    def cluster_transformations(self, R, T, dist_thresh, ang_thresh):
        R9, T3 = _f32(R).reshape(-1, 9), _f32(T).reshape(-1, 3)
        lab = np.full(len(R9), -1, dtype=np.int32)
        nc = self.lib.ref_cluster_transformations(_p(R9, _fp), _p(T3, _fp), len(R9), float(dist_thresh), float(ang_thresh), _p(lab, _ip))
        return nc, lab

    def compute_overlap(self, src_ds, tgt_ds, R, T, centers, query_radius, inlier_distance):
        s, t = _f32(src_ds).reshape(-1, 3), _f32(tgt_ds).reshape(-1, 3)
        R9, T3, C3 = _f32(R).reshape(-1, 9), _f32(T).reshape(-1, 3), _f32(centers).reshape(-1, 3)
        ov = np.zeros(len(R9), np.float32)
        cnt = np.zeros(len(R9), np.int32)
        self.lib.ref_compute_overlap(_p(s, _fp), len(s), _p(t, _fp), len(t), _p(R9, _fp), _p(T3, _fp), _p(C3, _fp), len(R9),
                                     float(query_radius), float(inlier_distance), _p(ov, _fp), _p(cnt, _ip))
        return ov, cnt


def _ref_methods():
    def plane_parameters(self, normal, pos, xyz):
        """PlanePrimitiveShape::Parameters: (uv[n,2], u[3], v[3])"""
        nrm, p, a = _f32(normal), _f32(pos), _f32(xyz).reshape(-1, 3)
        uv = np.zeros((len(a), 2), np.float32)
        fr = np.zeros(6, np.float32)
        self.lib.ref_plane_parameters(_p(nrm, _fp), _p(p, _fp), _p(a, _fp), len(a), _p(uv, _fp), _p(fr, _fp))
        return uv, fr[:3].copy(), fr[3:].copy()

    def connected_component(self, xyzn, normal, pos, idx, bitmap_eps, do_filtering=True):
        """BitmapPrimitiveShape::ConnectedComponent on points idx: sorted member indices of the largest component"""
        a, nrm, p, ix = _f32(xyzn).reshape(-1, 6), _f32(normal), _f32(pos), _i32(idx)
        out = np.zeros(max(len(ix), 1), np.int32)
        k = self.lib.ref_connected_component(_p(a, _fp), len(a), _p(nrm, _fp), _p(p, _fp), _p(ix, _ip), len(ix), float(bitmap_eps),
                                             1 if do_filtering else 0, _p(out, _ip))
        return out[:k].copy()

    def plane_ls_fit(self, xyzn, idx):
        a, ix = _f32(xyzn).reshape(-1, 6), _i32(idx)
        n, p = np.zeros(3, np.float32), np.zeros(3, np.float32)
        ok = self.lib.ref_plane_ls_fit(_p(a, _fp), len(a), _p(ix, _ip), len(ix), _p(n, _fp), _p(p, _fp))
        return bool(ok), n, p

    def weighted_score(self, xyzn, normal, pos, idx, epsilon, normal_thresh=0.8):
        a, nrm, p, ix = _f32(xyzn).reshape(-1, 6), _f32(normal), _f32(pos), _i32(idx)
        return float(self.lib.ref_weighted_score(_p(a, _fp), len(a), _p(nrm, _fp), _p(p, _fp), _p(ix, _ip), len(ix), float(epsilon), float(normal_thresh)))

    def refine_candidate(self, xyzn, normal, pos, epsilon, normal_thresh, bitmap_eps, min_support, assigned=None):
        """acceptance chain of one candidate (Detect, R/RansacShapeDetector.cpp:613-655): (size, normal, pos, members, (accepted, tried))"""
        a, nrm, p = _f32(xyzn).reshape(-1, 6), _f32(normal), _f32(pos)
        asg = _i32(assigned) if assigned is not None else None
        on, op = np.zeros(3, np.float32), np.zeros(3, np.float32)
        out = np.zeros(len(a), np.int32)
        tr = np.zeros(2, np.int32)
        with self.q():
            k = self.lib.ref_refine_candidate(_p(a, _fp), len(a), _p(asg, _ip) if asg is not None else None, _p(nrm, _fp), _p(p, _fp),
                                              float(epsilon), float(normal_thresh), float(bitmap_eps), int(min_support), _p(on, _fp), _p(op, _fp),
                                              _p(out, _ip), _p(tr, _ip))
        return int(k), on, op, out[:k].copy(), (int(tr[0]), int(tr[1]))

    def load_ply_ref(self, path):
        """load_ply_cloud (PLADE/util.cpp:1505-1546): (n, 6) float32 or None"""
        with self.q():
            k = self.lib.ref_load_ply(os.fsencode(path), b"ply")
        if k < 0:
            return None
        return self.blob("ply", np.float32).reshape(-1, 6)

    for f in (plane_parameters, connected_component, plane_ls_fit, weighted_score, refine_candidate, load_ply_ref):
        setattr(Ref, f.__name__, f)


_ref_methods()


def load_ply(path):
    """binary little-endian float x y z nx ny nz (the sample_data layout) -> (n, 6) float32."""
    with open(path, "rb") as f:
        n = 0
        while True:
            line = f.readline().decode("ascii", "replace").strip()
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
            if line == "end_header":
                break
        return np.fromfile(f, dtype="<f4", count=n * 6).reshape(n, 6)
