"""ctypes wrapper of the CPU oracle (oracle/s2m_oracle.cpp).

TEST INFRASTRUCTURE ONLY: import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs, never from sc-a-loam_b200/.
PARITY UNPINNED (see the header of s2m_oracle.cpp).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libs2m_oracle.so")
_REF = os.path.join(_HERE, "_ref", "libnanoflann_ref.so")


class Stats(ctypes.Structure):
    _fields_ = [("n_corner_in", ctypes.c_int), ("n_surf_in", ctypes.c_int),
                ("n_corner_ds", ctypes.c_int), ("n_surf_ds", ctypes.c_int),
                ("n_map_corner", ctypes.c_int), ("n_map_surf", ctypes.c_int),
                ("n_edge", ctypes.c_int * 2), ("n_plane", ctypes.c_int * 2),
                ("optimized", ctypes.c_int), ("lm_iters", ctypes.c_int * 2),
                ("lm_term", ctypes.c_int * 2),
                ("cost_initial", ctypes.c_double * 2), ("cost_final", ctypes.c_double * 2),
                ("cand_corner", ctypes.c_double), ("cand_surf", ctypes.c_double),
                ("t_ms", ctypes.c_double * 8)]


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in ("s2m_oracle.cpp", "scan_registration.cpp")]
    stale = (not os.path.exists(_LIB)) or os.path.getmtime(_LIB) < max(os.path.getmtime(f) for f in srcs)
    if force or stale or (os.path.exists("/root/reference") and not os.path.exists(_REF)):
        subprocess.check_call(["make", "-C", _HERE, "CXX=g++"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = ctypes.CDLL(_LIB)
        vp, ci, cf, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double
        L.orc_create.restype = vp
        L.orc_create.argtypes = [cf, cf]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_options.argtypes = [vp, ci, ci, ci]
        L.orc_register.argtypes = [vp, vp, ci, vp, ci, vp, vp, vp, vp, vp]
        L.orc_get_correction.argtypes = [vp, vp, vp]
        L.orc_map_upload.argtypes = [vp, vp, ci, vp, ci]
        L.orc_get_local_map.argtypes = [vp, ci, vp, vp, ci]
        L.orc_get_map.argtypes = [vp, ci, vp, ci]
        L.orc_get_window.argtypes = [vp, vp]
        L.orc_get_surround.argtypes = [vp, vp, ci]
        L.orc_debug_knn.argtypes = [vp, ci, vp, vp, ci, ci, vp, vp]
        L.orc_voxel_grid.argtypes = [vp, ci, cf, vp]
        L.orc_knn.argtypes = [vp, ci, vp, ci, ci, vp, vp]
        L.orc_eig3.argtypes = [vp, vp, vp]
        L.orc_plane_qr.argtypes = [vp, vp]
        L.orc_edge_factor.argtypes = [vp, vp, vp, vp, vp, vp]
        L.orc_plane_factor.argtypes = [vp, vp, cd, vp, vp, vp]
        L.orc_solve.argtypes = [vp, vp, ci, vp, ci, vp, vp, vp]
        L.orc_trace_sizes.argtypes = [vp, vp]
        L.orc_trace_cloud.argtypes = [vp, ci, vp]
        L.orc_trace_knn.argtypes = [vp, ci, ci, vp, vp, vp]
        L.orc_trace_lm.argtypes = [vp, ci, vp, vp, vp, vp, vp]
        _lib = L
    return _lib


def ref_lib():
    """The reference tree's own nanoflann KD-tree (oracle/_ref), or None if not built."""
    if not os.path.exists(_REF):
        return None
    R = ctypes.CDLL(_REF)
    R.ref_nanoflann_knn5.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                     ctypes.c_void_p, ctypes.c_void_p]
    return R


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


class Oracle:
    """One laserMapping instance (rows A..W of SURVEY.md section 8a) on the CPU."""

    def __init__(self, line_res=0.4, plane_res=0.8, use_kdtree=True, skip_optimization=False, trace=False):
        self.L = lib()
        self.h = self.L.orc_create(line_res, plane_res)
        self.L.orc_set_options(self.h, int(use_kdtree), int(skip_optimization), int(trace))
        self.stats = Stats()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    def register(self, corner, surf, q_wodom, t_wodom):
        corner, surf = _f32(corner).reshape(-1, 4), _f32(surf).reshape(-1, 4)
        q, t = _f64(q_wodom), _f64(t_wodom)
        qo, to = np.zeros(4), np.zeros(3)
        rc = self.L.orc_register(self.h, corner.ctypes.data, len(corner), surf.ctypes.data, len(surf),
                                 q.ctypes.data, t.ctypes.data, qo.ctypes.data, to.ctypes.data,
                                 ctypes.byref(self.stats))
        return rc, qo, to

    def map_upload(self, corner, surf):
        corner, surf = _f32(corner).reshape(-1, 4), _f32(surf).reshape(-1, 4)
        return self.L.orc_map_upload(self.h, corner.ctypes.data, len(corner), surf.ctypes.data, len(surf))

    def local_map(self, cls, centre_t):
        c = _f64(centre_t)
        n = self.L.orc_get_local_map(self.h, cls, c.ctypes.data, None, 0)
        out = np.zeros((max(n, 1), 4), np.float32)
        self.L.orc_get_local_map(self.h, cls, c.ctypes.data, out.ctypes.data, n)
        return out[:n]

    def get_map(self, cls):
        n = self.L.orc_get_map(self.h, cls, None, 0)
        out = np.zeros((max(n, 1), 4), np.float32)
        self.L.orc_get_map(self.h, cls, out.ctypes.data, n)
        return out[:n]

    def surround(self):
        n = self.L.orc_get_surround(self.h, None, 0)
        out = np.zeros((max(n, 1), 4), np.float32)
        self.L.orc_get_surround(self.h, out.ctypes.data, n)
        return out[:n]

    def window(self):
        c = np.zeros(3, np.int32)
        self.L.orc_get_window(self.h, c.ctypes.data)
        return c

    def debug_knn(self, cls, centre_t, q_xyz, method=0):
        q = _f32(q_xyz).reshape(-1, 3)
        c = _f64(centre_t)
        idx = np.zeros((len(q), 5), np.int32)
        d2 = np.zeros((len(q), 5), np.float32)
        self.L.orc_debug_knn(self.h, cls, c.ctypes.data, q.ctypes.data, len(q), method,
                             idx.ctypes.data, d2.ctypes.data)
        return idx, d2

    def trace_cloud(self, which):
        sizes = np.zeros(4, np.int32)
        self.L.orc_trace_sizes(self.h, sizes.ctypes.data)
        out = np.zeros((max(int(sizes[which]), 1), 4), np.float32)
        n = self.L.orc_trace_cloud(self.h, which, out.ctypes.data)
        return out[:n]

    def trace_knn(self, outer, cls):
        sizes = np.zeros(4, np.int32)
        self.L.orc_trace_sizes(self.h, sizes.ctypes.data)
        n = int(sizes[cls])
        idx = np.zeros((max(n, 1), 5), np.int32)
        d2 = np.zeros((max(n, 1), 5), np.float32)
        used = np.zeros(max(n, 1), np.uint8)
        m = self.L.orc_trace_knn(self.h, outer, cls, idx.ctypes.data, d2.ctypes.data, used.ctypes.data)
        return idx[:m], d2[:m], used[:m]

    def trace_lm(self, outer):
        pose, sums, iters = np.zeros(7), np.zeros(28), np.zeros((4, 6))
        n, term = ctypes.c_int(), ctypes.c_int()
        self.L.orc_trace_lm(self.h, outer, pose.ctypes.data, sums.ctypes.data, iters.ctypes.data,
                            ctypes.byref(n), ctypes.byref(term))
        return pose, sums, iters, n.value, term.value


def voxel_grid(pts, leaf):
    pts = _f32(pts).reshape(-1, 4)
    out = np.zeros((max(len(pts), 1), 4), np.float32)
    n = lib().orc_voxel_grid(pts.ctypes.data, len(pts), leaf, out.ctypes.data)
    return out[:n]


class Odometer:
    """laserOdometry.cpp:220-591 restated (DISTORTION 0): feed the four feature clouds of each sweep."""

    def __init__(self):
        L = lib()
        L.orc_odom_create.restype = ctypes.c_void_p
        L.orc_odom_destroy.argtypes = [ctypes.c_void_p]
        self.L, self.h = L, ctypes.c_void_p(L.orc_odom_create())
        self.counts = np.zeros(4, np.int32)
        self.para = np.zeros(7)
        self._n = (0, 0)

    def __del__(self):
        try:
            self.L.orc_odom_destroy(self.h)
        except Exception:
            pass

    def step(self, sharp, flat, less_sharp, less_flat):
        """-> (q_w_curr[4] x,y,z,w, t_w_curr[3]); self.para = (q_last_curr, t_last_curr), self.counts =
        [corner pass 0, corner pass 1, plane pass 0, plane pass 1]"""
        a = [_f32(x).reshape(-1, 4) for x in (sharp, flat, less_sharp, less_flat)]
        q, t = np.zeros(4), np.zeros(3)
        args = []
        for x in a:
            args += [ctypes.c_void_p(x.ctypes.data), len(x)]
        self.L.orc_odom_step(self.h, *args, ctypes.c_void_p(q.ctypes.data), ctypes.c_void_p(t.ctypes.data),
                             ctypes.c_void_p(self.para.ctypes.data), ctypes.c_void_p(self.counts.ctypes.data))
        self._n = (len(a[0]), len(a[1]))
        return q, t

    def trace(self, opt):
        """correspondence indices of pass `opt` of the last step: edge (n_sharp,2), plane (n_flat,3); -1 = none"""
        e = np.full((self._n[0], 2), -1, np.int32)
        p = np.full((self._n[1], 3), -1, np.int32)
        if self.counts.sum() or True:
            self.L.orc_odom_trace(self.h, opt, ctypes.c_void_p(e.ctypes.data), ctypes.c_void_p(p.ctypes.data))
        return e, p


SENSORS = {"HDL64": 0, "VLP16": 1, "OS1-64": 2, "HDL32": 3}
FEATURE_CLOUDS = ("full", "sharp", "less_sharp", "flat", "less_flat")


def scan_registration(sensor, xyz, minimum_range, tie_rule=0):
    """scanRegistration.cpp:116-454 restated: raw sweep (n,3) -> dict of the five clouds it publishes
    (full = /velodyne_cloud_2, sharp, less_sharp, flat, less_flat), each (m,4) xyzi."""
    xyz = _f32(xyz).reshape(-1, 3)
    cap = max(len(xyz), 1)
    bufs = [np.zeros((cap, 4), np.float32) for _ in FEATURE_CLOUDS]
    cnt = [ctypes.c_int() for _ in FEATURE_CLOUDS]
    args = []
    for b, c in zip(bufs, cnt):
        args += [ctypes.c_void_p(b.ctypes.data), ctypes.byref(c)]
    rc = lib().orc_scan_registration(SENSORS[sensor], ctypes.c_double(minimum_range), ctypes.c_void_p(xyz.ctypes.data), len(xyz),
                                     tie_rule, cap, *args)
    assert rc == 0
    return {k: b[:c.value].copy() for k, b, c in zip(FEATURE_CLOUDS, bufs, cnt)}


def ring_of(sensor, xyz):
    xyz = _f32(xyz).reshape(-1, 3)
    out = np.zeros(len(xyz), np.int32)
    lib().orc_ring_of(SENSORS[sensor], ctypes.c_void_p(xyz.ctypes.data), len(xyz), ctypes.c_void_p(out.ctypes.data))
    return out


def knn(map_xyzi, q_xyz, method=0):
    m, q = _f32(map_xyzi).reshape(-1, 4), _f32(q_xyz).reshape(-1, 3)
    idx = np.zeros((len(q), 5), np.int32)
    d2 = np.zeros((len(q), 5), np.float32)
    lib().orc_knn(m.ctypes.data, len(m), q.ctypes.data, len(q), method, idx.ctypes.data, d2.ctypes.data)
    return idx, d2


def ref_knn(map_xyzi, q_xyz):
    R = ref_lib()
    assert R is not None
    m, q = _f32(map_xyzi).reshape(-1, 4), _f32(q_xyz).reshape(-1, 3)
    idx = np.zeros((len(q), 5), np.int32)
    d2 = np.zeros((len(q), 5), np.float32)
    R.ref_nanoflann_knn5(m.ctypes.data, len(m), q.ctypes.data, len(q), idx.ctypes.data, d2.ctypes.data)
    return idx, d2
