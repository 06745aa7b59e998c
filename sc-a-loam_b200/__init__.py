"""sc-a-loam_b200 -- B200-native scan-to-map registration (host-side Python view).

Thin ctypes binding of the C ABI in include/s2m.h (csrc/libs2m.so).  The product
is the CUDA library; this module only marshals numpy / torch buffers into it, the
way a ROS shim would marshal PointCloud2 payloads.

There is NO CPU fallback: importing works without a GPU (so the build and the
symbol table can be checked), but creating a context raises if the library or a
CUDA device is missing.  Nothing here imports or calls oracle/.

Directory name has a hyphen, so load it with importlib (see __graft_entry__.py):
    pkg = load_package()           # -> module 'sc_a_loam_b200'
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("S2M_LIB") or os.path.join(_HERE, "csrc", "libs2m.so")  # S2M_LIB: tuning variants only

S2M_OK = 0
S2M_MAP_TOO_SMALL = 1

# every symbol include/s2m.h declares (tests check the library exports all of them)
EXPORTS = [
    "s2m_default_params", "s2m_create", "s2m_destroy", "s2m_strerror", "s2m_last_error", "s2m_set_stream",
    "s2m_register", "s2m_register_batch", "s2m_register_batch_dev", "s2m_register_batch_submit",
    "s2m_register_batch_wait", "s2m_get_correction",
    "s2m_transform_cloud", "s2m_map_upload", "s2m_map_download", "s2m_pcd_write", "s2m_pcd_read", "s2m_checkpoint_save",
    "s2m_checkpoint_load", "s2m_get_local_map", "s2m_get_surround",
    "s2m_get_window", "s2m_debug_knn", "s2m_debug_knn_fallbacks", "s2m_debug_guard_check", "s2m_trace_cloud", "s2m_trace_knn", "s2m_trace_lm",
    "s2m_launch_count", "s2m_set_profiling", "s2m_k4_profile", "s2m_phase_profile", "s2m_shard_unique_id", "s2m_shard_slab", "s2m_shard_init",
    "s2m_shard_profile",
    "s2m_odom_create", "s2m_odom_step_batch",
    "s2m_fx_create", "s2m_fx_destroy", "s2m_fx_last_error", "s2m_fx_extract", "s2m_fx_offsets", "s2m_fx_download",
    "s2m_fx_device_cloud", "s2m_fx_launch_count",
]


class Params(ctypes.Structure):
    _fields_ = [("line_res", ctypes.c_float), ("plane_res", ctypes.c_float), ("device", ctypes.c_int),
                ("batch", ctypes.c_int), ("cap_corner_in", ctypes.c_int), ("cap_surf_in", ctypes.c_int),
                ("cap_map_corner", ctypes.c_int), ("cap_map_surf", ctypes.c_int),
                ("skip_optimization", ctypes.c_int), ("trace", ctypes.c_int),
                ("shard_rank", ctypes.c_int), ("shard_world", ctypes.c_int), ("lanes", ctypes.c_int)]


class Stats(ctypes.Structure):
    _fields_ = [("n_corner_in", ctypes.c_int), ("n_surf_in", ctypes.c_int),
                ("n_corner_ds", ctypes.c_int), ("n_surf_ds", ctypes.c_int),
                ("n_map_corner", ctypes.c_int), ("n_map_surf", ctypes.c_int),
                ("n_edge", ctypes.c_int * 2), ("n_plane", ctypes.c_int * 2),
                ("optimized", ctypes.c_int), ("lm_iters", ctypes.c_int * 2), ("lm_term", ctypes.c_int * 2),
                ("cost_initial", ctypes.c_double * 2), ("cost_final", ctypes.c_double * 2)]


class S2MError(RuntimeError):
    pass


_lib = None


def load_library(path=LIB_PATH):
    """dlopen csrc/libs2m.so and declare the prototypes. Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise S2MError("%s not built: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                       "There is no CPU fallback." % path)
    L = ctypes.CDLL(path)
    vp, ci, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
    L.s2m_default_params.argtypes = [ctypes.POINTER(Params)]
    L.s2m_default_params.restype = None
    L.s2m_create.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(vp)]
    L.s2m_destroy.argtypes = [vp]
    L.s2m_destroy.restype = None
    L.s2m_strerror.argtypes = [ci]
    L.s2m_strerror.restype = ctypes.c_char_p
    L.s2m_last_error.argtypes = [vp]
    L.s2m_last_error.restype = ctypes.c_char_p
    L.s2m_set_stream.argtypes = [vp, vp]
    L.s2m_register.argtypes = [vp, vp, ci, vp, ci, vp, vp, vp, vp, vp]
    L.s2m_register_batch.argtypes = [vp] + [vp] * 11
    L.s2m_register_batch_dev.argtypes = [vp] + [vp] * 11
    L.s2m_register_batch_submit.argtypes = [vp] + [vp] * 11 + [ctypes.c_int]
    L.s2m_register_batch_wait.argtypes = [vp]
    L.s2m_get_correction.argtypes = [vp, ci, vp, vp]
    L.s2m_transform_cloud.argtypes = [vp, ci, vp, ci, vp]
    L.s2m_map_upload.argtypes = [vp, ci, vp, ci, vp, ci]
    L.s2m_map_download.argtypes = [vp, ci, ci, vp, ci]
    L.s2m_odom_create.argtypes = [ci, ci, ci, ci, ci, ci, ci, vp]
    L.s2m_odom_step_batch.argtypes = [vp] + [vp] * 8 + [ci] + [vp] * 4
    L.s2m_fx_create.argtypes = [vp, vp]
    L.s2m_fx_destroy.argtypes = [vp]
    L.s2m_fx_destroy.restype = None
    L.s2m_fx_last_error.argtypes = [vp]
    L.s2m_fx_last_error.restype = ctypes.c_char_p
    L.s2m_fx_extract.argtypes = [vp, vp, vp, ci]
    L.s2m_fx_offsets.argtypes = [vp, ci, vp]
    L.s2m_fx_download.argtypes = [vp, ci, vp, ci]
    L.s2m_fx_device_cloud.argtypes = [vp, ci]
    L.s2m_fx_device_cloud.restype = vp
    L.s2m_fx_launch_count.argtypes = [vp]
    L.s2m_fx_launch_count.restype = ctypes.c_longlong
    L.s2m_pcd_write.argtypes = [ctypes.c_char_p, vp, ci]
    L.s2m_pcd_read.argtypes = [ctypes.c_char_p, vp, ci]
    L.s2m_checkpoint_save.argtypes = [vp, ci, ctypes.c_char_p]
    L.s2m_checkpoint_load.argtypes = [vp, ci, ctypes.c_char_p]
    L.s2m_get_local_map.argtypes = [vp, ci, ci, vp, vp, ci]
    L.s2m_get_surround.argtypes = [vp, ci, vp, ci]
    L.s2m_get_window.argtypes = [vp, ci, vp]
    L.s2m_debug_knn.argtypes = [vp, ci, ci, vp, vp, ci, vp, vp]
    L.s2m_trace_cloud.argtypes = [vp, ci, ci, vp, ci]
    L.s2m_trace_knn.argtypes = [vp, ci, ci, ci, vp, vp, vp, ci]
    L.s2m_trace_lm.argtypes = [vp, ci, ci, vp, vp, vp, vp, vp]
    L.s2m_launch_count.argtypes = [vp]
    L.s2m_launch_count.restype = cll
    L.s2m_set_profiling.argtypes = [vp, ci]
    L.s2m_k4_profile.argtypes = [vp, ci, vp, vp, vp]
    L.s2m_debug_knn_fallbacks.argtypes = [vp]
    L.s2m_debug_knn_fallbacks.restype = ctypes.c_longlong
    L.s2m_phase_profile.argtypes = [vp, ci, vp]
    L.s2m_shard_unique_id.argtypes = [vp]
    L.s2m_shard_slab.argtypes = [ci, ci, vp, vp]
    L.s2m_shard_init.argtypes = [vp, vp]
    L.s2m_shard_profile.argtypes = [vp, ci, vp, vp]
    _lib = L
    return L


def shard_slab(rank, world):
    """[lo, hi) of world x owned by `rank` in a `world`-way sharded map (no GPU needed)."""
    lo, hi = ctypes.c_float(), ctypes.c_float()
    rc = load_library().s2m_shard_slab(rank, world, ctypes.byref(lo), ctypes.byref(hi))
    if rc != 0:
        raise S2MError("bad shard rank/world")
    return lo.value, hi.value


class Odometer:
    """laserOdometry.cpp:220-591 for `batch` independent sequences on the device (include/s2m.h, s2m_odom_*)."""

    def __init__(self, batch=1, cap_sharp=4096, cap_flat=8192, cap_less_sharp=1 << 14, cap_less_flat=1 << 16, device=0,
                 trace=False):
        self.L = load_library()
        h = ctypes.c_void_p()
        rc = self.L.s2m_odom_create(device, batch, cap_sharp, cap_flat, cap_less_sharp, cap_less_flat, 1 if trace else 0,
                                    ctypes.byref(h))
        if rc != 0:
            raise S2MError("s2m_odom_create: %s (no CUDA device? the product has no CPU path)" % self.L.s2m_strerror(rc).decode())
        self.h, self.batch = h, batch
        self.para = np.zeros((batch, 7))
        self.counts = np.zeros((batch, 4), np.int32)

    def close(self):
        if self.h:
            self.L.s2m_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise S2MError("%s: %s" % (self.L.s2m_strerror(rc).decode(), self.L.s2m_last_error(self.h).decode()))
        return rc

    def step_batch(self, sharp, sharp_off, flat, flat_off, less_sharp, ls_off, less_flat, lf_off, device_ptrs=False):
        """four packed xyzi clouds (host arrays, or raw device pointers if device_ptrs) with B+1 offsets each
        -> (q_w_curr[B,4], t_w_curr[B,3]); self.para[B,7], self.counts[B,4]"""
        B = self.batch
        offs = [np.ascontiguousarray(o, np.int32) for o in (sharp_off, flat_off, ls_off, lf_off)]
        assert all(len(o) == B + 1 for o in offs)
        if device_ptrs:
            ptrs = [ctypes.c_void_p(int(x)) for x in (sharp, flat, less_sharp, less_flat)]
        else:
            keep = [_f32(x).reshape(-1, 4) for x in (sharp, flat, less_sharp, less_flat)]
            ptrs = [ctypes.c_void_p(x.ctypes.data) for x in keep]
        q, t = np.zeros((B, 4)), np.zeros((B, 3))
        rc = self.L.s2m_odom_step_batch(self.h, ptrs[0], offs[0].ctypes.data, ptrs[1], offs[1].ctypes.data, ptrs[2],
                                        offs[2].ctypes.data, ptrs[3], offs[3].ctypes.data, 1 if device_ptrs else 0,
                                        q.ctypes.data, t.ctypes.data, self.para.ctypes.data, self.counts.ctypes.data)
        self._check(rc)
        return q, t

    def step(self, sharp, flat, less_sharp, less_flat):
        """single sequence, host arrays -> (q_w_curr[4], t_w_curr[3])"""
        a = [_f32(x).reshape(-1, 4) for x in (sharp, flat, less_sharp, less_flat)]
        q, t = self.step_batch(a[0], [0, len(a[0])], a[1], [0, len(a[1])], a[2], [0, len(a[2])], a[3], [0, len(a[3])])
        return q[0], t[0]

    def trace(self, outer, cls, slot=0, cap=1 << 16):
        """correspondences of pass `outer` for the sharp (cls 0) / flat (cls 1) queries of the last step:
        (idx[n,3] = closest, second, third index into the previous cloud or -1, used[n])"""
        idx = np.zeros((cap, 5), np.int32)
        d2 = np.zeros((cap, 5), np.float32)
        used = np.zeros(cap, np.uint8)
        n = self._check(self.L.s2m_trace_knn(self.h, slot, outer, cls, idx.ctypes.data, d2.ctypes.data, used.ctypes.data, cap))
        return idx[:n, :3].copy(), used[:n].astype(bool)

    def launch_count(self):
        return self.L.s2m_launch_count(self.h)


class FxParams(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int), ("batch", ctypes.c_int), ("cap_points", ctypes.c_int), ("sensor", ctypes.c_int),
                ("minimum_range", ctypes.c_double)]


SENSORS = {"HDL64": 0, "VLP16": 1, "OS1-64": 2, "HDL32": 3}
FX_CLOUDS = {"full": 0, "sharp": 1, "less_sharp": 2, "flat": 3, "less_flat": 4}


class FeatureExtractor:
    """scanRegistration.cpp:116-454 for a batch of raw sweeps on the device (include/s2m.h, s2m_fx_*)."""

    def __init__(self, sensor, minimum_range, batch=1, cap_points=1 << 17, device=0):
        self.L = load_library()
        p = FxParams(device, batch, cap_points, SENSORS[sensor], float(minimum_range))
        h = ctypes.c_void_p()
        rc = self.L.s2m_fx_create(ctypes.byref(p), ctypes.byref(h))
        if rc != 0:
            raise S2MError("s2m_fx_create: %s (no CUDA device? the product has no CPU path)" % self.L.s2m_strerror(rc).decode())
        self.h, self.batch = h, batch

    def close(self):
        if self.h:
            self.L.s2m_fx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise S2MError("%s: %s" % (self.L.s2m_strerror(rc).decode(), self.L.s2m_fx_last_error(self.h).decode()))
        return rc

    def extract(self, xyz, off, device=False):
        """xyz: packed (n,3) float32 host array (or a raw device pointer if device); off: B+1 sweep offsets."""
        off = np.ascontiguousarray(off, np.int32)
        assert len(off) == self.batch + 1
        if device:
            ptr = ctypes.c_void_p(int(xyz))
        else:
            xyz = _f32(xyz).reshape(-1, 3)
            ptr = ctypes.c_void_p(xyz.ctypes.data)
        self._check(self.L.s2m_fx_extract(self.h, ptr, off.ctypes.data, 1 if device else 0))

    def offsets(self, which):
        off = np.zeros(self.batch + 1, np.int32)
        self._check(self.L.s2m_fx_offsets(self.h, FX_CLOUDS[which], off.ctypes.data))
        return off

    def cloud(self, which):
        """-> (packed (m,4) xyzi over all sweeps, per-sweep offsets)"""
        n = self._check(self.L.s2m_fx_download(self.h, FX_CLOUDS[which], None, 0))
        out = np.zeros((max(n, 1), 4), np.float32)
        self._check(self.L.s2m_fx_download(self.h, FX_CLOUDS[which], out.ctypes.data, n))
        return out[:n], self.offsets(which)

    def device_cloud(self, which):
        return self.L.s2m_fx_device_cloud(self.h, FX_CLOUDS[which])

    def launch_count(self):
        return self.L.s2m_fx_launch_count(self.h)


def pcd_write(path, xyzi):
    """PCD v0.7 binary x y z intensity float32, as the reference's savePCDFileBinary writes it (no GPU needed)."""
    a = _f32(xyzi).reshape(-1, 4)
    rc = load_library().s2m_pcd_write(os.fsencode(path), a.ctypes.data, len(a))
    if rc != 0:
        raise S2MError("cannot write %s" % path)


def pcd_read(path):
    L = load_library()
    n = L.s2m_pcd_read(os.fsencode(path), None, 0)
    if n < 0:
        raise S2MError("cannot read %s" % path)
    out = np.zeros((max(n, 1), 4), np.float32)
    if L.s2m_pcd_read(os.fsencode(path), out.ctypes.data, n) != n:
        raise S2MError("cannot read %s" % path)
    return out[:n]


def default_params():
    p = Params()
    load_library().s2m_default_params(ctypes.byref(p))
    return p


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


class Registrar:
    """One context = `batch` independent laserMapping instances on one GPU.

    Mirrors the reference node's state and per-frame step (laserMapping.cpp:232-906):
    register() is one pass of process() for one sequence, register_batch() one pass
    for every slot.
    """

    def __init__(self, line_res=0.4, plane_res=0.8, device=0, batch=1, cap_corner_in=16384,
                 cap_surf_in=131072, cap_map_corner=1 << 19, cap_map_surf=1 << 20,
                 skip_optimization=False, trace=False, shard_rank=0, shard_world=1, lanes=0):
        self.L = load_library()
        p = default_params()
        p.line_res, p.plane_res, p.device, p.batch = line_res, plane_res, device, batch
        p.cap_corner_in, p.cap_surf_in = cap_corner_in, cap_surf_in
        p.cap_map_corner, p.cap_map_surf = cap_map_corner, cap_map_surf
        p.skip_optimization, p.trace = int(skip_optimization), int(trace)
        p.shard_rank, p.shard_world = shard_rank, shard_world
        p.lanes = lanes
        self.params = p
        self.batch = batch
        h = ctypes.c_void_p()
        rc = self.L.s2m_create(ctypes.byref(p), ctypes.byref(h))
        if rc != 0 or not h:
            raise S2MError("s2m_create failed: %s (no CUDA device? there is no CPU fallback)"
                           % self.L.s2m_strerror(rc).decode())
        self.h = h
        self.stats = Stats()
        self._bstats = (Stats * batch)()
        self._status = np.zeros(batch, np.int32)
        self._jobs = []

    def close(self):
        if getattr(self, "h", None):
            self.L.s2m_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc < 0:
            raise S2MError("%s: %s" % (self.L.s2m_strerror(rc).decode(), self.L.s2m_last_error(self.h).decode()))
        return rc

    def set_stream(self, cuda_stream_handle):
        self._check(self.L.s2m_set_stream(self.h, ctypes.c_void_p(cuda_stream_handle)))

    # ---- one sequence -------------------------------------------------------
    def register(self, corner, surf, q_wodom, t_wodom):
        """-> (status, q_w[4], t_w[3]); status 0 ok, 1 map too small (pose = guess)."""
        corner, surf = _f32(corner).reshape(-1, 4), _f32(surf).reshape(-1, 4)
        q, t = _f64(q_wodom), _f64(t_wodom)
        qo, to = np.zeros(4), np.zeros(3)
        rc = self.L.s2m_register(self.h, corner.ctypes.data, len(corner), surf.ctypes.data, len(surf),
                                 q.ctypes.data, t.ctypes.data, qo.ctypes.data, to.ctypes.data,
                                 ctypes.byref(self.stats))
        return self._check(rc), qo, to

    # ---- every slot -----------------------------------------------------------
    def register_batch(self, corner, corner_off, surf, surf_off, q_wodom, t_wodom, active=None, device_ptrs=False):
        """corner/surf: packed (n,4) float32 host arrays, or raw device pointers (ints) if device_ptrs.
        -> (status[B], q_w[B,4], t_w[B,3]); per-slot stats in self.batch_stats."""
        B = self.batch
        co, so = np.ascontiguousarray(corner_off, np.int32), np.ascontiguousarray(surf_off, np.int32)
        assert len(co) == B + 1 and len(so) == B + 1
        q, t = _f64(q_wodom).reshape(B, 4), _f64(t_wodom).reshape(B, 3)
        qo, to = np.zeros((B, 4)), np.zeros((B, 3))
        act = None if active is None else np.ascontiguousarray(active, np.int32)
        if device_ptrs:
            cp, sp = ctypes.c_void_p(int(corner)), ctypes.c_void_p(int(surf))
            fn = self.L.s2m_register_batch_dev
        else:
            corner, surf = _f32(corner).reshape(-1, 4), _f32(surf).reshape(-1, 4)
            cp, sp = corner.ctypes.data, surf.ctypes.data
            fn = self.L.s2m_register_batch
        rc = fn(self.h, cp, co.ctypes.data, sp, so.ctypes.data, q.ctypes.data, t.ctypes.data,
                None if act is None else act.ctypes.data, qo.ctypes.data, to.ctypes.data,
                ctypes.cast(self._bstats, ctypes.c_void_p), self._status.ctypes.data)
        self._check(rc)
        return self._status.copy(), qo, to

    def register_batch_ptr(self, corner_ptr, corner_off, surf_ptr, surf_off, q_wodom, t_wodom, device):
        """Raw-pointer variant (ints): host pointers -> s2m_register_batch, device pointers -> _dev."""
        B = self.batch
        co, so = np.ascontiguousarray(corner_off, np.int32), np.ascontiguousarray(surf_off, np.int32)
        q, t = _f64(q_wodom).reshape(B, 4), _f64(t_wodom).reshape(B, 3)
        qo, to = np.zeros((B, 4)), np.zeros((B, 3))
        fn = self.L.s2m_register_batch_dev if device else self.L.s2m_register_batch
        rc = fn(self.h, ctypes.c_void_p(int(corner_ptr)), co.ctypes.data, ctypes.c_void_p(int(surf_ptr)), so.ctypes.data,
                q.ctypes.data, t.ctypes.data, None, qo.ctypes.data, to.ctypes.data,
                ctypes.cast(self._bstats, ctypes.c_void_p), self._status.ctypes.data)
        self._check(rc)
        return self._status, qo, to

    def submit(self, corner_ptr, corner_off, surf_ptr, surf_off, q_wodom, t_wodom, device=False):
        """Asynchronous batch call (lanes >= 1): raw pointers (ints) to the packed clouds, which must stay
        valid until the matching wait().  At most two frames in flight."""
        B = self.batch
        job = dict(co=np.ascontiguousarray(corner_off, np.int32), so=np.ascontiguousarray(surf_off, np.int32),
                   q=_f64(q_wodom).reshape(B, 4), t=_f64(t_wodom).reshape(B, 3), qo=np.zeros((B, 4)), to=np.zeros((B, 3)),
                   status=np.zeros(B, np.int32), stats=(Stats * B)())
        rc = self.L.s2m_register_batch_submit(self.h, ctypes.c_void_p(int(corner_ptr)), job["co"].ctypes.data,
                                              ctypes.c_void_p(int(surf_ptr)), job["so"].ctypes.data, job["q"].ctypes.data,
                                              job["t"].ctypes.data, None, job["qo"].ctypes.data, job["to"].ctypes.data,
                                              ctypes.cast(job["stats"], ctypes.c_void_p), job["status"].ctypes.data,
                                              1 if device else 0)
        self._check(rc)
        self._jobs.append(job)

    def wait(self):
        """Completes the oldest submitted frame -> (status[B], q_w[B,4], t_w[B,3])."""
        rc = self.L.s2m_register_batch_wait(self.h)
        job = self._jobs.pop(0)
        self._check(rc)
        ctypes.memmove(self._bstats, job["stats"], ctypes.sizeof(self._bstats))
        return job["status"], job["qo"], job["to"]

    @property
    def batch_stats(self):
        return self._bstats

    # ---- state ----------------------------------------------------------------
    def correction(self, slot=0):
        q, t = np.zeros(4), np.zeros(3)
        self._check(self.L.s2m_get_correction(self.h, slot, q.ctypes.data, t.ctypes.data))
        return q, t

    def window(self, slot=0):
        c = np.zeros(3, np.int32)
        self._check(self.L.s2m_get_window(self.h, slot, c.ctypes.data))
        return c

    def transform_cloud(self, pts, slot=0):
        pts = _f32(pts).reshape(-1, 4)
        out = np.zeros_like(pts)
        self._check(self.L.s2m_transform_cloud(self.h, slot, pts.ctypes.data, len(pts), out.ctypes.data))
        return out

    def map_upload(self, corner, surf, slot=0):
        corner, surf = _f32(corner).reshape(-1, 4), _f32(surf).reshape(-1, 4)
        return self._check(self.L.s2m_map_upload(self.h, slot, corner.ctypes.data, len(corner),
                                                 surf.ctypes.data, len(surf)))

    def map_download(self, cls, slot=0):
        n = self._check(self.L.s2m_map_download(self.h, slot, cls, None, 0))
        out = np.zeros((max(n, 1), 4), np.float32)
        self._check(self.L.s2m_map_download(self.h, slot, cls, out.ctypes.data, n))
        return out[:n]

    def checkpoint_save(self, prefix, slot=0):
        self._check(self.L.s2m_checkpoint_save(self.h, slot, os.fsencode(prefix)))

    def checkpoint_load(self, prefix, slot=0):
        """-> number of points that fell outside the restored window"""
        return self._check(self.L.s2m_checkpoint_load(self.h, slot, os.fsencode(prefix)))

    def local_map(self, cls, centre_t, slot=0):
        c = _f64(centre_t)
        cap = int(self.params.cap_map_corner if cls == 0 else self.params.cap_map_surf)
        out = np.zeros((cap, 4), np.float32)
        n = self._check(self.L.s2m_get_local_map(self.h, slot, cls, c.ctypes.data, out.ctypes.data, cap))
        return out[:n].copy()

    def surround(self, slot=0):
        n = self._check(self.L.s2m_get_surround(self.h, slot, None, 0))
        out = np.zeros((max(n, 1), 4), np.float32)
        self._check(self.L.s2m_get_surround(self.h, slot, out.ctypes.data, n))
        return out[:n]

    def debug_knn(self, cls, centre_t, q_xyz, slot=0):
        q = _f32(q_xyz).reshape(-1, 3)
        c = _f64(centre_t)
        idx = np.zeros((len(q), 5), np.int32)
        d2 = np.zeros((len(q), 5), np.float32)
        self._check(self.L.s2m_debug_knn(self.h, slot, cls, c.ctypes.data, q.ctypes.data, len(q),
                                         idx.ctypes.data, d2.ctypes.data))
        return idx, d2

    # ---- trace of the last call (trace=True) --------------------------------------
    def trace_cloud(self, cls, slot=0):
        cap = int(self.params.cap_corner_in if cls == 0 else self.params.cap_surf_in)
        out = np.zeros((cap, 4), np.float32)
        n = self._check(self.L.s2m_trace_cloud(self.h, slot, cls, out.ctypes.data, cap))
        return out[:n].copy()

    def trace_knn(self, outer, cls, slot=0):
        cap = int(self.params.cap_corner_in if cls == 0 else self.params.cap_surf_in)
        idx = np.zeros((cap, 5), np.int32)
        d2 = np.zeros((cap, 5), np.float32)
        used = np.zeros(cap, np.uint8)
        n = self._check(self.L.s2m_trace_knn(self.h, slot, outer, cls, idx.ctypes.data, d2.ctypes.data,
                                             used.ctypes.data, cap))
        return idx[:n].copy(), d2[:n].copy(), used[:n].copy()

    def trace_lm(self, outer, slot=0):
        pose, sums, iters = np.zeros(7), np.zeros(28), np.zeros((4, 6))
        n, term = ctypes.c_int(), ctypes.c_int()
        self._check(self.L.s2m_trace_lm(self.h, slot, outer, pose.ctypes.data, sums.ctypes.data,
                                        iters.ctypes.data, ctypes.byref(n), ctypes.byref(term)))
        return pose, sums, iters, n.value, term.value

    # ---- sharded map (BASELINE config 5) ----------------------------------------------------
    @staticmethod
    def shard_unique_id():
        buf = ctypes.create_string_buffer(128)
        rc = load_library().s2m_shard_unique_id(buf)
        if rc != 0:
            raise S2MError("s2m_shard_unique_id failed (NCCL not loadable)")
        return buf.raw

    def shard_init(self, id128):
        self._check(self.L.s2m_shard_init(self.h, ctypes.create_string_buffer(id128, 128)))

    def shard_profile(self, reset=True):
        ms, n = ctypes.c_double(), ctypes.c_longlong()
        self._check(self.L.s2m_shard_profile(self.h, int(reset), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def knn_fallbacks(self):
        """queries the grouped kNN kernel handed to the thread-per-query search so far (tuning counter)"""
        return int(self.L.s2m_debug_knn_fallbacks(self.h))

    def guard_check(self):
        """debug (S2M_GUARD_BYTES set at create): guard words overwritten so far, 0 = no out-of-bounds store"""
        return self._check(self.L.s2m_debug_guard_check(self.h))

    # ---- measurement ----------------------------------------------------------------
    def launch_count(self):
        return int(self.L.s2m_launch_count(self.h))

    def set_profiling(self, on=True, count_candidates=False):
        """events at the phase boundaries of every frame; count_candidates: also the candidate term of K4's bytes"""
        self._check(self.L.s2m_set_profiling(self.h, (2 if count_candidates else 1) if on else 0))

    PHASES = ["input", "voxel", "index", "associate", "solve", "update", "readback"]

    def phase_profile(self, reset=True):
        ms = np.zeros(len(self.PHASES))
        self._check(self.L.s2m_phase_profile(self.h, int(reset), ms.ctypes.data))
        return dict(zip(self.PHASES, ms.tolist()))

    def k4_profile(self, reset=True):
        ms, n, b = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
        self._check(self.L.s2m_k4_profile(self.h, int(reset), ctypes.byref(ms), ctypes.byref(n), ctypes.byref(b)))
        return ms.value, n.value, b.value
