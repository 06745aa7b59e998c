"""Seeded synthetic inputs for the scan-to-map path (ctypes wrapper of harness/synth.cpp).

Input generation only: not the product path, not the oracle. See synth.cpp for what
each generator restates (scanRegistration.cpp:142-420 for the feature split).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libs2m_harness.so")

SENSORS = {"HDL64": 0, "VLP16": 1, "OS1-64": 2}
# minimum_range / mapping resolutions per launch file (launch/aloam_velodyne_HDL_64.launch:9-12,
# aloam_velodyne_VLP_16.launch:9-12, aloam_mulran.launch:9-12)
LAUNCH = {
    "HDL64": dict(minimum_range=5.0, line_res=0.4, plane_res=0.8),
    "VLP16": dict(minimum_range=0.1, line_res=0.2, plane_res=0.4),
    "OS1-64": dict(minimum_range=0.5, line_res=0.4, plane_res=0.8),
}


def build(force=False):
    src = os.path.join(_HERE, "synth.cpp")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-fopenmp", "-std=c++17", "-shared", "-fPIC", src, "-o", _LIB])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = ctypes.CDLL(_LIB)
        vp, ci, cd, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_uint64
        L.synth_trajectory.argtypes = [u64, ci, cd, vp]
        L.synth_odometry.argtypes = [u64, ci, vp, cd, cd, vp]
        L.synth_scan.argtypes = [u64, ci, vp, ci, cd, vp, ci]
        L.synth_features.argtypes = [ci, cd, vp, ci, vp, ci, vp, vp, ci, vp, vp, ci, vp]
        L.synth_surfaces.argtypes = [u64, cd, cd, cd, cd, vp, ci, vp, vp, ci, vp]
        _lib = L
    return _lib


def set_threads(n):
    """OpenMP threads used by scan() (torchrun exports OMP_NUM_THREADS=1 to every rank)"""
    lib().synth_set_threads(int(n))


def trajectory(seed, n_frames, step_m=1.0):
    poses = np.zeros((n_frames, 7), np.float64)
    lib().synth_trajectory(seed, n_frames, step_m, poses.ctypes.data)
    return poses


def odometry(seed, true_poses, sigma_t=0.02, sigma_r_deg=0.1):
    out = np.zeros_like(true_poses)
    tp = np.ascontiguousarray(true_poses, np.float64)
    lib().synth_odometry(seed, len(tp), tp.ctypes.data, sigma_t, np.deg2rad(sigma_r_deg), out.ctypes.data)
    return out


def scan(seed, sensor, pose7, frame, range_sigma=0.02):
    cap = 130000
    xyz = np.zeros((cap, 3), np.float32)
    p = np.ascontiguousarray(pose7, np.float64)
    n = lib().synth_scan(seed, SENSORS[sensor], p.ctypes.data, frame, range_sigma, xyz.ctypes.data, cap)
    assert n >= 0
    return xyz[:n].copy()


def features(sensor, xyz, minimum_range=None, want_full=False):
    """raw sweep -> (corner_last xyzi, surf_last xyzi[, full xyzi]) as laserMapping receives them."""
    if minimum_range is None:
        minimum_range = LAUNCH[sensor]["minimum_range"]
    n = len(xyz)
    corner = np.zeros((max(n, 1), 4), np.float32)
    surf = np.zeros((max(n, 1), 4), np.float32)
    full = np.zeros((max(n, 1), 4), np.float32) if want_full else None
    nc, ns, nf = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    xyz = np.ascontiguousarray(xyz, np.float32)
    rc = lib().synth_features(SENSORS[sensor], minimum_range, xyz.ctypes.data, n,
                              corner.ctypes.data, len(corner), ctypes.byref(nc),
                              surf.ctypes.data, len(surf), ctypes.byref(ns),
                              full.ctypes.data if want_full else None, len(surf), ctypes.byref(nf))
    assert rc == 0
    if want_full:
        return corner[:nc.value].copy(), surf[:ns.value].copy(), full[:nf.value].copy()
    return corner[:nc.value].copy(), surf[:ns.value].copy()


def sequence(seed, sensor, n_frames, step_m=1.0, sigma_t=0.02, sigma_r_deg=0.1, range_sigma=0.02):
    """A replayable sequence: true poses, odometry poses, per-frame (corner, surf)."""
    truth = trajectory(seed, n_frames, step_m)
    odom = odometry(seed, truth, sigma_t, sigma_r_deg)
    frames = []
    for f in range(n_frames):
        xyz = scan(seed, sensor, truth[f], f, range_sigma)
        frames.append(features(sensor, xyz))
    return truth, odom, frames


def surfaces(seed, half=525.0, surf_step=0.8, corner_step=0.4, sigma=0.02, cap=1 << 23):
    """Every world surface inside |x|,|y| <= half sampled directly (a saturated window
    without driving it) -> (corner xyzi, surf xyzi), world frame."""
    corner = np.zeros((cap, 4), np.float32)
    surf = np.zeros((cap, 4), np.float32)
    nc, ns = ctypes.c_int(), ctypes.c_int()
    rc = lib().synth_surfaces(seed, half, surf_step, corner_step, sigma, corner.ctypes.data, cap, ctypes.byref(nc),
                              surf.ctypes.data, cap, ctypes.byref(ns))
    assert rc == 0, "surface sample larger than cap"
    return corner[:nc.value].copy(), surf[:ns.value].copy()


def street_pose(xs, k, heading_quadrant=0):
    """A sensor pose [qx qy qz qw tx ty tz] on the centre line y' = 80 k of the street grid at street abscissa
    xs (never inside a building), 1.73 m above the terrain, heading along the street (+ quadrant * 90 deg)."""
    yaw = np.deg2rad(17.0)  # synth.cpp kGridYaw
    ys = 80.0 * k
    c, s = np.cos(yaw), np.sin(yaw)
    x, y = c * xs - s * ys, s * xs + c * ys
    g = 0.3 * np.sin(2 * np.pi * xs / 40.0) * np.cos(2 * np.pi * ys / 40.0)  # synth.cpp World::ground_s
    h = yaw + 0.5 * np.pi * heading_quadrant
    return np.array([0.0, 0.0, np.sin(h / 2), np.cos(h / 2), x, y, g + 1.73])
