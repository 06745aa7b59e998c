#!/usr/bin/env python
"""Generates tests/golden/kaist03.npz from the REAL data the reference ships
(/root/reference/utils/sample_data/KAIST03: 21 OS1-64 keyframe scans written by the reference pipeline at
laserPosegraphOptimization.cpp:695 from laserMapping's /velodyne_cloud_registered_local, and the poses
it saved for them, laserPosegraphOptimization.cpp:236-259).  Run in the authoring container only.

Content (float32 clouds are xyzi):
  map_corner, map_surf : the mapping node's map after scans 0..9 were inserted AT THE REFERENCE'S OWN
                         POSES (oracle, skip_optimization: rows B, C, V, I, W only)
  corner_k, surf_k     : feature clouds of scans k = 10..14 (the restated scanRegistration split of
                         harness/, ring = the OS1-64 rule of scanRegistration.cpp:205-213)
  ref_q, ref_t         : the reference's saved poses of scans 10..14 (x,y,z,w / metres)
  guess_q, guess_t     : those poses perturbed by U(-0.2,0.2) m and U(-1,1) deg per axis (seed 20261018)
  oracle_q, oracle_t   : the oracle's registered poses from those guesses (stream: each frame is inserted
                         at its own registered pose before the next one)
kaist03_scan10.npz: xyz of real scan 10 and the ring number the reference stored in its intensity channel
(scanRegistration.cpp:251-252), for tests/test_scan_registration.py.
The test then asserts (a) the oracle -- and the CUDA path -- pull every perturbed guess back to within
6 cm / 0.45 deg of the pose the reference itself saved, and (b) CUDA == oracle to the usual tolerance.
"""
import os
import sys

import numpy as np
from scipy.spatial.transform import Rotation as Rot

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import harness  # noqa: E402
import oracle  # noqa: E402

SRC = "/root/reference/utils/sample_data/KAIST03/"
N_MAP, FRAMES = 10, range(10, 15)


def read_pcd(path):
    raw = open(path, "rb").read()
    start = raw.index(b"DATA binary\n") + len(b"DATA binary\n")
    header = raw[:start].decode()
    assert "FIELDS x y z intensity" in header and "SIZE 4 4 4 4" in header
    n = int([ln for ln in header.splitlines() if ln.startswith("POINTS")][0].split()[1])
    return np.frombuffer(raw[start:start + 16 * n], np.float32).reshape(n, 4).copy()


def main():
    oracle.build()
    poses = np.loadtxt(SRC + "optimized_poses.txt")[:21].reshape(21, 3, 4)
    q_ref = np.array([Rot.from_matrix(p[:, :3]).as_quat() for p in poses])
    t_ref = poses[:, :, 3].copy()
    feats = [harness.features("OS1-64", read_pcd(SRC + "Scans/%06d.pcd" % k)[:, :3]) for k in range(max(FRAMES) + 1)]
    builder = oracle.Oracle(0.4, 0.8, skip_optimization=True)      # aloam_mulran.launch:11-12
    for k in range(N_MAP):
        builder.register(feats[k][0], feats[k][1], q_ref[k], t_ref[k])
    out = {"map_corner": builder.get_map(0), "map_surf": builder.get_map(1)}
    rng = np.random.default_rng(20261018)
    O = oracle.Oracle(0.4, 0.8)
    O.map_upload(out["map_corner"], out["map_surf"])
    gq, gt, oq, ot = [], [], [], []
    for k in FRAMES:
        dq = Rot.from_rotvec(np.deg2rad(rng.uniform(-1, 1, 3)))
        gq.append((Rot.from_quat(q_ref[k]) * dq).as_quat())
        gt.append(t_ref[k] + rng.uniform(-0.2, 0.2, 3))
        rc, q, t = O.register(feats[k][0], feats[k][1], gq[-1], gt[-1])
        assert rc == 0
        oq.append(q)
        ot.append(t)
        out["corner_%d" % k], out["surf_%d" % k] = feats[k]
        err_t = np.linalg.norm(t - t_ref[k])
        err_r = np.rad2deg((Rot.from_quat(q).inv() * Rot.from_quat(q_ref[k])).magnitude())
        print("scan %d: guess %.3f m -> %.3f m, %.2f deg from the reference's pose" % (k, np.linalg.norm(gt[-1] - t_ref[k]), err_t, err_r))
    out.update(frames=np.array(list(FRAMES)), ref_q=q_ref[list(FRAMES)], ref_t=t_ref[list(FRAMES)], guess_q=np.array(gq),
               guess_t=np.array(gt), oracle_q=np.array(oq), oracle_t=np.array(ot))
    path = os.path.join(HERE, "kaist03.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
    # one real sweep for the feature-extraction tests: xyz of scan 10 in the order the reference stored it
    raw = read_pcd(SRC + "Scans/000010.pcd")
    path = os.path.join(HERE, "kaist03_scan10.npz")
    nxt = read_pcd(SRC + "Scans/000011.pcd")   # the following keyframe, for the odometry test
    np.savez_compressed(path, xyz=raw[:, :3].copy(), ring=np.rint(raw[:, 3]).astype(np.int8), xyz_next=nxt[:, :3].copy())
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__" and "--keyframes" not in sys.argv:
    main()


def make_keyframes():
    """tests/golden/kaist03_keyframes.npz: xyz (float32, exact) of ALL 21 real keyframe scans, packed with offsets,
    and the poses the reference saved for them (optimized_poses.txt rows 0..20) -- the input of the -m gpu test that
    runs the whole chain (features -> odometry -> mapping) on the device (tests/test_odometry.py)."""
    poses = np.loadtxt(SRC + "optimized_poses.txt")[:21].reshape(21, 3, 4)
    scans = [read_pcd(SRC + "Scans/%06d.pcd" % k)[:, :3] for k in range(21)]
    off = np.cumsum([0] + [len(s) for s in scans]).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "kaist03_keyframes.npz"), xyz=np.concatenate(scans).astype(np.float32),
                        off=off, ref_poses=poses)
    print("kaist03_keyframes.npz:", off[-1], "points")


if __name__ == "__main__" and "--keyframes" in sys.argv:
    make_keyframes()
