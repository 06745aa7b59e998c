#!/usr/bin/env python
"""Generates tests/golden/*.npz -- run in the authoring container (needs /root/reference for the
nanoflann vectors).  Seeds are fixed; re-running must reproduce the committed files bit for bit.

  knn_nanoflann.npz : kNN(5) answers of the KD-tree the REFERENCE TREE vendors
                      (/root/reference/include/scancontext/nanoflann.hpp via oracle/_ref) on a small
                      voxel-filtered map -- vectors produced by reference code, not by the oracle.
  voxel_grid.npz    : oracle VoxelGrid outputs (PCL 1.8 semantics, stable order) on seeded clouds.
  vlp16_stream.npz  : inputs + oracle poses / counters of a 6-frame VLP-16 stream (0.2 / 0.4).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import harness  # noqa: E402
import oracle  # noqa: E402


def main():
    oracle.build()
    rng = np.random.default_rng(20261018)
    # ---- kNN vectors from the reference tree's nanoflann ----
    truth, odom, frames = harness.sequence(20261018, "VLP16", 4, step_m=0.5)
    O = oracle.Oracle(0.2, 0.4, skip_optimization=True)
    for f in range(4):
        O.register(frames[f][0], frames[f][1], truth[f, :4], truth[f, 4:])
    surf_map = O.get_map(1)
    sel = rng.choice(len(surf_map), 6000, replace=False)
    sel.sort()
    mp = surf_map[sel]
    q = (mp[rng.integers(0, len(mp), 1500), :3] + rng.normal(size=(1500, 3)).astype(np.float32) * 0.25).astype(np.float32)
    q = np.r_[q, rng.uniform(-60, 60, (100, 3)).astype(np.float32)]
    assert oracle.ref_lib() is not None, "oracle/_ref not built: /root/reference must be mounted"
    idx, d2 = oracle.ref_knn(mp, q)
    np.savez_compressed(os.path.join(HERE, "knn_nanoflann.npz"), map_xyzi=mp, q_xyz=q, idx=idx, d2=d2)
    # ---- voxel grid ----
    clouds, outs = {}, {}
    for name, (n, span, leaf) in {"a": (4000, 30.0, 0.4), "b": (6000, 80.0, 0.8), "c": (1500, 6.0, 0.2)}.items():
        c = rng.uniform(-span, span, (n, 4)).astype(np.float32)
        c[:, 2] *= 0.05
        clouds[name] = c
        outs[name] = oracle.voxel_grid(c, leaf)
    np.savez_compressed(os.path.join(HERE, "voxel_grid.npz"), **{"in_" + k: v for k, v in clouds.items()},
                        **{"out_" + k: v for k, v in outs.items()}, leaves=np.array([0.4, 0.8, 0.2], np.float32))
    # ---- VLP-16 stream ----
    truth, odom, frames = harness.sequence(7, "VLP16", 6, step_m=0.5)
    O = oracle.Oracle(0.2, 0.4)
    poses, counters = [], []
    for f, (c, s) in enumerate(frames):
        rc, qq, tt = O.register(c, s, odom[f, :4], odom[f, 4:])
        poses.append(np.r_[qq, tt])
        st = O.stats
        counters.append([rc, st.n_corner_ds, st.n_surf_ds, st.n_map_corner, st.n_map_surf, st.n_edge[0], st.n_edge[1],
                         st.n_plane[0], st.n_plane[1], st.lm_iters[0], st.lm_iters[1]])
    # thin the clouds so the fixture stays small: keep every 3rd surf point
    np.savez_compressed(os.path.join(HERE, "vlp16_stream.npz"), odom=odom, poses=np.array(poses), counters=np.array(counters),
                        **{"corner_%d" % f: frames[f][0] for f in range(6)}, **{"surf_%d" % f: frames[f][1] for f in range(6)})
    for f in os.listdir(HERE):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
