"""Small end-to-end run for compute-sanitizer (memcheck): 3 slots, 4 frames, upload, knn debug, shard filters."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import harness
from __graft_entry__ import load_package
pkg = load_package()
truth, odom, frames = harness.sequence(3, "VLP16", 5, step_m=0.5)
B = 3
R = pkg.Registrar(0.2, 0.4, batch=B, cap_corner_in=4096, cap_surf_in=16384, cap_map_corner=1 << 15, cap_map_surf=1 << 16, trace=True)
for f in range(4):
    fr = [max(f - b, 0) for b in range(B)]
    corner = np.concatenate([frames[i][0] for i in fr]); surf = np.concatenate([frames[i][1] for i in fr])
    co = np.cumsum([0] + [len(frames[i][0]) for i in fr]).astype(np.int32)
    so = np.cumsum([0] + [len(frames[i][1]) for i in fr]).astype(np.int32)
    q = np.array([odom[i, :4] for i in fr]); t = np.array([odom[i, 4:] for i in fr])
    st, qo, to = R.register_batch(corner, co, surf, so, q, t, np.array([1 if f - b >= 0 else 0 for b in range(B)], np.int32))
print("status", st, R.batch_stats[0].n_edge[1], R.batch_stats[0].n_plane[1])
m0, m1 = R.map_download(0), R.map_download(1)
R2 = pkg.Registrar(0.2, 0.4, cap_corner_in=1 << 15, cap_surf_in=1 << 16, cap_map_corner=1 << 15, cap_map_surf=1 << 16, shard_rank=1, shard_world=2)
R2.map_upload(m0, m1)
idx, d2 = R2.debug_knn(1, odom[3, 4:], m1[:500, :3] + 0.05)
R2.register(frames[4][0], frames[4][1], odom[4, :4], odom[4, 4:])
print("ok", len(m0), len(m1), int((idx[:, 0] >= 0).sum()), R2.stats.n_plane[1], len(R.surround()), R.transform_cloud(frames[0][0]).shape)
