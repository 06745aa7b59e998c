"""Worker of tests/test_sharded.py: run under torchrun with 2+ ranks, one GPU each.
Every rank holds the x-slab of the map it owns (+1 m halo); sums are allreduced per evaluation."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import harness  # noqa: E402
import oracle  # noqa: E402
from conftest import load_package, rot_angle  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    s2m = load_package()
    truth, odom, frames = harness.sequence(20261018, "HDL64", 12)
    O = oracle.Oracle(0.4, 0.8)
    R = s2m.Registrar(0.4, 0.8, device=local, shard_rank=rank, shard_world=world)
    ids = [s2m.Registrar.shard_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    R.shard_init(ids[0])
    R.set_profiling(True)
    worst_t = worst_r = 0.0
    for f, (c, s) in enumerate(frames):
        rc, q, t = R.register(c, s, odom[f, :4], odom[f, 4:])
        ro, qo, to = O.register(c, s, odom[f, :4], odom[f, 4:])
        assert rc == ro, (f, rc, ro)
        worst_t = max(worst_t, float(np.linalg.norm(t - to)))
        worst_r = max(worst_r, rot_angle(q, qo))
        # every rank must hold the same pose, bit for bit
        mine = torch.tensor(np.r_[q, t], device="cuda")
        ref = mine.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(mine, ref), f
        assert (R.stats.n_edge[1], R.stats.n_plane[1]) == (O.stats.n_edge[1], O.stats.n_plane[1]), f
    n_local = torch.tensor([len(R.map_download(0)) + len(R.map_download(1))], device="cuda")
    n_all = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(n_all, n_local)
    n_full = len(O.get_map(0)) + len(O.get_map(1))
    ms, n = R.shard_profile()
    if rank == 0:
        parts = [int(x.item()) for x in n_all]
        print("SHARDED_OK worst_t %.3e worst_r %.3e map parts %s of %d allreduce %.1f us x %d" %
              (worst_t, worst_r, parts, n_full, 1e3 * ms / max(n, 1), n), flush=True)
        assert worst_t < 1e-4 and worst_r < 1e-5
        assert all(p < n_full for p in parts) and sum(parts) >= n_full  # each rank holds a strict subset
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
