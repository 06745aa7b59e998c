"""world_size-2 gloo tests of the N>1 host logic (no GPU): the batch-replay work split of bench.py
(ranks ray-cast disjoint worlds and exchange them; every rank then holds the same inputs; slots get
rank-distinct odometry) and the slab assignment of the sharded-map mode."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, load_package


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    bench.N_WORLDS = 4  # keep the CPU test small
    bench.PREFILL = 0   # (mature maps are built on the GPU; none here)
    worlds, maps = bench.prepare_worlds(None, 3, rank, world, 0)
    assert len(worlds) == 4 and all(len(w[2]) == 3 for w in worlds) and len(maps) == 4
    digest = np.array([float(np.sum([c.sum() + s.sum() for c, s in w[2]])) for w in worlds])
    gathered = [None] * world
    dist.all_gather_object(gathered, digest)
    assert all(np.array_equal(gathered[0], g) for g in gathered)  # every rank holds the same worlds
    odo = bench.slot_odometry(worlds, 6, rank)
    first = [None] * world
    dist.all_gather_object(first, odo[0][2].copy())
    if world > 1:
        assert not np.array_equal(first[0], first[1])  # rank-distinct drift => independent sequences per GPU
    pkg = load_package()
    slabs = [pkg.shard_slab(r, world) for r in range(world)]
    mine = pkg.shard_slab(rank, world)
    assert slabs[rank] == mine
    np.save(os.path.join(out_dir, "slab_%d.npy" % rank), np.array(mine))
    dist.barrier()
    dist.destroy_process_group()


def test_batch_replay_split_and_slabs_world2(built, tmp_path):
    mp.spawn(_worker, args=(2, 29641, str(tmp_path)), nprocs=2, join=True)
    s0, s1 = np.load(tmp_path / "slab_0.npy"), np.load(tmp_path / "slab_1.npy")
    assert np.isneginf(s0[0]) and np.isposinf(s1[1]) and s0[1] == s1[0]  # contiguous, no gap, no overlap


def test_slabs_partition_the_window(s2m, built):
    for world in (1, 2, 3, 4, 8):
        slabs = [s2m.shard_slab(r, world) for r in range(world)]
        assert np.isneginf(slabs[0][0]) and np.isposinf(slabs[-1][1])
        for a, b in zip(slabs[:-1], slabs[1:]):
            assert a[1] == b[0] and a[0] < a[1]
            assert (a[1] + 25.0) % 50.0 == 0.0  # slab faces are cube faces (laserMapping.cpp:313-315)


def test_rank_binding_is_harmless_without_a_gpu(built, monkeypatch):
    """bench.bind_near_gpu: one rank is never bound; several ranks bind to the cores NVML reports as nearest to their
    GPU and fall back to "none (...)" -- leaving the affinity untouched -- when NVML or the device is not there."""
    sys.path.insert(0, ROOT)
    import bench
    before = os.sched_getaffinity(0)
    assert bench.bind_near_gpu(0, 1) == "none"
    monkeypatch.setenv("S2M_NO_BIND", "1")
    assert bench.bind_near_gpu(0, 8) == "none"
    monkeypatch.delenv("S2M_NO_BIND")
    msg = bench.bind_near_gpu(0, 8)
    assert isinstance(msg, str) and (msg.startswith("none") or "cores near GPU" in msg)
    if msg.startswith("none"):
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
