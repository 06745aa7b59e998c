#!/usr/bin/env python
"""bench_sharded.py -- BASELINE config 5: one registration stream against a spatially sharded map.

Launch with torchrun (1 rank per GPU). Every rank holds the x-slab of the map it owns (+1 m halo);
each LM evaluation allreduces one 32-double block per slot over NCCL/NVLink. Reports registrations/s
(CUDA events, max over ranks) and the allreduce latency -- the allreduce is latency-bound and its
cost is reported, not hidden. A secondary measurement; the driver's contract lives in bench.py.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def giant_map(n_side, spacing, seed):
    """A dense synthetic background map: jittered ground lattice over n_side*spacing metres (surf)
    and vertical pole lines (corner), laid around the street-grid world of harness/."""
    rng = np.random.default_rng(seed)
    g = (np.arange(n_side, dtype=np.float32) - n_side / 2) * spacing
    x, y = np.meshgrid(g, g, indexing="ij")
    x = x.ravel() + rng.uniform(-0.1, 0.1, x.size).astype(np.float32)
    y = y.ravel() + rng.uniform(-0.1, 0.1, y.size).astype(np.float32)
    z = (0.3 * np.sin(2 * np.pi * x / 40) * np.cos(2 * np.pi * y / 40)).astype(np.float32) - 3.0  # below the real ground
    surf = np.stack([x, y, z, np.zeros_like(x)], 1)
    px, py = np.meshgrid(g[::40], g[::40], indexing="ij")
    zz = np.arange(0, 6, 0.2, dtype=np.float32)
    corner = np.stack([np.repeat(px.ravel(), len(zz)) + 200.0, np.repeat(py.ravel(), len(zz)) + 200.0,
                       np.tile(zz, px.size), np.zeros(px.size * len(zz), np.float32)], 1).astype(np.float32)
    return corner, surf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--map-side", type=int, default=2000, help="ground lattice points per side")
    ap.add_argument("--map-spacing", type=float, default=0.5)
    ap.add_argument("--cap-map-surf", type=int, default=1 << 23, help="surf capacity of the window (points)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import harness
    from __graft_entry__ import load_package
    pkg = load_package()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    truth, odom, frames = harness.sequence(20261018, "HDL64", args.warmup + args.frames)
    corner, surf = giant_map(args.map_side, args.map_spacing, 5)
    n_bg = len(corner) + len(surf)
    R = pkg.Registrar(0.4, 0.8, device=local, cap_corner_in=max(1 << 14, len(corner)), cap_surf_in=max(1 << 17, len(surf)),
                      cap_map_corner=1 << 21, cap_map_surf=max(args.cap_map_surf, len(surf) + (1 << 20)), shard_rank=rank, shard_world=world)
    if world > 1:
        ids = [pkg.Registrar.shard_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        R.shard_init(ids[0])
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    R.set_stream(stream.cuda_stream)
    t0 = time.time()
    dropped = R.map_upload(corner, surf)  # every rank sees the whole background map and keeps its slab (+halo)
    n_up = n_bg
    up_s = time.time() - t0
    for f in range(args.warmup):
        R.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
    R.set_profiling(True)
    R.shard_profile(reset=True)
    R.phase_profile(reset=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = []
    for f in range(args.warmup, args.warmup + args.frames):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        R.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
        e1.record(stream)
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ar_ms, ar_n = R.shard_profile()
    phases = R.phase_profile()
    n_map = len(R.map_download(0)) + len(R.map_download(1))
    if rank == 0:
        print(json.dumps({"workload": "sharded_map_single_stream", "n_gpus": world, "frames": args.frames,
                          "registrations_per_s": args.frames / (ms * 1e-3), "ms_per_registration": ms / args.frames,
                          "allreduce_us_avg": 1e3 * ar_ms / max(ar_n, 1), "allreduces_per_registration": ar_n / args.frames,
                          "allreduce_ms_per_registration": ar_ms / args.frames,
                          "map_points_rank0": n_map, "background_points_uploaded": n_up, "not_stored_on_rank0": dropped, "upload_s": round(up_s, 2),
                          "phase_ms_per_registration": {k: round(v / args.frames, 4) for k, v in phases.items()}}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
