#!/usr/bin/env python
"""bench_sharded.py -- BASELINE config 5: a spatially sharded giant map.

One context per GPU, every GPU holds the x-slab of cube columns it owns (+ 1 m halo) of EVERY slot's
map; a batch of B scans placed uniformly over the 21 x 21 cube window is registered per step, each
query answered by the rank whose slab holds it, and every LM evaluation allreduces one 32-double
block per slot over NCCL/NVLink (latency-bound: its cost is reported, not hidden).

Workload.  Slot b stands at street-grid offset (80 i_b, 80 j_b) m of the synthetic world (the grid
has period 80 m, the ground period 40 m, so the trajectory and the terrain repeat there while the
buildings differ), i.e. at world x, y spread over +-400 m of the window.  Its map is what the
reference would hold there after the prefill frames PLUS a dense scatter ("vegetation", SURVEY 8d
config 5) uniform in the 240 x 240 x 32 m box around the sensor (the height band the sweeps reach) -- inside the 5 x 5 x 3 valid block, so
every point of it is gathered, indexed and within reach of the kNN of the queries; total >= 50 M map
points over the slots.  The scatter makes the search heavy (tens to a hundred candidates per query);
it is map ballast for the throughput measurement, not a claim about mapping quality.

Used by bench.py (`sharded` object of its JSON line) and runnable on its own under torchrun.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRID_YAW = np.deg2rad(17.0)   # harness/synth.cpp kGridYaw
PITCH = 80.0                  # harness/synth.cpp kPitch


def slot_offsets(n_slots):
    """World (x, y) offsets of the slots: street-grid multiples of 80 m, x spread evenly over the window."""
    c, s = np.cos(GRID_YAW), np.sin(GRID_YAW)
    out = []
    for b in range(n_slots):
        i = int(round(-4 + 8.0 * b / max(n_slots - 1, 1)))
        j = (5 * b) % 9 - 4
        xs, ys = PITCH * i, PITCH * j
        out.append((c * xs - s * ys, s * xs + c * ys))
    return np.array(out)


def scatter(rng, centre, n, z_lo=-3.0, z_hi=29.0, half=120.0):
    p = np.empty((n, 4), np.float32)
    p[:, 0] = centre[0] + rng.uniform(-half, half, n)
    p[:, 1] = centre[1] + rng.uniform(-half, half, n)
    p[:, 2] = centre[2] + rng.uniform(z_lo, z_hi, n)
    p[:, 3] = 0.0
    return p


def run_sharded(pkg, torch, dist, rank, world, local_rank, slots=16, fill_corner=3_000_000, fill_surf=1_000_000,
                prefill=2, warmup=2, steps=5, check_slots=(0,), seed=20261018):
    import harness
    harness.set_threads(max(1, (os.cpu_count() or 1) // max(world, 1)))
    B = slots
    n_frames = prefill + warmup + steps
    offs = slot_offsets(B)
    truth0 = harness.trajectory(seed, n_frames, 1.0)
    seqs = []
    for b in range(B):
        tr = truth0.copy()
        tr[:, 4] += offs[b, 0]
        tr[:, 5] += offs[b, 1]
        od = harness.odometry(seed * 7919 + 31 * b, tr, 0.02, 0.1)
        frames = [harness.features("HDL64", harness.scan(seed, "HDL64", tr[f], f, 0.02)) for f in range(n_frames)]
        seqs.append((tr, od, frames))
    max_c = max(len(f[0]) for s in seqs for f in s[2]) + 64
    max_s = max(len(f[1]) for s in seqs for f in s[2]) + 64
    cap_c, cap_s = fill_corner + (1 << 19), fill_surf + (1 << 19)
    kw = dict(cap_corner_in=max(max_c, 1 << 18), cap_surf_in=max(max_s, 1 << 18), cap_map_corner=cap_c, cap_map_surf=cap_s)
    R = pkg.Registrar(0.4, 0.8, device=local_rank, batch=B, shard_rank=rank, shard_world=world, **kw)
    if world > 1:
        ids = [pkg.Registrar.shard_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        R.shard_init(ids[0])
    stream = torch.cuda.Stream()
    R.set_stream(stream.cuda_stream)
    # ---- the maps: every rank sees every point and keeps what touches its slab (+halo) ----
    t0 = time.perf_counter()
    uploaded = 0
    fills = {}
    for b in range(B):
        rng = np.random.default_rng(seed + 1000 + b)
        fc, fs = scatter(rng, seqs[b][0][0, 4:], fill_corner), scatter(rng, seqs[b][0][0, 4:], fill_surf)
        R.map_upload(fc, fs, slot=b)
        uploaded += len(fc) + len(fs)
        if b in check_slots:
            fills[b] = (fc, fs)
    upload_s = time.perf_counter() - t0

    def step(f):
        cs = [seqs[b][2][f][0] for b in range(B)]
        ss = [seqs[b][2][f][1] for b in range(B)]
        co = np.cumsum([0] + [len(c) for c in cs]).astype(np.int32)
        so = np.cumsum([0] + [len(c) for c in ss]).astype(np.int32)
        q = np.array([seqs[b][1][f, :4] for b in range(B)])
        t = np.array([seqs[b][1][f, 4:] for b in range(B)])
        return R.register_batch(np.concatenate(cs), co, np.concatenate(ss), so, q, t)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for f in range(prefill + warmup):
        step(f)
    R.set_profiling(True)
    R.shard_profile(reset=True)
    R.phase_profile(reset=True)
    barrier()
    poses = []
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for f in range(prefill + warmup, n_frames):
            st, q, t = step(f)
            poses.append(np.c_[q, t])
        e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    ar_ms, ar_n = R.shard_profile()
    phases = R.phase_profile()
    stats = [R.batch_stats[b] for b in range(B)]
    n_local = sum(st.n_map_corner + st.n_map_surf for st in stats)  # points this rank holds in the valid blocks (= searched)
    poses = np.array(poses)
    ph_names = ["input", "voxel", "index", "associate", "solve", "update", "readback"]
    mine = torch.tensor(np.r_[ms, phases["associate"], float(n_local), [phases[k] for k in ph_names]], dtype=torch.float64, device="cuda")
    pose_t = torch.tensor(poses.view(np.int64).ravel().copy(), device="cuda")
    if world > 1:
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        ref = pose_t.clone()
        dist.broadcast(ref, src=0)
        same = torch.tensor([int(torch.equal(ref, pose_t))], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        identical = bool(same.item())
    else:
        allv, identical = [mine], True
    allv = np.array([v.cpu().numpy() for v in allv])
    # ---- agreement with an unsharded context (one GPU, whole map) on a few slots ----
    agree = None
    if rank == 0 and check_slots:
        worst_t = worst_r = 0.0
        for b in check_slots:
            U = pkg.Registrar(0.4, 0.8, device=local_rank, **kw)
            U.map_upload(*fills[b])
            tr, od, frames = seqs[b]
            for f in range(n_frames):
                rc, q, t = U.register(frames[f][0], frames[f][1], od[f, :4], od[f, 4:])
                if f >= prefill + warmup:
                    ps = poses[f - prefill - warmup, b]
                    worst_t = max(worst_t, float(np.linalg.norm(t - ps[4:])))
                    d = abs(float(np.dot(q, ps[:4])))
                    worst_r = max(worst_r, 2.0 * float(np.arccos(min(1.0, d))))
            U.close()
        agree = {"slots_checked": list(check_slots), "max_translation_diff_m": worst_t, "max_rotation_diff_rad": worst_r}
    R.close()
    torch.cuda.empty_cache()
    ms_max = float(allv[:, 0].max())
    assoc = allv[:, 1] / steps
    res = {
        "workload": "sharded_giant_map: %d scans per step placed over the window, x-slabs of cube columns per GPU (+1 m halo), "
                    "allreduce of 32 doubles per slot per LM evaluation" % B,
        "n_gpus": world, "slots": B, "steps": steps, "warmup": warmup, "prefill_frames": prefill,
        "registrations_per_s": B * steps / (ms_max * 1e-3), "ms_per_step": ms_max / steps,
        "map_points_uploaded": int(uploaded), "map_points_in_valid_blocks_all_ranks": int(allv[:, 2].sum()),
        "map_points_in_valid_blocks_per_rank": [int(v) for v in allv[:, 2]],
        "allreduce_us_avg": 1e3 * ar_ms / max(ar_n, 1), "allreduces_per_step": ar_n / steps,
        "allreduce_ms_per_step": ar_ms / steps,
        "association_ms_per_step_per_rank": [round(float(v), 4) for v in assoc],
        "query_imbalance_max_over_mean": float(assoc.max() / max(assoc.mean(), 1e-12)),
        "all_ranks_bit_identical": identical, "agreement_with_one_gpu": agree,
        "phase_ms_per_step_per_rank": {k: [round(float(allv[r, 3 + i]) / steps, 3) for r in range(len(allv))] for i, k in enumerate(ph_names)},
        "phase_note": "a rank's index / solve phases contain its allreduces, i.e. the time it waits for the slowest rank",
        "upload_s": round(upload_s, 2),
        "slot_x_offsets_m": [round(float(v), 1) for v in offs[:, 0]],
    }
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--slots", type=int, default=16)
    ap.add_argument("--fill-corner", type=int, default=3_000_000)
    ap.add_argument("--fill-surf", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from __graft_entry__ import load_package
    pkg = load_package()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run_sharded(pkg, torch, dist, rank, world, local, slots=args.slots, fill_corner=args.fill_corner,
                      fill_surf=args.fill_surf, warmup=args.warmup, steps=args.steps)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
