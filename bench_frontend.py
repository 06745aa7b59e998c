#!/usr/bin/env python
"""bench_frontend.py -- the whole front end on the device (SURVEY 8f rows N2 + N3 + the mapping path).

A secondary measurement (the driver's contract lives in bench.py). `--batch` synthetic HDL-64 sequences
advance together; per step and sequence: raw sweep (resident in HBM, or copied from pinned host memory
with --host) -> s2m_fx_extract -> s2m_odom_step_batch -> s2m_register_batch_dev, the clouds handed on as
device pointers. Reports frames/s of every stage and of the chain (host-timed: each call ends with a
stream synchronisation), with the CPU restatements timed beside them on one core.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=14)
    ap.add_argument("--warmup", type=int, default=4, help="leading frames left out of the timing (first sweep only initialises)")
    ap.add_argument("--worlds", type=int, default=8)
    ap.add_argument("--host", action="store_true", help="raw sweeps start in pinned host memory (H2D inside the timed region)")
    ap.add_argument("--cpu-frames", type=int, default=6)
    args = ap.parse_args()
    import torch
    import harness
    import oracle
    from __graft_entry__ import load_package
    pkg = load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench_frontend.py: no CUDA device; the product has no CPU path")
    sensor, mr = "HDL64", harness.LAUNCH["HDL64"]["minimum_range"]
    B, n = args.batch, args.frames
    worlds = []
    for w in range(args.worlds):
        tr = harness.trajectory(100 + w, n)
        worlds.append([harness.scan(100 + w, sensor, tr[f], f) for f in range(n)])
    steps = []
    for f in range(n):
        sw = [worlds[b % args.worlds][f] for b in range(B)]
        off = np.cumsum([0] + [len(x) for x in sw]).astype(np.int32)
        t = torch.from_numpy(np.concatenate(sw))
        steps.append((t.pin_memory() if args.host else t.cuda(), off))
    cap = max(int(np.diff(o).max()) for _, o in steps) + 64
    F = pkg.FeatureExtractor(sensor, mr, batch=B, cap_points=cap)
    D = pkg.Odometer(batch=B, cap_sharp=1024, cap_flat=2048, cap_less_sharp=8192, cap_less_flat=40960)
    M = pkg.Registrar(0.4, 0.8, batch=B, cap_corner_in=8192, cap_surf_in=40960, cap_map_corner=1 << 18, cap_map_surf=1 << 18)
    names = ("sharp", "flat", "less_sharp", "less_flat")
    t_fx = t_od = t_map = 0.0
    counts = None
    for f in range(n):
        src, off = steps[f]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if args.host:
            F.extract(src.numpy(), off)
        else:
            F.extract(src.data_ptr(), off, device=True)
        t1 = time.perf_counter()
        dev = {k: F.device_cloud(k) for k in names}
        offs = {k: F.offsets(k) for k in names}
        q_od, t_odo = D.step_batch(dev["sharp"], offs["sharp"], dev["flat"], offs["flat"], dev["less_sharp"], offs["less_sharp"],
                                   dev["less_flat"], offs["less_flat"], device_ptrs=True)
        t2 = time.perf_counter()
        M.register_batch_ptr(dev["less_sharp"], offs["less_sharp"], dev["less_flat"], offs["less_flat"], q_od, t_odo, True)
        t3 = time.perf_counter()
        if f >= args.warmup:
            t_fx += t1 - t0; t_od += t2 - t1; t_map += t3 - t2
            counts = D.counts.mean(0)
    k = (n - args.warmup) * B
    # CPU restatements, one sequence, one core
    Oo, Om = oracle.Odometer(), oracle.Oracle(0.4, 0.8)
    c_fx = c_od = c_map = 0.0
    m = min(args.cpu_frames + 1, n)
    for f in range(m):
        t0 = time.perf_counter()
        A = oracle.scan_registration(sensor, worlds[0][f], mr)
        t1 = time.perf_counter()
        qo, to = Oo.step(A["sharp"], A["flat"], A["less_sharp"], A["less_flat"])
        t2 = time.perf_counter()
        Om.register(A["less_sharp"], A["less_flat"], qo, to)
        t3 = time.perf_counter()
        if f >= 1:
            c_fx += t1 - t0; c_od += t2 - t1; c_map += t3 - t2
    cm = m - 1
    print(json.dumps({
        "workload": "hdl64_front_end_chain", "batch": B, "timed_steps": n - args.warmup, "raw_sweeps": "pinned host" if args.host else "resident in HBM",
        "frames_per_s": {"features": k / t_fx, "odometry": k / t_od, "mapping": k / t_map, "chain": k / (t_fx + t_od + t_map)},
        "ms_per_step": {"features": 1e3 * t_fx / (n - args.warmup), "odometry": 1e3 * t_od / (n - args.warmup),
                        "mapping": 1e3 * t_map / (n - args.warmup)},
        "odometry_correspondences_mean": {"corner": float(counts[1]), "plane": float(counts[3])},
        "cpu_baseline": {"unit": "frames/s", "cores": 1, "kind": "port", "features": cm / c_fx, "odometry": cm / c_od, "mapping": cm / c_map,
                         "chain": cm / (c_fx + c_od + c_map),
                         "note": "oracle/ restatements; the odometry one searches its nearest neighbour by brute force (the reference uses a KD-tree)"}}), flush=True)


if __name__ == "__main__":
    main()
