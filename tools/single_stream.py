#!/usr/bin/env python
"""Latency of ONE sequence (the node's real-time case, BASELINE configs 1 / 2): s2m_register per frame from
host clouds, wall clock around each call, after the map has grown for `--prefill` frames; the CPU
restatement timed beside it."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import harness  # noqa: E402
import oracle  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sensor", default="HDL64")
ap.add_argument("--frames", type=int, default=80)
ap.add_argument("--prefill", type=int, default=40)
args = ap.parse_args()
pkg = load_package()
lr, pr = harness.LAUNCH[args.sensor]["line_res"], harness.LAUNCH[args.sensor]["plane_res"]
truth, odom, frames = harness.sequence(20261018, args.sensor, args.frames, step_m=1.0 if args.sensor == "HDL64" else 0.5)
R = pkg.Registrar(lr, pr)
O = oracle.Oracle(lr, pr)
tg, to = [], []
for f in range(args.frames):
    t0 = time.perf_counter()
    R.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
    t1 = time.perf_counter()
    O.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
    t2 = time.perf_counter()
    if f >= args.prefill:
        tg.append(t1 - t0)
        to.append(t2 - t1)
print(json.dumps({"workload": "single_stream_latency", "sensor": args.sensor, "frames_timed": len(tg),
                  "gpu_ms_per_registration": {"median": 1e3 * float(np.median(tg)), "p95": 1e3 * float(np.percentile(tg, 95))},
                  "cpu_oracle_ms_per_registration": {"median": 1e3 * float(np.median(to))},
                  "points": {"corner_in": len(frames[-1][0]), "surf_in": len(frames[-1][1]), "map_corner": int(R.stats.n_map_corner),
                             "map_surf": int(R.stats.n_map_surf)}}))
