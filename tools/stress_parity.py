#!/usr/bin/env python
"""One-off stress run (not part of the test suite): many seeds x sensors x step lengths, GPU path vs oracle.
  mode A (skip_optimization): map evolution must be bit-identical (rows B, C, V, I, W)
  mode B (optimisation on):   counters identical, poses within 1e-4 m / 1e-5 rad
Prints one line per failure and a summary; exit code 1 on any failure."""
import argparse
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import harness  # noqa: E402
import oracle  # noqa: E402
from __graft_entry__ import load_package  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=12)
    ap.add_argument("--frames", type=int, default=10)
    args = ap.parse_args()
    pkg = load_package()
    fails = runs = 0
    worst = 0.0
    for seed in range(1000, 1000 + args.seeds):
        rng = np.random.default_rng(seed)
        sensor = ["VLP16", "HDL64", "OS1-64"][seed % 3]
        step = float(rng.choice([0.3, 1.0, 2.5, 7.0]))
        lr, pr = harness.LAUNCH[sensor]["line_res"], harness.LAUNCH[sensor]["plane_res"]
        truth, odom, frames = harness.sequence(seed, sensor, args.frames, step_m=step, sigma_t=float(rng.choice([0.02, 0.1])),
                                               sigma_r_deg=float(rng.choice([0.1, 0.5])))
        for skip in (True, False):
            runs += 1
            R = pkg.Registrar(lr, pr, skip_optimization=skip)
            O = oracle.Oracle(lr, pr, skip_optimization=skip)
            ok = True
            for f in range(args.frames):
                rg, qg, tg = R.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
                ro, qo, to = O.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
                sg, so = R.stats, O.stats
                same = rg == ro and (sg.n_map_corner, sg.n_map_surf, sg.n_corner_ds, sg.n_surf_ds) == \
                    (so.n_map_corner, so.n_map_surf, so.n_corner_ds, so.n_surf_ds)
                if not skip:
                    same = same and list(sg.n_edge) == list(so.n_edge) and list(sg.n_plane) == list(so.n_plane)
                    dt = float(np.linalg.norm(tg - to))
                    dq = float(np.abs(qg - qo).max())
                    worst = max(worst, dt)
                    same = same and dt < 1e-4 and dq < 1e-5
                if not same:
                    print("FAIL seed %d %s step %.1f skip %s frame %d" % (seed, sensor, step, skip, f), flush=True)
                    ok = False
                    break
            if ok and skip:
                for cls in (0, 1):
                    a, b = R.map_download(cls), O.get_map(cls)
                    if a.shape != b.shape or not np.array_equal(bits(a), bits(b)):
                        print("FAIL seed %d %s step %.1f: final map differs (cls %d)" % (seed, sensor, step, cls), flush=True)
                        ok = False
            fails += not ok
            R.close()
    # front end: feature extraction bit-identical, odometry correspondences + poses
    for seed in range(2000, 2000 + args.seeds):
        rng = np.random.default_rng(seed)
        sensor = ["VLP16", "HDL64", "OS1-64"][seed % 3]
        step = float(rng.choice([0.3, 1.0, 2.0]))
        mr = harness.LAUNCH[sensor]["minimum_range"]
        n = 5
        truth = harness.trajectory(seed, n, step)
        F = pkg.FeatureExtractor(sensor, mr, batch=1)
        D = pkg.Odometer(trace=True)
        Oo = oracle.Odometer()
        runs += 1
        ok = True
        for f in range(n):
            xyz = harness.scan(seed, sensor, truth[f], f, range_sigma=float(rng.choice([0.0, 0.02, 0.05])))
            F.extract(xyz, np.array([0, len(xyz)], np.int32))
            A = oracle.scan_registration(sensor, xyz, mr)
            for k in A:
                got, _ = F.cloud(k)
                if got.shape != A[k].shape or not np.array_equal(bits(got), bits(A[k])):
                    print("FAIL fx seed %d %s frame %d cloud %s" % (seed, sensor, f, k), flush=True)
                    ok = False
            names = ("sharp", "flat", "less_sharp", "less_flat")
            qd, td = D.step(*[A[k] for k in names])
            qo, to = Oo.step(*[A[k] for k in names])
            good = list(D.counts[0]) == list(Oo.counts) and np.linalg.norm(td - to) < 1e-4 and np.abs(qd - qo).max() < 1e-5
            if good and f > 0:
                for outer in range(2):
                    e, p = Oo.trace(outer)
                    good = good and np.array_equal(D.trace(outer, 0)[0][:, :2], e) and np.array_equal(D.trace(outer, 1)[0], p)
            if not good:
                print("FAIL odom seed %d %s step %.1f frame %d" % (seed, sensor, step, f), flush=True)
                ok = False
            if not ok:
                break
        fails += not ok
        F.close()
        D.close()
    print("stress: %d runs, %d failures, worst pose difference %.2e m" % (runs, fails, worst))
    sys.exit(1 if fails else 0)


if __name__ == "__main__":
    main()
