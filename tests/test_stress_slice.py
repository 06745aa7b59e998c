"""A slice of tools/stress_parity.py inside the suite: a few random (seed, sensor, step length, odometry noise)
replays, CUDA path through the C ABI against the oracle.
  optimisation off: final maps bit-identical (rows B, C, V, I, W)
  optimisation on:  guard / counter agreement every frame, poses within 1e-4 m / 1e-5 rad
The full run (60 seeds x 12 frames plus the front end) stays a tool; its last result is quoted in DESIGN.md."""
import numpy as np
import pytest

import harness
import oracle
from conftest import rot_angle

pytestmark = pytest.mark.gpu

TOL_T, TOL_R = 1e-4, 1e-5


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("seed", [1003, 1004, 1005, 1010, 1017, 1021])
def test_random_replay(s2m, built, seed):
    rng = np.random.default_rng(seed)
    sensor = ["VLP16", "HDL64", "OS1-64"][seed % 3]
    step = float(rng.choice([0.3, 1.0, 2.5, 7.0]))
    lr, pr = harness.LAUNCH[sensor]["line_res"], harness.LAUNCH[sensor]["plane_res"]
    n = 7
    truth, odom, frames = harness.sequence(seed, sensor, n, step_m=step, sigma_t=float(rng.choice([0.02, 0.1])),
                                           sigma_r_deg=float(rng.choice([0.1, 0.5])))
    for skip in (True, False):
        R = s2m.Registrar(lr, pr, skip_optimization=skip)
        O = oracle.Oracle(lr, pr, skip_optimization=skip)
        for f in range(n):
            rg, qg, tg = R.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
            ro, qo, to = O.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
            sg, so = R.stats, O.stats
            assert rg == ro, (skip, f)
            assert (sg.n_map_corner, sg.n_map_surf, sg.n_corner_ds, sg.n_surf_ds) == \
                (so.n_map_corner, so.n_map_surf, so.n_corner_ds, so.n_surf_ds), (skip, f)
            if not skip:
                assert list(sg.n_edge) == list(so.n_edge) and list(sg.n_plane) == list(so.n_plane), f
                assert np.linalg.norm(tg - to) < TOL_T and rot_angle(qg, qo) < TOL_R, f
        if skip:
            for cls in (0, 1):
                a, b = R.map_download(cls), O.get_map(cls)
                assert a.shape == b.shape and np.array_equal(bits(a), bits(b)), cls
        R.close()
