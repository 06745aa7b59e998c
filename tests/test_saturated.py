"""BASELINE config 3: Ouster OS1-64 scans (aloam_mulran.launch:9-12: minimum_range 0.5, 0.4 / 0.8 m)
registered against a SATURATED 21 x 21 x 11 cube window.  The window is filled the direct way SURVEY 8d allows:
every surface of the synthetic world inside it is sampled (harness.surfaces), uploaded raw, and every cube a
sensor can see is re-filtered by visiting the 3 x 3 block positions that cover the window without shifting it.
CUDA path (through the C ABI) against the oracle: saturated maps bit for bit, kNN of the first outer iteration
bit for bit inside the real flow, same accepted correspondences, poses within 1e-4 m / 1e-5 rad."""
import numpy as np
import pytest

import harness
import oracle
from conftest import rot_angle

pytestmark = pytest.mark.gpu

SEED = 20261018
TOL_T, TOL_R = 1e-4, 1e-5


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def saturate(ctx):
    e = np.zeros((0, 4), np.float32)
    for cx in (-250.0, 0.0, 250.0):
        for cy in (-250.0, 0.0, 250.0):
            ctx.register(e, e, [0, 0, 0, 1.0], [cx, cy, 0.0])


def test_os1_64_against_a_saturated_window(s2m, built):
    corner, surf = harness.surfaces(SEED)
    assert len(corner) > 500_000 and len(surf) > 2_000_000
    R = s2m.Registrar(0.4, 0.8, trace=True, cap_corner_in=1 << 17, cap_surf_in=1 << 18, cap_map_corner=1 << 21,
                      cap_map_surf=1 << 22)
    O = oracle.Oracle(0.4, 0.8, trace=True, use_kdtree=False)
    assert R.map_upload(corner, surf) == O.map_upload(corner, surf)
    saturate(R)
    saturate(O)
    assert np.array_equal(R.window(), O.window()) and list(R.window()) == [10, 10, 5]
    for cls in (0, 1):
        got, want = R.map_download(cls), O.get_map(cls)
        assert got.shape == want.shape and len(got) > 500_000
        assert np.array_equal(bits(got), bits(want))
    rng = np.random.default_rng(3)
    worst_t = worst_r = 0.0
    # poses spread over the window, jumping hundreds of metres from one call to the next (the valid block, the
    # pending lists and the cell index are rebuilt from scratch every time)
    for trial, (xs, k, quad) in enumerate([(137.0, 2, 1), (-301.5, -3, 0), (22.0, 0, 2), (288.0, -1, 3), (-95.0, 3, 1)]):
        pose = harness.street_pose(xs, k, quad)
        c, s = harness.features("OS1-64", harness.scan(SEED, "OS1-64", pose, trial))
        guess = pose.copy()
        guess[4:] += rng.uniform(-0.2, 0.2, 3)  # SURVEY 8d "T_init distribution"
        rg, qg, tg = R.register(c, s, guess[:4], guess[4:])
        ro, qo, to = O.register(c, s, guess[:4], guess[4:])
        assert rg == ro == 0, trial
        assert (R.stats.n_map_corner, R.stats.n_map_surf) == (O.stats.n_map_corner, O.stats.n_map_surf)
        assert R.stats.n_map_surf > 80_000  # a full ground level of cubes plus facades
        for cls in (0, 1):
            ig, dg, ug = R.trace_knn(0, cls)
            io, do, uo = O.trace_knn(0, cls)
            gate = do[:, 4] < 1.0
            assert len(ig) == len(io) and gate.sum() > 100
            assert np.array_equal(ig[gate], io[gate]) and np.array_equal(bits(dg[gate]), bits(do[gate]))
            assert np.array_equal(ug, uo)
        assert list(R.stats.n_edge) == list(O.stats.n_edge) and list(R.stats.n_plane) == list(O.stats.n_plane)
        worst_t = max(worst_t, float(np.linalg.norm(tg - to)))
        worst_r = max(worst_r, rot_angle(qg, qo))
        assert np.linalg.norm(to - pose[4:]) < 0.15  # and the registration pulled the guess back to the truth
    assert worst_t < TOL_T and worst_r < TOL_R, (worst_t, worst_r)
