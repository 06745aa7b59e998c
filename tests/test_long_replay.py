"""Long replay: 120 HDL-64 frames (a straight, a 90-degree turn and another straight of the closed-block
trajectory), GPU path and oracle each on its own map with optimisation ON.  Catches rare divergences
(threshold flips, a centroid rounded across a voxel boundary, window bookkeeping) that short tests miss."""
import numpy as np
import pytest

import harness
import oracle
from conftest import rot_angle

pytestmark = pytest.mark.gpu


def test_long_replay_matches_oracle(s2m, built):
    n = 120
    truth, odom, frames = harness.sequence(20261019, "HDL64", n)
    R = s2m.Registrar(0.4, 0.8)
    O = oracle.Oracle(0.4, 0.8)
    worst_t = worst_r = 0.0
    mismatched_counts = 0
    for f in range(n):
        rg, qg, tg = R.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
        ro, qo, to = O.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
        assert rg == ro
        worst_t = max(worst_t, float(np.linalg.norm(tg - to)))
        worst_r = max(worst_r, rot_angle(qg, qo))
        sg, so = R.stats, O.stats
        mismatched_counts += (list(sg.n_edge) != list(so.n_edge)) + (list(sg.n_plane) != list(so.n_plane))
        assert (sg.n_map_corner, sg.n_map_surf) == (so.n_map_corner, so.n_map_surf), f
    assert worst_t < 1e-4 and worst_r < 1e-5, (worst_t, worst_r)   # north_star tolerance
    assert worst_t < 1e-7 and worst_r < 1e-8, (worst_t, worst_r)   # what we actually see
    assert mismatched_counts == 0                                   # same correspondences accepted in every solve
    # the maps: same size, and (nearly) every point bit-identical
    for cls in (0, 1):
        a, b = R.map_download(cls), O.get_map(cls)
        assert a.shape == b.shape
        diff = int((a.view(np.uint32) != b.view(np.uint32)).any(1).sum())
        assert diff <= max(3, len(a) // 10000), (cls, diff, len(a))
    # drift check: the mapping result stays close to the truth while odometry has drifted away
    assert np.linalg.norm(tg - truth[n - 1, 4:]) < 0.5 * np.linalg.norm(odom[n - 1, 4:] - truth[n - 1, 4:])


def test_straight_drive_with_window_shift_matches_oracle(s2m, built):
    """450 m straight along the street in 6 m hops (far more than the registration can follow -- the point
    is the bookkeeping): the 21x21x11 window shifts under optimisation ON (laserMapping.cpp:324-508), cubes
    are recycled, new cubes turn valid with raw points pending.  GPU and oracle each keep their own map."""
    n = 75
    yaw = np.deg2rad(17.0)   # street axis of the synthetic world
    truth = np.zeros((n, 7))
    truth[:, 2], truth[:, 3] = np.sin(yaw / 2), np.cos(yaw / 2)
    d = np.arange(n) * 6.0
    truth[:, 4], truth[:, 5] = d * np.cos(yaw), d * np.sin(yaw)
    odom = harness.odometry(9, truth)
    R = s2m.Registrar(0.4, 0.8)
    O = oracle.Oracle(0.4, 0.8)
    worst_t = worst_r = 0.0
    shifted = False
    for f in range(n):
        c, s = harness.features("HDL64", harness.scan(9, "HDL64", truth[f], f))
        rg, qg, tg = R.register(c, s, odom[f, :4], odom[f, 4:])
        ro, qo, to = O.register(c, s, odom[f, :4], odom[f, 4:])
        assert rg == ro, f
        assert np.array_equal(R.window(), O.window()), f
        shifted = shifted or not np.array_equal(O.window(), [10, 10, 5])
        worst_t = max(worst_t, float(np.linalg.norm(tg - to)))
        worst_r = max(worst_r, rot_angle(qg, qo))
        sg, so = R.stats, O.stats
        assert (sg.n_map_corner, sg.n_map_surf) == (so.n_map_corner, so.n_map_surf), f
        assert list(sg.n_edge) == list(so.n_edge) and list(sg.n_plane) == list(so.n_plane), f
    assert shifted
    assert worst_t < 1e-4 and worst_r < 1e-5, (worst_t, worst_r)
