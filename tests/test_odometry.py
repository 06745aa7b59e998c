"""scan-to-scan odometry (laserOdometry.cpp:220-591, lidarFactor.hpp:57-104).

CPU part: the restatement (oracle.Odometer) tracks a synthetic trajectory.  GPU part: s2m_odom_* against
it -- identical correspondences (closest / second / third index of every query, both passes), identical
correspondence counts, relative and integrated poses within 1e-4 m / 1e-5 rad (observed ~1e-12) --
and the whole front end on the device: raw sweep -> features -> odometry -> mapping."""
import os

import numpy as np
import pytest

import harness
import oracle
from conftest import rot_angle

TOL_T, TOL_R = 1e-4, 1e-5


def features_stream(sensor, seed, n, step):
    truth = harness.trajectory(seed, n, step)
    mr = harness.LAUNCH[sensor]["minimum_range"]
    out = []
    for f in range(n):
        out.append(oracle.scan_registration(sensor, harness.scan(seed, sensor, truth[f], f), mr))
    return truth, out


def rel_truth(truth, f):
    from scipy.spatial.transform import Rotation as Rot
    R0 = Rot.from_quat(truth[0, :4])
    return (R0.inv() * Rot.from_quat(truth[f, :4])).as_quat(), R0.inv().apply(truth[f, 4:] - truth[0, 4:])


def test_oracle_odometry_tracks_the_trajectory(built):
    truth, feats = features_stream("VLP16", 20261018, 6, 0.5)
    O = oracle.Odometer()
    for f, A in enumerate(feats):
        q, t = O.step(A["sharp"], A["flat"], A["less_sharp"], A["less_flat"])
        if f == 0:
            assert np.array_equal(q, [0, 0, 0, 1]) and np.array_equal(t, [0, 0, 0]) and O.counts.sum() == 0   # :267-271
            continue
        assert O.counts[0] > 100 and O.counts[2] > 200
        qt, tt = rel_truth(truth, f)
        assert np.linalg.norm(t - tt) < 0.05 and rot_angle(q, qt) < np.deg2rad(0.5), f
        assert abs(np.linalg.norm(O.para[4:]) - 0.5) < 0.05       # the relative motion of one 0.5 m step


KAIST = "/root/reference/utils/sample_data/KAIST03/"


@pytest.mark.skipif(not os.path.exists(KAIST), reason="reference tree not mounted (authoring container only)")
def test_restated_front_end_on_the_reference_real_keyframes(built, s2m):
    """the 21 real OS1-64 keyframes the reference ships (about 1.2 m apart, i.e. far
    coarser than the 10 Hz sweeps the odometry is meant for) through the three restatements -- features,
    odometry, mapping -- land within decimetres of the poses the reference saved over the 23.6 m stretch."""
    poses = np.loadtxt(KAIST + "optimized_poses.txt")[:21].reshape(21, 3, 4)
    Oo, Om = oracle.Odometer(), oracle.Oracle(0.4, 0.8)      # aloam_mulran.launch:11-12
    worst = 0.0
    for k in range(21):
        xyz = s2m.pcd_read(KAIST + "Scans/%06d.pcd" % k)[:, :3]
        A = oracle.scan_registration("OS1-64", xyz, 0.5)
        qo, to = Oo.step(A["sharp"], A["flat"], A["less_sharp"], A["less_flat"])
        rc, qm, tm = Om.register(A["less_sharp"], A["less_flat"], qo, to)
        err = float(np.linalg.norm(tm - poses[k, :, 3]))
        worst = max(worst, err)
        if 1 <= k <= 10:
            assert err < 0.10, (k, err)
            assert err < np.linalg.norm(to - poses[k, :, 3])      # the mapping corrects the (lagging) odometry
    assert worst < 0.30 and np.linalg.norm(poses[20, :, 3]) > 23.0


@pytest.mark.gpu
@pytest.mark.parametrize("sensor,step", [("VLP16", 0.5), ("HDL64", 1.0)])
def test_cuda_odometry_matches_oracle(s2m, built, sensor, step):
    truth, feats = features_stream(sensor, 7, 5, step)
    O = oracle.Odometer()
    R = s2m.Odometer(trace=True, cap_less_flat=1 << 16)
    for f, A in enumerate(feats):
        qo, to = O.step(A["sharp"], A["flat"], A["less_sharp"], A["less_flat"])
        qg, tg = R.step(A["sharp"], A["flat"], A["less_sharp"], A["less_flat"])
        assert list(R.counts[0]) == list(O.counts), f
        if f > 0:
            for outer in range(2):
                e, p = O.trace(outer)
                ge, used_e = R.trace(outer, 0)
                gp, used_p = R.trace(outer, 1)
                assert np.array_equal(ge[:, :2], e), (f, outer)
                assert np.array_equal(gp, p), (f, outer)
                assert np.array_equal(used_e, e[:, 1] >= 0) and np.array_equal(used_p, (p[:, 1] >= 0) & (p[:, 2] >= 0))
        assert np.linalg.norm(R.para[0, 4:] - O.para[4:]) < TOL_T and rot_angle(R.para[0, :4], O.para[:4]) < TOL_R, f
        assert np.linalg.norm(tg - to) < TOL_T and rot_angle(qg, qo) < TOL_R, f
    assert R.launch_count() > 0


@pytest.mark.gpu
def test_cuda_front_end_on_two_real_keyframes(s2m, built):
    """two consecutive real OS1-64 keyframes of the reference (tests/golden/kaist03_scan10.npz): features and
    odometry on the device against the restatements -- real ring populations (> 3000 points in a bucket),
    real geometry."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kaist03_scan10.npz"))
    F = s2m.FeatureExtractor("OS1-64", 0.5, batch=1, cap_points=40000)
    D = s2m.Odometer(trace=True, cap_less_flat=1 << 16)
    O = oracle.Odometer()
    names = ("sharp", "flat", "less_sharp", "less_flat")
    for xyz in (g["xyz"], g["xyz_next"]):
        F.extract(xyz, np.array([0, len(xyz)], np.int32))
        args = []
        for k in names:
            args += [F.device_cloud(k), F.offsets(k)]
        qd, td = D.step_batch(*args, device_ptrs=True)
        A = oracle.scan_registration("OS1-64", xyz, 0.5)
        qo, to = O.step(*[A[k] for k in names])
        assert list(D.counts[0]) == list(O.counts)
    assert O.counts[0] > 100 and O.counts[2] > 50
    for outer in range(2):
        e, p = O.trace(outer)
        assert np.array_equal(D.trace(outer, 0)[0][:, :2], e) and np.array_equal(D.trace(outer, 1)[0], p)
    assert np.linalg.norm(td[0] - to) < TOL_T and rot_angle(qd[0], qo) < TOL_R
    assert 0.3 < np.linalg.norm(to) < 2.0      # the keyframes are about 1 m apart


@pytest.mark.gpu
def test_cuda_odometry_degenerate_inputs(s2m, built):
    """empty and tiny clouds, queries far from everything, ring numbers outside 0..255"""
    rng = np.random.default_rng(4)
    e = np.zeros((0, 4), np.float32)

    def cloud(n, rings, spread=5.0):
        p = np.c_[rng.uniform(-spread, spread, (n, 3)), np.sort(rng.integers(0, rings, n)) + rng.uniform(0, 0.09, n)]
        return p.astype(np.float32)

    D, O = s2m.Odometer(trace=True), oracle.Odometer()
    seq = [(e, e, e, e),                                              # nothing at all
           (cloud(5, 3), cloud(9, 3), cloud(20, 4), cloud(31, 4)),     # fewer previous points than one chunk
           (e, e, cloud(40, 4), cloud(70, 4)),                         # no queries
           (cloud(30, 4), cloud(50, 4), e, cloud(33, 4)),              # empty corner target next time
           (cloud(30, 4), cloud(50, 4), cloud(64, 4), cloud(96, 4)),
           (cloud(20, 4) + np.float32([300, 0, 0, 0]), cloud(20, 4) + np.float32([0, 200, 0, 0]), cloud(64, 4), cloud(96, 4)),  # > 5 m from everything
           (cloud(30, 4), cloud(50, 4), cloud(64, 4), cloud(96, 4))]
    for i, (a, b, c, dd) in enumerate(seq):
        qg, tg = D.step(a, b, c, dd)
        qo, to = O.step(a, b, c, dd)
        assert list(D.counts[0]) == list(O.counts), i
        if i > 0 and len(a) + len(b) > 0:
            for outer in range(2):
                eo, po = O.trace(outer)
                assert np.array_equal(D.trace(outer, 0)[0][:, :2], eo) and np.array_equal(D.trace(outer, 1)[0], po), (i, outer)
        assert np.linalg.norm(tg - to) < 1e-4 and np.abs(qg - qo).max() < 1e-5, i    # tiny random clouds: poses may be wild, but equal
    bad = cloud(40, 4)
    bad[:, 3] += 300.0
    with pytest.raises(s2m.S2MError):
        D.step(cloud(5, 3), cloud(5, 3), bad, cloud(40, 4))


@pytest.mark.gpu
def test_cuda_odometry_batch_slots_are_independent(s2m, built):
    """two sequences in one context (the second lags one sweep) == two single contexts, bit for bit"""
    _, feats = features_stream("VLP16", 3, 5, 0.5)
    B2 = s2m.Odometer(batch=2)
    S = [s2m.Odometer(), s2m.Odometer()]
    names = ("sharp", "flat", "less_sharp", "less_flat")
    for f in range(1, 5):
        fr = [feats[f], feats[f - 1]]
        packed, offs = [], []
        for k in names:
            packed.append(np.concatenate([x[k] for x in fr]))
            offs.append(np.cumsum([0] + [len(x[k]) for x in fr]).astype(np.int32))
        q, t = B2.step_batch(packed[0], offs[0], packed[1], offs[1], packed[2], offs[2], packed[3], offs[3])
        for b in range(2):
            q1, t1 = S[b].step(*[fr[b][k] for k in names])
            assert np.array_equal(q[b], q1) and np.array_equal(t[b], t1), (f, b)


@pytest.mark.gpu
def test_whole_front_end_on_the_device(s2m, built):
    """raw sweep -> s2m_fx_extract -> s2m_odom_step_batch -> s2m_register_batch_dev, device pointers all the
    way (one H2D copy of the raw sweep per frame), against the CPU chain of the three restatements."""
    seed, n = 20261018, 6
    truth = harness.trajectory(seed, n, 0.5)
    F = s2m.FeatureExtractor("VLP16", 0.1, batch=1)
    D = s2m.Odometer()
    M = s2m.Registrar(0.2, 0.4)
    Oo, Om = oracle.Odometer(), oracle.Oracle(0.2, 0.4)
    for f in range(n):
        xyz = harness.scan(seed, "VLP16", truth[f], f)
        F.extract(xyz, np.array([0, len(xyz)], np.int32))
        dev = {k: F.device_cloud(k) for k in ("sharp", "flat", "less_sharp", "less_flat")}
        off = {k: F.offsets(k) for k in dev}
        q_od, t_od = D.step_batch(dev["sharp"], off["sharp"], dev["flat"], off["flat"], dev["less_sharp"], off["less_sharp"],
                                  dev["less_flat"], off["less_flat"], device_ptrs=True)
        st, q_w, t_w = M.register_batch_ptr(dev["less_sharp"], off["less_sharp"], dev["less_flat"], off["less_flat"],
                                            q_od[0], t_od[0], True)
        A = oracle.scan_registration("VLP16", xyz, 0.1)
        qo, to = Oo.step(A["sharp"], A["flat"], A["less_sharp"], A["less_flat"])
        rc, qm, tm = Om.register(A["less_sharp"], A["less_flat"], qo, to)
        assert st[0] == rc
        assert np.linalg.norm(t_od[0] - to) < TOL_T and rot_angle(q_od[0], qo) < TOL_R, f
        assert np.linalg.norm(t_w[0] - tm) < TOL_T and rot_angle(q_w[0], qm) < TOL_R, f
    qt, tt = rel_truth(truth, n - 1)
    # the odometry follows the true trajectory (the VLP-16 mapping on this young, street-aligned map lags
    # along the street -- an algorithm property both chains share, see test_gpu_parity)
    assert np.linalg.norm(t_od[0] - tt) < 0.05


@pytest.mark.gpu
def test_cuda_chain_on_all_21_real_keyframes_of_the_reference(s2m, built):
    """SURVEY 8c item 5 on the DEVICE: the 21 real OS1-64 keyframes the reference ships (committed fixture
    tests/golden/kaist03_keyframes.npz, generator tests/golden/make_kaist03.py --keyframes) through
    s2m_fx_extract -> s2m_odom_step_batch -> s2m_register_batch_dev with device pointers all the way.  Anchors:
    (a) the poses the reference itself saved for these scans (optimized_poses.txt rows 0..20): within 10 cm for
    keyframes 1..10, 30 cm over the 23.6 m stretch -- the same bars the CPU restatements meet in the authoring
    container; (b) the CPU chain of the three restatements on the same scans: odometry and mapping poses within
    1e-4 m / 1e-5 rad, every feature cloud bit-identical."""
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kaist03_keyframes.npz"))
    xyz_all, off, ref = g["xyz"], g["off"], g["ref_poses"]
    assert len(off) == 22 and ref.shape == (21, 3, 4)
    F = s2m.FeatureExtractor("OS1-64", 0.5, batch=1, cap_points=40000)   # aloam_mulran.launch:9
    D = s2m.Odometer(cap_less_flat=1 << 16)
    M = s2m.Registrar(0.4, 0.8)                                          # aloam_mulran.launch:11-12
    Oo, Om = oracle.Odometer(), oracle.Oracle(0.4, 0.8)
    names = ("sharp", "flat", "less_sharp", "less_flat")
    worst = 0.0
    for k in range(21):
        xyz = np.ascontiguousarray(xyz_all[off[k]:off[k + 1]])
        F.extract(xyz, np.array([0, len(xyz)], np.int32))
        dev = {n: F.device_cloud(n) for n in names}
        o = {n: F.offsets(n) for n in names}
        q_od, t_od = D.step_batch(dev["sharp"], o["sharp"], dev["flat"], o["flat"], dev["less_sharp"], o["less_sharp"],
                                  dev["less_flat"], o["less_flat"], device_ptrs=True)
        st, q_w, t_w = M.register_batch_ptr(dev["less_sharp"], o["less_sharp"], dev["less_flat"], o["less_flat"],
                                            q_od[0], t_od[0], True)
        A = oracle.scan_registration("OS1-64", xyz, 0.5)
        for n in names:
            assert np.array_equal(F.cloud(n)[0].view(np.uint32), A[n].view(np.uint32)), (k, n)
        qo, to = Oo.step(*[A[n] for n in names])
        rc, qm, tm = Om.register(A["less_sharp"], A["less_flat"], qo, to)
        assert st[0] == rc, k
        assert np.linalg.norm(t_od[0] - to) < TOL_T and rot_angle(q_od[0], qo) < TOL_R, k
        assert np.linalg.norm(t_w[0] - tm) < TOL_T and rot_angle(q_w[0], qm) < TOL_R, k
        err = float(np.linalg.norm(t_w[0] - ref[k, :, 3]))
        worst = max(worst, err)
        if 1 <= k <= 10:
            assert err < 0.10, (k, err)
            assert err < np.linalg.norm(t_od[0] - ref[k, :, 3])      # the mapping corrects the (lagging) odometry
    assert worst < 0.30 and np.linalg.norm(ref[20, :, 3]) > 23.0
