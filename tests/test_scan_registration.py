"""feature extraction (scanRegistration.cpp:116-454).

CPU part: the restatement oracle/scan_registration.cpp against the ring numbers the REFERENCE
itself stored in the KAIST03 scans it ships (authoring container only), the two tie rules of the
sector sort, and basic structure.  GPU part: s2m_fx_* bit-identical to the oracle on all five
clouds, for batches of sweeps of three sensors, and handed on the device to the mapping call."""
import os

import numpy as np
import pytest

import harness
import oracle

CLOUDS = ("full", "sharp", "less_sharp", "flat", "less_flat")
KAIST = "/root/reference/utils/sample_data/KAIST03/Scans/"


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def sweeps(sensor, seed, n):
    tr = harness.trajectory(seed, n)
    return [harness.scan(seed, sensor, tr[f], f) for f in range(n)]


def test_oracle_structure_and_tie_rules(built):
    for sensor in ("HDL64", "VLP16", "OS1-64"):
        mr = harness.LAUNCH[sensor]["minimum_range"]
        for xyz in sweeps(sensor, 11, 2):
            A = oracle.scan_registration(sensor, xyz, mr, tie_rule=0)
            S = oracle.scan_registration(sensor, xyz, mr, tie_rule=1)   # std::sort with the reference's comparator
            for k in CLOUDS:
                assert np.array_equal(bits(A[k]), bits(S[k])), (sensor, k)
            ring = np.rint(A["full"][:, 3]).astype(int)
            assert (np.diff(ring) >= 0).all()                            # ring-major (:261-267)
            frac = A["full"][:, 3] - ring
            assert frac.min() > -0.02 and frac.max() < 0.15            # 0.1 * relTime (:251-252); the synthetic sweep overshoots a turn slightly
            nr = len(np.unique(ring))
            assert len(A["sharp"]) <= 2 * 6 * nr and len(A["less_sharp"]) <= 20 * 6 * nr and len(A["flat"]) <= 4 * 6 * nr
            # sharp is a subset of less_sharp, in order
            ls = {tuple(r) for r in bits(A["less_sharp"]).tolist()}
            assert all(tuple(r) in ls for r in bits(A["sharp"]).tolist())
            assert len(A["less_flat"]) > 10 * len(A["flat"])
    # degenerate inputs
    e = oracle.scan_registration("VLP16", np.zeros((0, 3), np.float32), 0.1)
    assert all(len(e[k]) == 0 for k in CLOUDS)
    bad = np.full((100, 3), np.nan, np.float32)
    assert len(oracle.scan_registration("VLP16", bad, 0.1)["full"]) == 0


@pytest.mark.skipif(not os.path.exists(KAIST), reason="reference tree not mounted (authoring container only)")
def test_oracle_ring_rule_matches_rings_stored_by_the_reference(built, s2m):
    """The KAIST03 scans carry ring + 0.1*relTime in their intensity, written by the reference's
    scanRegistration: the restated OS1-64 ring rule reproduces the ring of all 765 919 points."""
    total = 0
    for k in range(21):
        p = s2m.pcd_read(KAIST + "%06d.pcd" % k)
        r = oracle.ring_of("OS1-64", p[:, :3])
        d = p[:, 3] - r
        assert (r >= 0).all() and d.min() > -1e-3 and d.max() < 0.101, k
        total += len(p)
    assert total == 765919


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kaist03_scan10.npz")


def test_oracle_on_a_real_scan_of_the_reference(built):
    """tests/golden/kaist03_scan10.npz: a real OS1-64 sweep the reference shipped, with the ring number the
    reference's own scanRegistration stored per point.  The restated ring rule gives those rings; the
    re-extracted ring-major cloud keeps the stored order; the feature counts have the documented caps."""
    g = np.load(GOLD)
    xyz, ring = g["xyz"], g["ring"].astype(np.int32)
    assert np.array_equal(oracle.ring_of("OS1-64", xyz), ring)
    A = oracle.scan_registration("OS1-64", xyz, 0.5)     # minimum_range of aloam_mulran.launch:9
    assert len(A["full"]) == len(xyz) and np.array_equal(bits(A["full"][:, :3]), bits(xyz))   # already ring-major: stable
    assert np.array_equal(np.rint(A["full"][:, 3]).astype(np.int32), ring)
    rings = len(np.unique(ring))
    assert 0 < len(A["sharp"]) <= 2 * 6 * rings and len(A["sharp"]) < len(A["less_sharp"]) <= 20 * 6 * rings
    assert 0 < len(A["flat"]) <= 4 * 6 * rings and len(A["less_flat"]) > 5000


@pytest.mark.gpu
def test_cuda_features_on_a_real_scan_of_the_reference(s2m, built):
    g = np.load(GOLD)
    xyz = g["xyz"]
    want = oracle.scan_registration("OS1-64", xyz, 0.5)
    F = s2m.FeatureExtractor("OS1-64", 0.5, batch=1, cap_points=len(xyz) + 16)
    F.extract(xyz, np.array([0, len(xyz)], np.int32))
    for k in CLOUDS:
        got, _ = F.cloud(k)
        assert got.shape == want[k].shape and np.array_equal(bits(got), bits(want[k])), k
    full, _ = F.cloud("full")
    assert np.array_equal(np.rint(full[:, 3]).astype(np.int32), g["ring"].astype(np.int32))   # the rings the reference stored


@pytest.mark.gpu
@pytest.mark.parametrize("sensor,batch", [("VLP16", 1), ("HDL64", 3), ("OS1-64", 2)])
def test_cuda_features_bit_identical_to_oracle(s2m, built, sensor, batch):
    mr = harness.LAUNCH[sensor]["minimum_range"]
    sw = sweeps(sensor, 5, batch)
    sw[-1] = sw[-1].copy()
    sw[-1][::97] = np.nan                       # removeNaNFromPointCloud (:138)
    sw[-1][5::211] *= 1e-3                      # removeClosedPointCloud (:139)
    F = s2m.FeatureExtractor(sensor, mr, batch=batch)
    for rep in range(2):                        # second call: no state may leak between calls
        if rep == 1:
            sw = sw[::-1]
        off = np.cumsum([0] + [len(x) for x in sw]).astype(np.int32)
        F.extract(np.concatenate(sw), off)
        want = [oracle.scan_registration(sensor, x, mr) for x in sw]
        for k in CLOUDS:
            got, o = F.cloud(k)
            for b in range(batch):
                g, w = got[o[b]:o[b + 1]], want[b][k]
                assert g.shape == w.shape, (k, b, g.shape, w.shape)
                assert np.array_equal(bits(g), bits(w)), (k, b)
    assert F.launch_count() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("sensor", ["VLP16", "HDL32", "HDL64", "OS1-64"])
def test_cuda_ring_and_time_rules_at_the_bucket_edges(s2m, built, sensor):
    """Adversarial sweep: elevation angles packed around every ring-bucket edge of the sensor's rule
    (:168-213) at offsets from 1e-7 to 1e-2 degrees, azimuths covering the half-turn and wrap-around
    thresholds of the relative-time rule (:218-250).  The device screens these decisions in FP32 and
    falls back to the exact FP64 path near a threshold: the ring-major cloud must still be the oracle's."""
    rng = np.random.default_rng(42)
    edges = {"VLP16": np.arange(-16.0, 17.0, 2.0), "HDL32": np.arange(-92.0 / 3, 12.0, 4.0 / 3),
             "HDL64": np.r_[2.0 - (np.arange(0, 34) - 0.5) / 3.0, -8.83 - (np.arange(0, 33) - 0.5) / 2.0, 2.0, -8.83, -24.33],
             "OS1-64": np.arange(-23.5, 107.0, 2.0)}[sensor]
    offs = np.r_[0.0, np.geomspace(1e-7, 1e-2, 12), -np.geomspace(1e-7, 1e-2, 12)]
    elev = np.deg2rad((edges[:, None] + offs[None, :]).ravel())
    elev = np.tile(elev, 12)
    n = len(elev)
    az0 = 0.3
    # one clockwise turn; extra samples right at the half turn and at the end of the sweep
    az = az0 - np.sort(np.r_[rng.uniform(0, 2 * np.pi + 0.02, n - 60), np.pi + np.linspace(-3e-4, 3e-4, 30),
                             2 * np.pi + np.linspace(-3e-4, 3e-4, 30)])
    rng.shuffle(elev)
    rad = rng.uniform(5.0, 40.0, n)
    xyz = np.c_[rad * np.cos(elev) * np.cos(az), rad * np.cos(elev) * np.sin(az), rad * np.sin(elev)].astype(np.float32)
    want = oracle.scan_registration(sensor, xyz, 0.1)
    assert len(want["full"]) > 0.3 * n
    F = s2m.FeatureExtractor(sensor, 0.1, batch=1, cap_points=n + 16)
    F.extract(xyz, np.array([0, n], np.int32))
    for k in CLOUDS:
        got, _ = F.cloud(k)
        assert got.shape == want[k].shape and np.array_equal(bits(got), bits(want[k])), k


@pytest.mark.gpu
def test_cuda_features_degenerate_sweeps(s2m, built):
    F = s2m.FeatureExtractor("VLP16", 0.1, batch=3, cap_points=40000)
    good = sweeps("VLP16", 3, 1)[0]
    tiny = good[:7]
    off = np.array([0, 0, len(tiny), len(tiny) + len(good)], np.int32)    # empty, < 12 points, normal
    F.extract(np.concatenate([tiny, good]), off)
    want = oracle.scan_registration("VLP16", good, 0.1)
    for k in CLOUDS:
        got, o = F.cloud(k)
        assert o[1] == 0 and o[2] == 0
        assert np.array_equal(bits(got[o[2]:o[3]]), bits(want[k]))
    with pytest.raises(s2m.S2MError):
        F.extract(np.zeros((50000, 3), np.float32), np.array([0, 50000, 50000, 50000], np.int32))


@pytest.mark.gpu
def test_features_feed_the_mapping_call_on_the_device(s2m, built):
    """raw sweeps -> s2m_fx_extract -> s2m_register_batch_dev with the device clouds: the poses equal the
    CPU chain oracle.scan_registration -> oracle mapper (1e-4 m / 1e-5 rad; observed ~1e-15)."""
    import torch  # only to prove nothing else is needed: the hand-off is raw device pointers
    assert torch.cuda.is_available()
    seed, n = 20261018, 5
    truth = harness.trajectory(seed, n, 0.5)
    odom = harness.odometry(seed, truth)
    F = s2m.FeatureExtractor("VLP16", 0.1, batch=1)
    R = s2m.Registrar(0.2, 0.4)
    O = oracle.Oracle(0.2, 0.4)
    for f in range(n):
        xyz = harness.scan(seed, "VLP16", truth[f], f)
        F.extract(xyz, np.array([0, len(xyz)], np.int32))
        st, q, t = R.register_batch_ptr(F.device_cloud("less_sharp"), F.offsets("less_sharp"), F.device_cloud("less_flat"),
                                        F.offsets("less_flat"), odom[f, :4], odom[f, 4:], True)
        A = oracle.scan_registration("VLP16", xyz, 0.1)
        rc, qo, to = O.register(A["less_sharp"], A["less_flat"], odom[f, :4], odom[f, 4:])
        assert st[0] == rc
        assert np.linalg.norm(t[0] - to) < 1e-4 and np.abs(q[0] - qo).max() < 1e-5, f
