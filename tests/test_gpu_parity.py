"""Parity of the CUDA path (through the C ABI) against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): kNN indices and float distances bit-exact; poses within
1e-4 m / 1e-5 rad of the reference-path restatement (observed agreement is ~1e-9)."""
import os

import numpy as np
import pytest

import harness
import oracle
from conftest import rot_angle

pytestmark = pytest.mark.gpu

TOL_T, TOL_R = 1e-4, 1e-5  # metres, radians (north_star)


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def seq_hdl(built):
    return harness.sequence(20261018, "HDL64", 12)


@pytest.fixture(scope="module")
def seq_vlp(built):
    return harness.sequence(7, "VLP16", 10, step_m=0.5)


@pytest.fixture(scope="module")
def seq_os1(built):
    return harness.sequence(13, "OS1-64", 8)


def pick(sensor, seq_hdl, seq_vlp, seq_os1):
    # launch-file parameters: HDL-64 0.4/0.8, VLP-16 0.2/0.4, OS1-64 (MulRan) 0.4/0.8
    return {"hdl": (seq_hdl, 0.4, 0.8), "vlp": (seq_vlp, 0.2, 0.4), "os1": (seq_os1, 0.4, 0.8)}[sensor]


def run_oracle(seq, n, line, plane, **kw):
    truth, odom, frames = seq
    O = oracle.Oracle(line, plane, **kw)
    poses = []
    for f in range(n):
        rc, q, t = O.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
        poses.append(np.r_[q, t])
    return O, np.array(poses)


def test_scan_voxel_filter_bit_exact(s2m, seq_hdl):
    """Row V on the incoming clouds: same points, same order, same bits."""
    truth, odom, frames = seq_hdl
    R = s2m.Registrar(0.4, 0.8, trace=True)
    O = oracle.Oracle(0.4, 0.8, trace=True)
    c, s = frames[0]
    R.register(c, s, odom[0, :4], odom[0, 4:])
    O.register(c, s, odom[0, :4], odom[0, 4:])
    for cls in (0, 1):
        got, want = R.trace_cloud(cls), O.trace_cloud(cls)
        assert got.shape == want.shape and len(got) > 100
        assert np.array_equal(bits(got), bits(want))
    assert R.stats.n_corner_ds == O.stats.n_corner_ds and R.stats.n_surf_ds == O.stats.n_surf_ds


def test_first_frame_takes_guard_path(s2m, seq_hdl):
    truth, odom, frames = seq_hdl
    R = s2m.Registrar(0.4, 0.8)
    rc, q, t = R.register(frames[0][0], frames[0][1], odom[0, :4], odom[0, 4:])
    assert rc == s2m.S2M_MAP_TOO_SMALL and R.stats.optimized == 0
    assert np.array_equal(np.r_[q, t], odom[0])  # identity correction: pose == odometry guess, bit for bit
    rc, q, t = R.register(frames[1][0], frames[1][1], odom[1, :4], odom[1, 4:])
    assert rc == 0 and R.stats.optimized == 1


@pytest.mark.parametrize("sensor", ["hdl", "vlp", "os1"])
def test_map_evolution_bit_exact_without_optimisation(s2m, seq_hdl, seq_vlp, seq_os1, sensor):
    """Rows B, C, V, I, W with pose = guess on both sides: maps stay bit-identical frame after frame."""
    seq, line, plane = pick(sensor, seq_hdl, seq_vlp, seq_os1)
    truth, odom, frames = seq
    R = s2m.Registrar(line, plane, skip_optimization=True)
    O = oracle.Oracle(line, plane, skip_optimization=True)
    for f in range(8):
        c, s = frames[f]
        _, qg, tg = R.register(c, s, odom[f, :4], odom[f, 4:])
        _, qo, to = O.register(c, s, odom[f, :4], odom[f, 4:])
        assert np.array_equal(np.r_[qg, tg], np.r_[qo, to])
        if f in (0, 3, 7):
            for cls in (0, 1):
                got, want = R.map_download(cls), O.get_map(cls)
                assert got.shape == want.shape, (f, cls, got.shape, want.shape)
                assert np.array_equal(bits(got), bits(want)), (f, cls)
    sg, so = R.surround(), O.surround()  # row X: /laser_cloud_surround, gathered on the device
    assert sg.shape == so.shape and np.array_equal(bits(sg), bits(so))
    lm_g, lm_o = R.local_map(1, odom[7, 4:]), O.local_map(1, odom[7, 4:])
    assert np.array_equal(bits(lm_g), bits(lm_o))


def test_knn_bit_exact_on_uploaded_map(s2m, seq_hdl):
    """Row K through s2m_debug_knn: indices and float distances identical to the canonical
    brute force and to the KD-tree, on the gather order both sides define."""
    O1, _ = run_oracle(seq_hdl, 8, 0.4, 0.8)
    truth, odom, frames = seq_hdl
    cm, sm = O1.get_map(0), O1.get_map(1)
    R = s2m.Registrar(0.4, 0.8)
    O = oracle.Oracle(0.4, 0.8)
    assert R.map_upload(cm, sm) == 0 and O.map_upload(cm, sm) == 0
    centre = odom[8, 4:]
    rng = np.random.default_rng(5)
    ties = 0
    for cls, mp in ((0, cm), (1, sm)):
        lg, lo = R.local_map(cls, centre), O.local_map(cls, centre)
        assert np.array_equal(bits(lg), bits(lo)) and len(lg) > 1000
        # queries: map points jittered (dense hits) plus far points (gate misses)
        q = mp[rng.integers(0, len(mp), 4000), :3] + rng.normal(size=(4000, 3)).astype(np.float32) * 0.3
        q = np.r_[q, rng.uniform(-200, 200, (200, 3))].astype(np.float32)
        ig, dg = R.debug_knn(cls, centre, q)
        ib, db = O.debug_knn(cls, centre, q, method=0)
        ik, dk = O.debug_knn(cls, centre, q, method=1)
        gate = db[:, 4] < 1.0
        assert gate.sum() > 1000
        assert np.array_equal(gate, dg[:, 4] < 1.0)          # same queries pass the reference's gate
        assert np.array_equal(ig[gate], ib[gate])            # indices bit-exact
        assert np.array_equal(bits(dg[gate]), bits(db[gate]))  # float distances bit-exact
        assert (ig[~gate] == -1).all()
        same = (ik[gate] == ib[gate]).all(1)
        ties += int((~same).sum())                           # KD-tree differs only on exact ties
        assert np.array_equal(bits(dk[gate]), bits(db[gate]))
    assert ties <= 2


def test_knn_ties_on_a_lattice_map(s2m, built):
    """Row K on a map whose points sit on a regular lattice: every query has many candidates at exactly equal float
    distances, including at the fifth place, so the (d2, index) tie rule decides both WHICH five neighbours are
    returned and their order.  Queries come clustered (whole groups of the grouped search share a 2 m block) and
    spread out; s2m_debug_knn also cross-checks the grouped search against the thread-per-query one on the device."""
    g = np.arange(-6.0, 6.01, 0.5, dtype=np.float32)
    X, Y, Z = np.meshgrid(g + 3.0, g - 2.0, g[:9] + 1.0, indexing="ij")
    lat = np.c_[X.ravel(), Y.ravel(), Z.ravel(), np.zeros(X.size)].astype(np.float32)
    rng = np.random.default_rng(11)
    cm = lat[rng.permutation(len(lat))]          # arrival order decides the index inside a cube's cloud
    sm = lat[::3].copy()
    R = s2m.Registrar(0.4, 0.8)
    O = oracle.Oracle(0.4, 0.8)
    assert R.map_upload(cm, sm) == 0 and O.map_upload(cm, sm) == 0
    centre = np.array([3.0, -2.0, 1.0])
    # lattice nodes, cell centres and face centres (maximal ties), jittered points, far points
    nodes = lat[rng.integers(0, len(lat), 1500), :3]
    q = np.r_[nodes, nodes + np.float32(0.25), nodes + np.array([0.25, 0.0, 0.0], np.float32),
              nodes + rng.normal(size=nodes.shape).astype(np.float32) * 0.1,
              rng.uniform(-40, 40, (300, 3))].astype(np.float32)
    for cls in (0, 1):
        lg, lo = R.local_map(cls, centre), O.local_map(cls, centre)
        assert np.array_equal(bits(lg), bits(lo)) and len(lg) > 1000
        for order in (np.arange(len(q)), np.lexsort((q[:, 0], q[:, 1], q[:, 2]))):
            ig, dg = R.debug_knn(cls, centre, q[order])
            ib, db = O.debug_knn(cls, centre, q[order], method=0)
            gate = db[:, 4] < 1.0
            assert gate.sum() > 1000
            assert np.array_equal(gate, dg[:, 4] < 1.0)
            assert np.array_equal(ig[gate], ib[gate])
            assert np.array_equal(bits(dg[gate]), bits(db[gate]))
            if cls == 0:
                tied_fifth = db[gate][:, 3] == db[gate][:, 4]
                assert tied_fifth.sum() > 500      # the tie rule really is exercised


@pytest.mark.parametrize("sensor", ["hdl", "vlp", "os1"])
def test_registration_matches_oracle_on_uploaded_map(s2m, seq_hdl, seq_vlp, seq_os1, sensor):
    """One full registration (rows A..W) from identical maps: kNN of the first outer iteration
    bit-exact inside the real flow, reduced normal equations and poses within tolerance -- for the
    three sensors / launch-file resolutions of BASELINE configs 1-3."""
    seq, line, plane = pick(sensor, seq_hdl, seq_vlp, seq_os1)
    truth, odom, frames = seq
    k = len(frames) - 1  # map from all frames but the last, which is then registered
    O1, _ = run_oracle(seq, k, line, plane)
    cm, sm = O1.get_map(0), O1.get_map(1)
    R = s2m.Registrar(line, plane, trace=True)
    O = oracle.Oracle(line, plane, trace=True, use_kdtree=False)
    R.map_upload(cm, sm)
    O.map_upload(cm, sm)
    c, s = frames[k]
    rg, qg, tg = R.register(c, s, odom[k, :4], odom[k, 4:])
    ro, qo, to = O.register(c, s, odom[k, :4], odom[k, 4:])
    assert rg == ro == 0
    assert (R.stats.n_map_corner, R.stats.n_map_surf) == (O.stats.n_map_corner, O.stats.n_map_surf)
    for cls in (0, 1):
        ig, dg, ug = R.trace_knn(0, cls)
        io, do, uo = O.trace_knn(0, cls)
        gate = do[:, 4] < 1.0
        assert len(ig) == len(io) and gate.sum() > 200
        assert np.array_equal(ig[gate], io[gate]) and np.array_equal(bits(dg[gate]), bits(do[gate]))
        assert np.array_equal(ug, uo)  # same correspondences accepted (no threshold flips)
    assert list(R.stats.n_edge) == list(O.stats.n_edge) and list(R.stats.n_plane) == list(O.stats.n_plane)
    for outer in (0, 1):
        pg, sg, itg, ng, termg = R.trace_lm(outer)
        po, so, ito, no, termo = O.trace_lm(outer)
        assert np.allclose(sg, so, rtol=1e-9, atol=1e-9 * np.abs(so).max())
        assert (ng, termg) == (no, termo)
        assert np.abs(pg - po).max() < 1e-9
    assert np.linalg.norm(tg - to) < TOL_T and rot_angle(qg, qo) < TOL_R
    assert np.linalg.norm(tg - to) < 1e-8  # what we actually observe


@pytest.mark.parametrize("sensor", ["hdl", "vlp", "os1"])
def test_stream_poses_within_tolerance(s2m, seq_hdl, seq_vlp, seq_os1, sensor):
    """N-frame replay, each side on its own map: per-scan poses agree within the north_star tolerance."""
    seq, line, plane = pick(sensor, seq_hdl, seq_vlp, seq_os1)
    truth, odom, frames = seq
    n = len(frames)
    O, poses_o = run_oracle(seq, n, line, plane)
    R = s2m.Registrar(line, plane)
    worst_t = worst_r = 0.0
    for f in range(n):
        rc, q, t = R.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
        worst_t = max(worst_t, float(np.linalg.norm(t - poses_o[f, 4:])))
        worst_r = max(worst_r, rot_angle(q, poses_o[f, :4]))
    assert worst_t < TOL_T and worst_r < TOL_R, (worst_t, worst_r)
    if sensor == "hdl":  # and mapping did its job: closer to the truth than the drifting odometry
        assert np.linalg.norm(t - truth[n - 1, 4:]) < np.linalg.norm(odom[n - 1, 4:] - truth[n - 1, 4:])
    qc, tc = R.correction()
    assert abs(np.linalg.norm(qc) - 1) < 1e-9


def test_perturbed_initial_guesses_against_a_mature_map(s2m, built):
    """SURVEY 8d "T_init distribution": truth o perturbation (U(-0.2,0.2) m, U(-1,1) deg per axis, and a
    second set at 0.5 m / 3 deg) against a map built from 40 frames at the true poses. Both sides start
    from the same uploaded map; poses must agree within tolerance for every draw."""
    from conftest import quat_mul
    truth, odom, frames = harness.sequence(20261018, "HDL64", 44)
    B = oracle.Oracle(0.4, 0.8, skip_optimization=True)
    for f in range(40):
        B.register(frames[f][0], frames[f][1], truth[f, :4], truth[f, 4:])
    cm, sm = B.get_map(0), B.get_map(1)
    assert len(cm) + len(sm) > 40000
    rng = np.random.default_rng(20261018)
    worst_t = worst_r = 0.0
    for trial, (dt, dr) in enumerate([(0.2, 1.0)] * 4 + [(0.5, 3.0)] * 2):
        f = 40 + trial % 4
        rv = np.deg2rad(rng.uniform(-dr, dr, 3))
        a = np.linalg.norm(rv)
        dq = np.r_[np.sin(a / 2) * rv / a, np.cos(a / 2)]
        q0 = quat_mul(truth[f, :4], dq)
        t0 = truth[f, 4:] + rng.uniform(-dt, dt, 3)
        R = s2m.Registrar(0.4, 0.8)
        O = oracle.Oracle(0.4, 0.8)
        R.map_upload(cm, sm)
        O.map_upload(cm, sm)
        rg, qg, tg = R.register(frames[f][0], frames[f][1], q0, t0)
        ro, qo, to = O.register(frames[f][0], frames[f][1], q0, t0)
        assert rg == ro == 0
        assert list(R.stats.n_edge) == list(O.stats.n_edge) and list(R.stats.n_plane) == list(O.stats.n_plane)
        worst_t = max(worst_t, float(np.linalg.norm(tg - to)))
        worst_r = max(worst_r, rot_angle(qg, qo))
        # the registration pulled the perturbed guess towards the truth
        assert np.linalg.norm(tg - truth[f, 4:]) < 0.5 * np.linalg.norm(t0 - truth[f, 4:]) + 0.05
    assert worst_t < TOL_T and worst_r < TOL_R, (worst_t, worst_r)


@pytest.mark.parametrize("lanes", [0, 1, 2, 3])
def test_batch_slots_are_independent_and_identical_to_single(s2m, seq_hdl, seq_vlp, lanes):
    """register_batch over 3 slots (plain context, or 1, 2, 3 concurrent lanes) == 3 single-slot contexts, bit for bit."""
    truth, odom, frames = seq_hdl
    n = 6
    B = 3
    RB = s2m.Registrar(0.4, 0.8, batch=B, lanes=lanes, cap_map_corner=1 << 17, cap_map_surf=1 << 18)
    singles = [s2m.Registrar(0.4, 0.8, cap_map_corner=1 << 17, cap_map_surf=1 << 18) for _ in range(B)]
    for f in range(n):
        # slot b replays the sequence with a lag of b frames (different inputs per slot, ragged sizes)
        fr = [max(f - b, 0) for b in range(B)]
        corner = np.concatenate([frames[i][0] for i in fr])
        surf = np.concatenate([frames[i][1] for i in fr])
        co = np.cumsum([0] + [len(frames[i][0]) for i in fr]).astype(np.int32)
        so = np.cumsum([0] + [len(frames[i][1]) for i in fr]).astype(np.int32)
        q = np.array([odom[i, :4] for i in fr])
        t = np.array([odom[i, 4:] for i in fr])
        active = np.array([1 if f - b >= 0 else 0 for b in range(B)], np.int32)
        st, qo, to = RB.register_batch(corner, co, surf, so, q, t, active)
        for b in range(B):
            if not active[b]:
                continue
            rc, q1, t1 = singles[b].register(frames[fr[b]][0], frames[fr[b]][1], odom[fr[b], :4], odom[fr[b], 4:])
            assert st[b] == rc
            assert np.array_equal(qo[b], q1) and np.array_equal(to[b], t1), (f, b)


def test_submit_wait_pipeline_matches_synchronous_calls(s2m, seq_hdl):
    """s2m_register_batch_submit/_wait with two frames in flight (the copy of frame f+1 overlapping frame f)
    returns, frame by frame, the bits of the synchronous calls; a third submit is refused; a plain
    context refuses the pair."""
    truth, odom, frames = seq_hdl
    n, B = 6, 2
    kw = dict(cap_map_corner=1 << 17, cap_map_surf=1 << 18)
    RA = s2m.Registrar(0.4, 0.8, batch=B, lanes=2, **kw)
    RS = s2m.Registrar(0.4, 0.8, batch=B, lanes=0, **kw)
    packed = []
    for f in range(n):
        fr = [f, max(f - 1, 0)]
        corner = np.ascontiguousarray(np.concatenate([frames[i][0] for i in fr]), np.float32)
        surf = np.ascontiguousarray(np.concatenate([frames[i][1] for i in fr]), np.float32)
        co = np.cumsum([0] + [len(frames[i][0]) for i in fr]).astype(np.int32)
        so = np.cumsum([0] + [len(frames[i][1]) for i in fr]).astype(np.int32)
        packed.append((corner, co, surf, so, np.array([odom[i, :4] for i in fr]), np.array([odom[i, 4:] for i in fr])))
    want = [RS.register_batch(*p) for p in packed]
    got = []
    for f, (corner, co, surf, so, q, t) in enumerate(packed):
        if f >= 2:
            got.append(RA.wait())
        RA.submit(corner.ctypes.data, co, surf.ctypes.data, so, q, t)
    with pytest.raises(s2m.S2MError):
        c, co, su, so, q, t = packed[-1]
        RA.submit(c.ctypes.data, co, su.ctypes.data, so, q, t)      # two already in flight
    got.append(RA.wait())
    got.append(RA.wait())
    for f in range(n):
        assert np.array_equal(got[f][0], want[f][0])
        assert np.array_equal(got[f][1], want[f][1]) and np.array_equal(got[f][2], want[f][2]), f
    for cls in (0, 1):
        for b in range(B):
            assert np.array_equal(bits(RA.map_download(cls, b)), bits(RS.map_download(cls, b)))
    with pytest.raises(s2m.S2MError):
        c, co, su, so, q, t = packed[0]
        RS.submit(c.ctypes.data, co, su.ctypes.data, so, q, t)


def test_checkpoint_resume_reproduces_the_uninterrupted_run(s2m, seq_hdl, tmp_path):
    """s2m_checkpoint_save after frame 3, s2m_checkpoint_load into a new context (slot 1 of a batch of 2):
    frames 4..7 give the bits of the run that never stopped; so does the map at the end."""
    truth, odom, frames = seq_hdl
    kw = dict(cap_map_corner=1 << 17, cap_map_surf=1 << 18)
    A = s2m.Registrar(0.4, 0.8, **kw)
    for f in range(4):
        A.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
    prefix = str(tmp_path / "ckpt")
    A.checkpoint_save(prefix)
    for ext in (".corner.pcd", ".surf.pcd", ".state"):
        assert os.path.getsize(prefix + ext) > 0
    assert len(s2m.pcd_read(prefix + ".surf.pcd")) == len(A.map_download(1))
    Bc = s2m.Registrar(0.4, 0.8, batch=2, **kw)
    assert Bc.checkpoint_load(prefix, slot=1) == 0
    assert np.array_equal(Bc.window(1), A.window()) and np.array_equal(Bc.correction(1)[1], A.correction()[1])
    e = np.zeros((0, 4), np.float32)
    for f in range(4, 8):
        rc, q, t = A.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
        co, so = np.array([0, 0, len(frames[f][0])], np.int32), np.array([0, 0, len(frames[f][1])], np.int32)
        st, qb, tb = Bc.register_batch(frames[f][0], co, frames[f][1], so, np.tile(odom[f, :4], (2, 1)), np.tile(odom[f, 4:], (2, 1)),
                                       active=[0, 1])
        assert st[1] == rc and np.array_equal(qb[1], q) and np.array_equal(tb[1], t), f
    for cls in (0, 1):
        assert np.array_equal(bits(A.map_download(cls)), bits(Bc.map_download(cls, 1)))


def test_window_shift_and_eviction(s2m, built):
    """Drive the sensor 400 m along +x in 60 m hops with a tiny synthetic cloud: the window must
    follow (cen decreases), cubes that leave it are dropped, exactly like the oracle."""
    rng = np.random.default_rng(9)
    R = s2m.Registrar(0.4, 0.8, skip_optimization=True)
    O = oracle.Oracle(0.4, 0.8, skip_optimization=True)
    for step in range(9):
        x = 60.0 * step - 30.0
        corner = np.c_[rng.uniform(-20, 20, (300, 3)), np.zeros(300)].astype(np.float32)
        surf = np.c_[rng.uniform(-140, 140, (3000, 2)), rng.uniform(-3, 3, 3000), np.zeros(3000)].astype(np.float32)
        q, t = np.array([0, 0, 0.0, 1.0]), np.array([x, -70.0 * step, 0.3 * step])
        R.register(corner, surf, q, t)
        O.register(corner, surf, q, t)
        assert np.array_equal(R.window(), O.window()), step
    assert not np.array_equal(O.window(), [10, 10, 5])
    for cls in (0, 1):
        got, want = R.map_download(cls), O.get_map(cls)
        assert got.shape == want.shape and np.array_equal(bits(got), bits(want))


def test_empty_and_tiny_inputs(s2m, built):
    R = s2m.Registrar(0.4, 0.8)
    e = np.zeros((0, 4), np.float32)
    rc, q, t = R.register(e, e, [0, 0, 0, 1.0], [1.0, 2.0, 3.0])
    assert rc == s2m.S2M_MAP_TOO_SMALL and np.array_equal(t, [1.0, 2.0, 3.0])
    one = np.array([[1, 2, 3, 0.5]], np.float32)
    rc, q, t = R.register(one, one, [0, 0, 0, 1.0], [0.0, 0.0, 0.0])
    assert rc == s2m.S2M_MAP_TOO_SMALL
    assert len(R.map_download(0)) == 1 and len(R.map_download(1)) == 1
    with pytest.raises(s2m.S2MError):
        R.register(np.zeros((20000, 4), np.float32), e, [0, 0, 0, 1.0], [0, 0, 0.0])  # > cap_corner_in


def test_transform_cloud_bit_exact(s2m, seq_hdl):
    truth, odom, frames = seq_hdl
    R = s2m.Registrar(0.4, 0.8)
    R.register(frames[0][0], frames[0][1], odom[0, :4], odom[0, 4:])
    pts = frames[0][1]
    got = R.transform_cloud(pts)
    q, t = odom[0, :4], odom[0, 4:]
    v = pts[:, :3].astype(np.float64)
    u = np.broadcast_to(q[:3], v.shape)
    uv = np.cross(u, v)
    uv = uv + uv
    want = ((v + q[3] * uv) + np.cross(u, uv) + t).astype(np.float32)
    assert np.array_equal(bits(got[:, :3]), bits(want)) and np.array_equal(got[:, 3], pts[:, 3])


def test_contexts_release_their_memory(s2m, built):
    """create / use / destroy every kind of context a few times: device memory returns to where it was."""
    import torch
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info()
    rng = np.random.default_rng(1)
    pts = np.c_[rng.uniform(-20, 20, (4000, 3)), np.zeros(4000)].astype(np.float32)
    for rep in range(4):
        R = s2m.Registrar(0.4, 0.8, batch=4, lanes=2)
        R.register_batch(np.tile(pts[:500], (4, 1)), np.arange(5) * 500, np.tile(pts, (4, 1)), np.arange(5) * 4000,
                         np.tile([0, 0, 0, 1.0], (4, 1)), np.zeros((4, 3)))
        R.close()
        F = s2m.FeatureExtractor("VLP16", 0.1, batch=2, cap_points=30000)
        F.close()
        D = s2m.Odometer(batch=2)
        D.close()
    torch.cuda.synchronize()
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < 64 << 20, (free0, free1)      # CUDA keeps a little for itself; a leak would be GBs
