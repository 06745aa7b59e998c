"""Error paths and state handling of the C ABI around the hot path (round-1 advisor findings):
frames are atomic, a checkpoint larger than the frame staging buffers loads, getters do not move
the window, capacities are validated at create."""
import numpy as np
import pytest

import harness
import oracle

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def seq(built):
    return harness.sequence(20261018, "HDL64", 10)


def test_checkpoint_larger_than_the_frame_buffers_resumes_bit_exact(s2m, seq, tmp_path):
    """cap_*_in (the per-frame staging) is far smaller than the map: s2m_map_upload / s2m_checkpoint_load push the
    map in chunks.  Chunked upload == one-shot upload, resume == uninterrupted run, maps and poses bit for bit."""
    truth, odom, frames = seq
    O = oracle.Oracle(0.4, 0.8)
    for f in range(6):
        O.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
    cm, sm = O.get_map(0), O.get_map(1)
    small_kw = dict(cap_corner_in=2048, cap_surf_in=3000, cap_map_corner=1 << 17, cap_map_surf=1 << 17)
    assert len(cm) > 4 * 2048 and len(sm) > 2 * 3000  # several chunks per class: the case round 1 could save but not load
    A = s2m.Registrar(0.4, 0.8, **small_kw)
    big = s2m.Registrar(0.4, 0.8, cap_map_corner=1 << 17, cap_map_surf=1 << 17)
    assert A.map_upload(cm, sm) == 0 and big.map_upload(cm, sm) == 0
    for cls in (0, 1):
        assert np.array_equal(bits(A.map_download(cls)), bits(big.map_download(cls)))

    def thin(f):  # frames that fit the small staging buffers
        return frames[f][0][:2000], frames[f][1][::12][:2900]

    for f in (6, 7):
        c, s = thin(f)
        ra, qa, ta = A.register(c, s, odom[f, :4], odom[f, 4:])
        rb, qb, tb = big.register(c, s, odom[f, :4], odom[f, 4:])
        assert ra == rb == 0 and np.array_equal(qa, qb) and np.array_equal(ta, tb), f
    prefix = str(tmp_path / "big")
    A.checkpoint_save(prefix)
    Bc = s2m.Registrar(0.4, 0.8, **small_kw)
    assert Bc.checkpoint_load(prefix) == 0
    for cls in (0, 1):
        assert np.array_equal(bits(A.map_download(cls)), bits(Bc.map_download(cls)))
    for f in (8, 9):
        c, s = thin(f)
        ra, qa, ta = A.register(c, s, odom[f, :4], odom[f, 4:])
        rb, qb, tb = Bc.register(c, s, odom[f, :4], odom[f, 4:])
        assert ra == rb and np.array_equal(qa, qb) and np.array_equal(ta, tb), f
    for cls in (0, 1):
        assert np.array_equal(bits(A.map_download(cls)), bits(Bc.map_download(cls)))
    # a checkpoint that cannot fit is refused BEFORE the slot is touched
    small = s2m.Registrar(0.4, 0.8, cap_map_corner=64, cap_map_surf=64)
    small.register(frames[0][0][:40], frames[0][1][:40], odom[0, :4], odom[0, 4:])
    w0, c0 = small.window().copy(), [x.copy() for x in small.correction()]
    with pytest.raises(s2m.S2MError):
        small.checkpoint_load(prefix)
    assert np.array_equal(small.window(), w0) and all(np.array_equal(a, b) for a, b in zip(small.correction(), c0))
    assert len(small.map_download(1)) > 0


def test_a_failed_frame_leaves_the_map_and_the_window_untouched(s2m, seq):
    """A frame that overflows cap_map_* returns S2M_ERR_CAPACITY and changes nothing: map, window, valid block
    and correction equal those of a context that never saw it, and so does everything registered afterwards."""
    truth, odom, frames = seq
    big = s2m.Registrar(0.4, 0.8, cap_map_corner=1 << 17, cap_map_surf=1 << 18)
    for f in range(3):
        big.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
    n_s = len(big.map_download(1))
    tight = s2m.Registrar(0.4, 0.8, cap_map_corner=1 << 17, cap_map_surf=n_s + 200)
    for f in range(3):
        tight.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
    for cls in (0, 1):
        assert np.array_equal(bits(big.map_download(cls)), bits(tight.map_download(cls)))
    # frame 3, shifted 120 m so that it would also move the valid block, brings thousands of new surf voxels
    far = odom[3, 4:] + np.array([120.0, 0.0, 0.0])
    with pytest.raises(s2m.S2MError):
        tight.register(frames[3][0], frames[3][1], odom[3, :4], far)
    assert np.array_equal(tight.window(), big.window())
    for cls in (0, 1):
        assert np.array_equal(bits(big.map_download(cls)), bits(tight.map_download(cls)))
    for a, b in zip(tight.correction(), big.correction()):
        assert np.array_equal(a, b)
    # both continue with a frame that fits: identical results
    few = frames[3][1][:150]
    ra, qa, ta = big.register(frames[3][0], few, odom[3, :4], odom[3, 4:])
    rb, qb, tb = tight.register(frames[3][0], few, odom[3, :4], odom[3, 4:])
    assert ra == rb and np.array_equal(qa, qb) and np.array_equal(ta, tb)
    for cls in (0, 1):
        assert np.array_equal(bits(big.map_download(cls)), bits(tight.map_download(cls)))


def test_getters_do_not_move_the_window(s2m, seq):
    """s2m_get_local_map / s2m_debug_knn around a far-away centre must not shift the live window."""
    truth, odom, frames = seq
    A = s2m.Registrar(0.4, 0.8)
    Bc = s2m.Registrar(0.4, 0.8)
    for f in range(3):
        A.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
        Bc.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
    far = odom[2, 4:] + np.array([420.0, -380.0, 0.0])
    A.local_map(1, far)
    A.debug_knn(0, far, np.zeros((4, 3), np.float32))
    assert np.array_equal(A.window(), Bc.window())
    for f in range(3, 6):
        ra, qa, ta = A.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
        rb, qb, tb = Bc.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
        assert ra == rb and np.array_equal(qa, qb) and np.array_equal(ta, tb)
    for cls in (0, 1):
        assert np.array_equal(bits(A.map_download(cls)), bits(Bc.map_download(cls)))


def test_register_on_a_multi_lane_context_goes_through_the_lane_worker(s2m, seq):
    truth, odom, frames = seq
    A = s2m.Registrar(0.4, 0.8)
    M = s2m.Registrar(0.4, 0.8, batch=2, lanes=2)
    for f in range(4):
        ra, qa, ta = A.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
        rm, qm, tm = M.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
        assert ra == rm and np.array_equal(qa, qm) and np.array_equal(ta, tm)


def test_blocking_host_waits_change_nothing_but_the_waiting(s2m, seq, monkeypatch):
    """S2M_SYNC=block: the three host waits of a frame sleep on blocking-sync events instead of spinning (hosts with
    fewer cores than ranks x lanes).  Same poses, bit for bit, for a plain and a two-lane context."""
    truth, odom, frames = seq
    A = s2m.Registrar(0.4, 0.8)
    ref = [A.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:]) for f in range(4)]
    monkeypatch.setenv("S2M_SYNC", "block")
    B = s2m.Registrar(0.4, 0.8)
    M = s2m.Registrar(0.4, 0.8, batch=2, lanes=2)
    for f in range(4):
        for R in (B, M):
            rc, q, t = R.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
            assert rc == ref[f][0] and np.array_equal(q, ref[f][1]) and np.array_equal(t, ref[f][2])


def test_capacities_are_validated_at_create(s2m, built):
    # leaf 4 m: 13 voxels per cube axis -> the arrival field of the update key is 18 bits
    with pytest.raises(s2m.S2MError):
        s2m.Registrar(4.0, 4.0, cap_corner_in=1 << 18, cap_surf_in=1024, cap_map_corner=1024, cap_map_surf=1024)
    R = s2m.Registrar(4.0, 4.0, cap_corner_in=(1 << 18) - 1, cap_surf_in=1024, cap_map_corner=1024, cap_map_surf=1024)
    R.close()


def test_guard_bands_stay_intact(s2m, seq, monkeypatch):
    """compute-sanitizer is not available on the GPU pool: with S2M_GUARD_BYTES every device buffer gets pattern
    bands before and after it, checked after every call.  A replay with optimisation, a window shift, a chunked
    upload, the getters, a batch on two lanes and the odometry: no band is touched."""
    monkeypatch.setenv("S2M_GUARD_BYTES", "4096")
    truth, odom, frames = seq
    R = s2m.Registrar(0.4, 0.8, trace=True, cap_map_corner=1 << 17, cap_map_surf=1 << 17)
    for f in range(5):
        rc, q, t = R.register(frames[f][0], frames[f][1], odom[f, :4], odom[f, 4:])
    far = odom[4, 4:] + np.array([460.0, 0.0, 0.0])     # shifts the window, evicts cubes
    R.register(frames[5][0], frames[5][1], odom[5, :4], far)
    R.local_map(1, far)
    R.debug_knn(1, far, frames[5][1][:300, :3])
    R.surround()
    R.transform_cloud(frames[5][1])
    cm, sm = R.map_download(0), R.map_download(1)
    assert R.guard_check() == 0
    S = s2m.Registrar(0.4, 0.8, cap_corner_in=1500, cap_surf_in=2000, cap_map_corner=1 << 17, cap_map_surf=1 << 17)
    S.map_upload(cm, sm)                                  # chunked
    S.register(frames[6][0][:1400], frames[6][1][:1900], odom[6, :4], far)
    assert S.guard_check() == 0
    M = s2m.Registrar(0.4, 0.8, batch=4, lanes=2)
    for f in range(3):
        c, s_ = frames[f]
        M.register_batch(np.tile(c, (4, 1)), np.arange(5) * len(c), np.tile(s_, (4, 1)), np.arange(5) * len(s_),
                         np.tile(odom[f, :4], (4, 1)), np.tile(odom[f, 4:], (4, 1)))
    assert M.guard_check() == 0
    D = s2m.Odometer()
    for f in range(2):
        A = oracle.scan_registration("HDL64", harness.scan(20261018, "HDL64", truth[f], f), 5.0)
        D.step(A["sharp"], A["flat"], A["less_sharp"], A["less_flat"])
    assert D.L.s2m_debug_guard_check(D.h) == 0
