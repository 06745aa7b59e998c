"""Host build of csrc/s2m_math.cuh (the very source the kernels compile) against the oracle."""
import ctypes

import numpy as np
import pytest

import oracle
from conftest import quat_mul


@pytest.fixture(scope="module")
def hm(built):
    H = ctypes.CDLL(built.HOSTMATH)
    H.hm_dist2.restype = ctypes.c_float
    H.hm_store_key.restype = ctypes.c_ulonglong
    H.hm_store_key.argtypes = [ctypes.c_int] * 3 + [ctypes.c_uint, ctypes.c_ulonglong]
    H.hm_voxel_rel.argtypes = [ctypes.c_float, ctypes.c_int, ctypes.c_float]
    H.hm_cube_of.argtypes = [ctypes.c_double]
    return H


def rand_q(rng, ang):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    return np.r_[ax * np.sin(ang / 2), np.cos(ang / 2)]


def _rot(q, v):
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return R @ v


def make_problem(rng, nb=400):
    qt, tt = rand_q(rng, rng.uniform(0, 3)), rng.normal(size=3) * 10
    kinds = np.zeros(nb, np.int32)
    dO, dH = np.zeros((nb, 10)), np.zeros((nb, 10))
    for i in range(nb):
        cp = rng.normal(size=3) * 20
        w = _rot(qt, cp) + tt
        if i % 3 == 0:
            u = rng.normal(size=3)
            u /= np.linalg.norm(u)
            c = w + u * rng.normal() * 0.5 + rng.normal(size=3) * (0.02 if i % 9 else 0.4)
            kinds[i] = 0
            dO[i, :3], dO[i, 3:6], dO[i, 6:9] = cp, c + 0.1 * u, c - 0.1 * u
            dH[i, :3], dH[i, 3:6], dH[i, 6:9] = cp, c, u
        else:
            n = rng.normal(size=3)
            n /= np.linalg.norm(n)
            d = -(n @ w) + rng.normal() * (0.02 if i % 7 else 0.5)
            kinds[i] = 1
            dO[i, :3], dO[i, 3:6], dO[i, 6] = cp, n, d
            dH[i] = dO[i]
    x0 = np.r_[quat_mul(rand_q(rng, np.deg2rad(rng.uniform(0, 2))), qt), tt + rng.normal(size=3) * 0.2]
    return kinds, dO, dH, x0


def test_lm_schedule_matches_oracle(hm):
    """Analytic Jacobians + 6x6 Cholesky LM (device code path) vs Jets + dense-QR LM (oracle)."""
    O = oracle.lib()
    rng = np.random.default_rng(0)
    for trial in range(25):
        kinds, dO, dH, x0 = make_problem(rng)
        xo, xh = x0.copy(), x0.copy()
        ni, te, ni2, te2 = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        O.orc_solve(kinds.ctypes.data, dO.ctypes.data, len(kinds), xo.ctypes.data, 4, None, ctypes.byref(ni), ctypes.byref(te))
        hm.hm_solve(kinds.ctypes.data, dH.ctypes.data, len(kinds), xh.ctypes.data, 4, None, ctypes.byref(ni2), ctypes.byref(te2))
        assert (ni.value, te.value) == (ni2.value, te2.value)
        assert np.abs(xo - xh).max() < 1e-11
        assert np.abs(xo - x0).max() > 1e-3  # the solve actually moved the pose


def test_lm_no_residuals(hm):
    x = np.array([0, 0, 0, 1.0, 1, 2, 3])
    ni, te = ctypes.c_int(), ctypes.c_int()
    hm.hm_solve(None, None, 0, x.ctypes.data, 4, None, ctypes.byref(ni), ctypes.byref(te))
    assert te.value == 5 and ni.value == 0 and np.array_equal(x, [0, 0, 0, 1.0, 1, 2, 3])


def test_xf_point_bit_exact(hm):
    """pointAssociateToMap: Eigen's q*p+t in double then float, no contraction."""
    rng = np.random.default_rng(1)
    q, t = rand_q(rng, 1.3), rng.normal(size=3) * 100
    pose = np.r_[q, t]
    p = (rng.normal(size=(5000, 3)) * 40).astype(np.float32)
    out = np.zeros_like(p)
    hm.hm_xf_point(pose.ctypes.data, p.ctypes.data, len(p), out.ctypes.data)
    v = p.astype(np.float64)
    u = q[:3]
    uv = np.cross(np.broadcast_to(u, v.shape), v)
    uv = uv + uv
    want = ((v + q[3] * uv) + np.cross(np.broadcast_to(u, v.shape), uv) + t).astype(np.float32)
    assert np.array_equal(out.view(np.uint32), want.view(np.uint32))


def test_cube_rule(hm):
    # int((v+25)/50), minus one when v+25 < 0: differs from floor at exact negative multiples
    cases = {0.0: 0, 24.999: 0, 25.0: 1, -25.0: 0, -25.0001: -1, -75.0: -2, -74.99: -1, -125.0: -3, 1e4: 200}
    for v, want in cases.items():
        assert hm.hm_cube_of(v) == want, v


def test_edge_and_plane_fit_match_oracle(hm):
    O = oracle.lib()
    rng = np.random.default_rng(2)
    flips = 0
    for _ in range(500):
        base = rng.normal(size=3) * 30
        if rng.uniform() < 0.5:
            nb = base + np.outer(rng.normal(size=5), rng.normal(size=3)) * 0.4 + rng.normal(size=(5, 3)) * 0.03
        else:
            nb = base + rng.normal(size=(5, 3)) * 0.3
        nb = np.ascontiguousarray(nb, np.float32)
        c, u = np.zeros(3), np.zeros(3)
        ok = hm.hm_edge_fit(nb.ctypes.data, c.ctypes.data, u.ctypes.data)
        p = nb.astype(np.float64)
        cen = p.sum(0) / 5.0
        m = np.ascontiguousarray((p - cen).T @ (p - cen))
        ev, evec = np.zeros(3), np.zeros(9)
        O.orc_eig3(m.ctypes.data, ev.ctypes.data, evec.ctypes.data)
        want_ok = ev[2] > 3 * ev[1]
        if bool(ok) != bool(want_ok):
            flips += 1
            continue
        assert np.allclose(c, cen, atol=1e-12)
        if ok:
            assert abs(abs(u @ evec.reshape(3, 3)[:, 2]) - 1) < 1e-10
        # plane
        n3, d = np.zeros(3), ctypes.c_double()
        okp = hm.hm_plane_fit(nb.ctypes.data, n3.ctypes.data, ctypes.byref(d))
        x = np.zeros(3)
        A = np.ascontiguousarray(p)
        O.orc_plane_qr(A.ctypes.data, x.ctypes.data)
        nn = np.linalg.norm(x)
        assert np.allclose(n3, x / nn, atol=1e-9) and abs(d.value - 1 / nn) < 1e-9 * max(1, 1 / nn)
        want_okp = bool((np.abs(p @ (x / nn) + 1 / nn) <= 0.2).all())
        assert bool(okp) == want_okp
    assert flips == 0


def test_store_key_order_is_gather_order(hm):
    """key order == (i, j, k) loop order, filtered entries (voxel z,y,x) before raw ones."""
    ks = [hm.hm_store_key(0, 0, 0, 0, 5), hm.hm_store_key(0, 0, 0, 0, 9), hm.hm_store_key(0, 0, 0, 1, 0),
          hm.hm_store_key(0, 0, 1, 0, 0), hm.hm_store_key(0, 1, -1, 0, 0), hm.hm_store_key(1, -5, -3, 0, 0)]
    assert ks == sorted(ks) and len(set(ks)) == len(ks)
    assert hm.hm_store_key(-1, 7, 3, 1, 2 ** 33 - 1) < hm.hm_store_key(0, -7, -3, 0, 0)


def test_voxel_rel_nonnegative_and_bounded(hm):
    rng = np.random.default_rng(3)
    for leaf in (0.2, 0.4, 0.8):
        inv = np.float32(1.0) / np.float32(leaf)
        for _ in range(2000):
            cube = int(rng.integers(-40, 40))
            lo = 50.0 * cube - 25.0
            p = np.float32(lo + rng.uniform(0, 50))
            if float(p) < lo:
                p = np.float32(lo)
            v = hm.hm_voxel_rel(float(p), cube, float(inv))
            assert 0 <= v <= int(50 / leaf) + 1
