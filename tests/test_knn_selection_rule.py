"""The exactness argument of the grouped kNN selection (csrc/s2m_kernels.cu, knn5_group step 3/4), checked as a
property on the CPU with numpy -- a model of the rule, not of the CUDA code (the CUDA code is compared with the oracle
and with the thread-per-query search in the -m gpu tests).

Rule: every candidate gets the key (float bits of d2 with the low B bits replaced by its position); the six smallest
keys are kept.  If the fifth and the sixth key differ above the position bits, the five kept candidates are exactly the
five smallest under the reference's order (d2, index); otherwise everything whose truncated distance is <= the fifth
key's is ordered exactly.  Seeds: the gate 1.0f (laserMapping.cpp:585 / :653) -- candidates at d2 >= 1 never enter."""
import numpy as np
import pytest


def select(d2, tags, bits):
    """-> (gate, indices of the five nearest in (d2, tag) order, took_second_pass) by the kernel's rule"""
    mask = np.uint32((1 << bits) - 1)
    u = d2.view(np.uint32)
    keys = (u & ~mask) | np.arange(len(d2), dtype=np.uint32)
    seed = np.float32(1.0).view(np.uint32)
    six = np.sort(np.r_[keys, np.full(6, seed, np.uint32)])[:6]
    k4, k5 = six[4], six[5]
    if k4 >= seed:
        return False, None, False
    if (k4 & ~mask) == (k5 & ~mask):  # near-tie at the fifth place: exact order of everything up to that truncated distance
        cand = np.nonzero((u & ~mask) <= (k4 & ~mask))[0]
        second = True
    else:
        cand = (six[:5] & mask).astype(np.int64)
        second = False
    order = sorted(cand.tolist(), key=lambda i: (d2[i], tags[i]))
    return True, order[:5], second


def exact(d2, tags):
    order = sorted(range(len(d2)), key=lambda i: (d2[i], tags[i]))
    if len(order) < 5 or not d2[order[4]] < np.float32(1.0):
        return False, None
    return True, order[:5]


@pytest.mark.parametrize("bits", [8, 10])
def test_truncated_key_selection_is_exact(bits):
    rng = np.random.default_rng(1234 + bits)
    second_passes = gates = 0
    for trial in range(3000):
        n = int(rng.integers(0, 1 << bits))
        kind = trial % 5
        if kind == 0:    # generic
            d2 = rng.uniform(0, 3, n)
        elif kind == 1:  # many exact ties (lattice-like)
            d2 = rng.choice(np.array([0.0, 0.0625, 0.125, 0.25, 0.3125, 0.5, 0.5625, 0.75, 0.99999994, 1.0, 1.25]), n)
        elif kind == 2:  # near-ties: values a few ulps apart around a common distance
            base = np.float32(rng.uniform(0.05, 0.9))
            d2 = (np.full(n, base, np.float32).view(np.uint32) + rng.integers(0, 3000, n).astype(np.uint32)).view(np.float32)
        elif kind == 3:  # around the gate
            d2 = (np.full(n, np.float32(1.0), np.float32).view(np.uint32) - np.uint32(40) + rng.integers(0, 80, n).astype(np.uint32)).view(np.float32)
        else:            # tiny and zero distances (denormal keys)
            d2 = rng.choice(np.array([0.0, 1e-45, 1e-40, 1e-38, 1e-30, 1e-10, 0.5]), n)
        d2 = np.ascontiguousarray(d2, np.float32)
        tags = rng.permutation(max(n, 1))[:n].astype(np.uint32)  # unique tags, unrelated to the pool position
        g1, i1, sp = select(d2, tags, bits)
        g0, i0 = exact(d2, tags)
        assert g1 == g0, (trial, kind)
        if g0:
            gates += 1
            assert i1 == i0, (trial, kind)
        second_passes += sp
    assert gates > 500 and second_passes > 50  # both branches of the rule were exercised
