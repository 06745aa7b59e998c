"""The validation argument of the parallel feature picks (csrc/s2m_fx.cu, select_kernel), checked as a property on the
CPU -- a model of the rule, not of the CUDA code (the CUDA code is compared bit for bit with oracle/scan_registration.cpp
in the -m gpu tests).

Reference (scanRegistration.cpp:292-394): the six sectors of a ring are walked in order; a pick marks up to five
neighbours on either side as taken (until a gap).  Rule under test: run every sector WITHOUT the marks of the sector
before it; then, in order, redo sector j+1 only if the marks sector j leaves on j+1's first five points hit a point
j+1 picked; carry the (possibly new) marks on."""
import numpy as np

K_LESS, K_SHARP, K_FLAT = 20, 2, 4


def run_sector(curv, gap, sp, ep, picked):
    """the reference's two walks over [sp, ep] on the shared `picked` array -> (sharp, less, flat) pick lists"""
    idx = np.arange(sp, ep + 1)
    order = idx[np.lexsort((idx, curv[sp:ep + 1]))]  # ascending (curvature, index): A8
    sharp, less, flat = [], [], []

    def suppress(i):
        picked[i] = 1
        for l in range(1, 6):
            if gap[i + l]:
                break
            picked[i + l] = 1
        for l in range(-1, -6, -1):
            if gap[i + l + 1]:
                break
            picked[i + l] = 1

    n_big = 0
    for i in order[::-1]:
        if not picked[i] and curv[i] > 0.1:
            n_big += 1
            if n_big <= K_SHARP:
                sharp.append(i); less.append(i)
            elif n_big <= K_LESS:
                less.append(i)
            else:
                break
            suppress(i)
    n_small = 0
    for i in order:
        if not picked[i] and curv[i] < 0.1:
            flat.append(i)
            n_small += 1
            if n_small >= K_FLAT:
                break
            suppress(i)
    return sharp, less, flat


def sectors(first, last):
    return [(first + (last - first) * j // 6, first + (last - first) * (j + 1) // 6 - 1) for j in range(6)]


def serial(curv, gap, first, last):
    picked = np.zeros(len(curv), np.uint8)
    return [run_sector(curv, gap, sp, ep, picked) for sp, ep in sectors(first, last)]


def speculative(curv, gap, first, last):
    secs = sectors(first, last)

    def alone(j, in_marks):
        sp, ep = secs[j]
        picked = np.zeros(len(curv), np.uint8)
        for t in range(5):
            if (in_marks >> t) & 1:
                picked[sp + t] = 1
        res = run_sector(curv, gap, sp, ep, picked)
        out = sum(int(picked[ep + 1 + t]) << t for t in range(5))
        chosen = set(res[1]) | set(res[2])
        pk = sum((1 << t) for t in range(5) if sp + t in chosen)
        return res, out, pk

    runs = [alone(j, 0) for j in range(6)]  # all six at the same time
    redone = 0
    for j in range(1, 6):
        marks = runs[j - 1][1]
        if marks & runs[j][2]:
            runs[j] = alone(j, marks)
            redone += 1
    return [r[0] for r in runs], redone


def test_speculative_sectors_equal_the_serial_walk():
    rng = np.random.default_rng(77)
    total_redone = 0
    for trial in range(400):
        n = int(rng.integers(60, 700))
        kind = trial % 4
        curv = rng.gamma(0.5, 0.2, n).astype(np.float32)
        if kind == 1:  # few distinct values: many ties
            curv = rng.choice(np.array([0.0, 0.05, 0.1, 0.2, 3.0], np.float32), n)
        gap = (rng.uniform(size=n) < (0.02 if kind != 2 else 0.3)).astype(np.uint8)
        first, last = 5, n - 6
        if kind == 3:  # sharp structure right at the sector borders
            for sp, ep in sectors(first, last):
                curv[max(first, sp - 3):sp + 4] = rng.uniform(1.0, 5.0, len(curv[max(first, sp - 3):sp + 4])).astype(np.float32)
            gap[:] = 0
        a = serial(curv, gap, first, last)
        b, redone = speculative(curv, gap, first, last)
        assert a == b, (trial, kind)
        total_redone += redone
    assert total_redone > 100  # the redo path was exercised
