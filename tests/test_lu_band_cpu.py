"""CPU model of the band-limited LU of csrc/lu.cu + csrc/lu_band.cuh.

The GPU kernels restrict the reference's dense loops (gauss_elim.h:9-60) to the
band of the coarse operator and re-schedule the solve over warps.  This file
re-enacts exactly that schedule in plain Python floats (IEEE doubles, no FMA)
-- same per-row operation order, same block/lane bookkeeping, same "which block
must be finished before this chunk" arithmetic -- and checks it against the
oracle's dense loops BIT FOR BIT (sign of zero included for the factor).  The
-m gpu tests then check the kernels themselves against the same oracle.
"""
import struct

import numpy as np
import pytest

W = 4  # kLuWarps


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def band_factor_model(a, bw):
    """k_lu_factor_band + k_lu_finish_lower on a dense copy"""
    a = a.copy()
    n = a.shape[0]
    for p in range(n - 1):
        w = min(bw, n - 1 - p)
        pinv = 1.0 / a[p, p]
        for rr in range(w):
            r = p + 1 + rr
            z = a[r, p] * pinv
            for cc in range(w):
                c = p + 1 + cc
                a[r, c] = a[r, c] - z * a[p, c]
    for r in range(n):
        for c in range(r):
            pinv = 1.0 / a[c, c]
            a[r, c] = a[r, c] * pinv if r - c <= bw else 0.0 * pinv
    return a


def extract_band(lu, bw):
    n = lu.shape[0]
    lb = np.zeros((max(bw, 1), n))
    ub = np.zeros((max(bw, 1), n))
    for d in range(1, bw + 1):
        for i in range(n):
            if i - d >= 0:
                lb[d - 1, i] = lu[i, i - d]
            if i + d < n:
                ub[d - 1, i] = lu[i, i + d]
    return lb, ub, np.diag(lu).copy()


def band_solve_model(lb, ub, ud, bw, b):
    """lu_band_solve, blocks processed in publication order; asserts that every
    value a chunk reads belongs to a block the chunk waited for"""
    n = len(b)
    NB = (n + 31) // 32
    xs = [float(v) for v in b] + [0.0] * (NB * 32 - n)
    done_f = 0
    for R in range(NB):
        i0 = 32 * R
        sums = [0.0] * 32
        dmax = min(bw, i0 + 31)
        dc = dmax
        while dc >= 1:
            dlow = max(dc - 15, 1)
            jmax = min(i0 + 31 - dlow, i0 - 1)
            need = (jmax >> 5) + 1 if jmax >= 0 else 0
            assert done_f >= need
            for t in range(16):
                d = dc - t
                for lane in range(32):
                    i = i0 + lane
                    if i < n and d >= 1 and d > lane and d <= i:
                        assert (i - d) >> 5 < need  # a finished block
                        sums[lane] = sums[lane] + lb[d - 1, i] * xs[i - d]
            dc -= 16
        assert done_f >= R
        tri = [[(lb[lane - jj - 1, i0 + lane] if (i0 + lane < n and lane - jj <= bw) else 0.0)
                for jj in range(lane)] for lane in range(32)]
        bi = [xs[i0 + lane] if i0 + lane < n else 0.0 for lane in range(32)]
        mine = [0.0] * 32
        for jj in range(32):
            z = bi[jj] - sums[jj]
            mine[jj] = z
            for lane in range(jj + 1, 32):
                sums[lane] = sums[lane] + tri[lane][jj] * z
        for lane in range(32):
            if i0 + lane < n:
                xs[i0 + lane] = mine[lane]
        done_f = R + 1
    done_b = 0
    for Rr in range(NB):
        R = NB - 1 - Rr
        i0 = 32 * R
        sums = [0.0] * 32
        dmax = min(bw, n - 1 - i0)
        dc = dmax
        while dc >= 1:
            dlow = max(dc - 15, 1)
            jmin = i0 + max(dlow, 32)
            need = NB - (jmin >> 5) if jmin < n else 0
            assert done_b >= need
            for t in range(16):
                d = dc - t
                for lane in range(32):
                    i = i0 + lane
                    if i < n and d >= 1 and d > 31 - lane and i + d < n:
                        assert NB - 1 - ((i + d) >> 5) < need
                        sums[lane] = sums[lane] + ub[d - 1, i] * xs[i + d]
            dc -= 16
        assert done_b >= Rr
        zi = [xs[i0 + lane] if i0 + lane < n else 0.0 for lane in range(32)]
        di = [ud[i0 + lane] if i0 + lane < n else 1.0 for lane in range(32)]
        mine = [0.0] * 32
        for jj in range(31, -1, -1):
            x = (zi[jj] - sums[jj]) / di[jj]
            mine[jj] = x
            for lane in range(jj):
                u = ub[jj - lane - 1, i0 + lane] if (i0 + lane < n and jj - lane <= bw and
                                                    i0 + jj < n) else 0.0
                sums[lane] = sums[lane] + u * x
        for lane in range(32):
            if i0 + lane < n:
                xs[i0 + lane] = mine[lane]
        done_b = Rr + 1
    return np.array(xs[:n])


@pytest.mark.parametrize("coarse", [(3, 3, 3), (5, 3, 3), (3, 5, 3), (5, 5, 3), (9, 3, 3)])
def test_band_model_equals_dense_oracle(orc, coarse):
    from oracle_lib import seeded
    n = int(np.prod(coarse))
    bw = min(coarse[1] * coarse[2], n - 1)
    A = orc.coarse_matrix(coarse, 0.125)
    want = A.copy()
    orc.lu_factor(want)
    got = band_factor_model(A, bw)
    assert np.array_equal(bits(got), bits(want)), "factor (sign of zero included)"
    lb, ub, ud = extract_band(got, bw)
    for seed in (1, 2):
        b = seeded((n,), seed)
        x_want = orc.lu_solve(want, b)
        x_got = band_solve_model(lb, ub, ud, bw, b)
        assert np.array_equal(bits(x_got), bits(x_want))


def test_band_model_on_a_generic_banded_matrix(orc):
    """not a Laplacian: random diagonally dominant band, n not a multiple of 32,
    bandwidth wider than two row blocks"""
    from oracle_lib import seeded
    n, bw = 150, 70
    A = seeded((n, n), 5)
    for r in range(n):
        for c in range(n):
            if abs(r - c) > bw:
                A[r, c] = 0.0
        A[r, r] = -(abs(A[r, r]) + 2.0 * bw)
    want = A.copy()
    orc.lu_factor(want)
    got = band_factor_model(A, bw)
    assert np.array_equal(bits(got), bits(want))
    lb, ub, ud = extract_band(got, bw)
    b = seeded((n,), 6)
    assert np.array_equal(bits(band_solve_model(lb, ub, ud, bw, b)), bits(orc.lu_solve(want, b)))


def test_signed_zero_multipliers_outside_the_band(orc):
    """the reference stores z = (+0)*(1/a_ii) = -0. below the band wherever the
    pivot is negative (interior rows: -6/h^2); the fix-up pass reproduces it"""
    A = orc.coarse_matrix((5, 3, 3), 0.25)
    orc.lu_factor(A)
    n = A.shape[0]
    neg = [struct.pack("<d", A[r, c]) == struct.pack("<d", -0.0)
           for r in range(n) for c in range(r) if r - c > 9]
    assert any(neg), "expected -0. multipliers below the band in the reference factor"
