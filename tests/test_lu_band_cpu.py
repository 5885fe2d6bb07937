"""CPU model of the band-limited, tiled LU of csrc/lu.cu + csrc/lu_band.cuh.

The GPU kernels restrict the reference's dense loops (gauss_elim.h:9-60) to the
band of the coarse operator and re-schedule the solve over warps.  This file
re-enacts exactly that schedule in plain Python floats (IEEE doubles, no FMA)
-- same per-row operation order, same block/lane bookkeeping, same "which block
must be finished before this chunk" arithmetic -- and checks it against the
oracle's dense loops BIT FOR BIT (sign of zero included for the factor).  The
-m gpu tests then check the kernels themselves against the same oracle.
"""
import math
import struct
from fractions import Fraction

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


def fma(a, b, c):
    """correctly rounded a*b + c (Python 3.12 has no math.fma), signed zeros included"""
    r = Fraction(a) * Fraction(b) + Fraction(c)
    if r != 0:
        return float(r)
    if a == 0 or b == 0:  # (+-0) + c with c = +-0: -0 only if both are -0
        prod_neg = math.copysign(1.0, a) * math.copysign(1.0, b) < 0
        if c == 0:
            return -0.0 if prod_neg and math.copysign(1.0, c) < 0 else 0.0
    return 0.0  # exact cancellation of non-zeros: +0 in round-to-nearest


def extract_tiles(lu, bw):
    """k_lu_extract_tiles: per block row NT+1 transposed 32x32 tiles of L and of U"""
    n = lu.shape[0]
    NT, NB = (bw + 31) // 32, (n + 31) // 32
    lt = np.zeros((NB, NT + 1, 32, 32))
    ut = np.zeros((NB, NT + 1, 32, 32))
    for R in range(NB):
        for t in range(NT + 1):
            for jj in range(32):
                for l in range(32):
                    row = 32 * R + l
                    col = 32 * (R - NT + t) + jj
                    if row < n and 0 <= col < row:
                        lt[R, t, jj, l] = lu[row, col]
                    col = 32 * (R + NT - t) + jj
                    if row < n and row < col < n:
                        ut[R, t, jj, l] = lu[row, col]
    ud = np.diag(lu).copy()
    return lt, ut, ud, 1.0 / ud, NT


def tile_solve_model(lt, ut, ud, rd, NT, b, exact):
    """lu_band_solve<EXACT>, blocks processed in publication order; asserts that every
    block a tile reads has been published; returns (x, every fast quotient was the
    correctly rounded one)"""
    n = len(b)
    NB = (n + 31) // 32
    xs = [float(v) for v in b] + [0.0] * (NB * 32 - n)
    ok = True
    done_f = 0
    for R in range(NB):
        sums = [0.0] * 32
        order = [(t, R - NT + t) for t in range(NT - 1)]
        if NT >= 1 and R >= 1:
            order.append((NT - 1, R - 1))
        for t, P in order:
            if P < 0:
                continue
            assert done_f >= P + 1
            for jj in range(32):
                for l in range(32):
                    sums[l] = sums[l] + lt[R, t, jj, l] * xs[32 * P + jj]
        assert done_f >= R
        bi = [xs[32 * R + l] if 32 * R + l < n else 0.0 for l in range(32)]
        mine = [0.0] * 32
        for jj in range(32):
            z = bi[jj] - sums[jj]
            mine[jj] = z
            for l in range(32):
                sums[l] = sums[l] + lt[R, NT, jj, l] * z
        for l in range(32):
            if 32 * R + l < n:
                xs[32 * R + l] = mine[l]
        done_f = R + 1
    done_b = 0
    for Rr in range(NB):
        R = NB - 1 - Rr
        sums = [0.0] * 32
        order = [(t, R + NT - t) for t in range(NT - 1)]
        if NT >= 1 and R + 1 < NB:
            order.append((NT - 1, R + 1))
        for t, P in order:
            if P >= NB:
                continue
            assert done_b >= NB - P
            for jj in range(31, -1, -1):
                for l in range(32):
                    sums[l] = sums[l] + ut[R, t, jj, l] * xs[32 * P + jj]
        assert done_b >= Rr
        zi = [xs[32 * R + l] if 32 * R + l < n else 0.0 for l in range(32)]
        di = [ud[32 * R + l] if 32 * R + l < n else 1.0 for l in range(32)]
        yi = [rd[32 * R + l] if 32 * R + l < n else 1.0 for l in range(32)]
        mine = [0.0] * 32
        for jj in range(31, -1, -1):
            t_ = zi[jj] - sums[jj]
            if exact:
                x = t_ / di[jj]
            else:
                q0 = t_ * yi[jj]
                x = fma(fma(-di[jj], q0, t_), yi[jj], q0)
                if struct.pack("<d", x) != struct.pack("<d", t_ / di[jj]):
                    ok = False
            mine[jj] = x
            for l in range(32):
                sums[l] = sums[l] + ut[R, NT, jj, l] * x
        for l in range(32):
            if 32 * R + l < n:
                xs[32 * R + l] = mine[l]
        done_b = Rr + 1
    return np.array(xs[:n]), ok


def band_solve_model(lu, bw, b):
    lt, ut, ud, rd, NT = extract_tiles(lu, bw)
    x, ok = tile_solve_model(lt, ut, ud, rd, NT, b, exact=False)
    if not ok:  # the kernel's fallback
        x, _ = tile_solve_model(lt, ut, ud, rd, NT, b, exact=True)
    return x, ok


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
    for seed in (1, 2):
        b = seeded((n,), seed)
        x_want = orc.lu_solve(want, b)
        x_got, fast_ok = band_solve_model(got, bw, b)
        assert np.array_equal(bits(x_got), bits(x_want))
        assert fast_ok, "the fast quotient should practically always be the rounded one"


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
    b = seeded((n,), 6)
    assert np.array_equal(bits(band_solve_model(got, bw, b)[0]), bits(orc.lu_solve(want, b)))


def test_fast_quotient_is_the_rounded_quotient():
    """q = fma(fma(-d, t*y, t), y, t*y) with y = RN(1/d) against t/d on random pairs, plus
    signed zeros and large magnitudes (the kernel checks every quotient it forms)"""
    rng = np.random.Generator(np.random.PCG64(7))
    n = 20000
    t = rng.uniform(-1, 1, n) * 10.0 ** rng.integers(-30, 30, n)
    d = rng.uniform(0.5, 4, n) * rng.choice([-1.0, 1.0], n) * 2.0 ** rng.integers(-20, 20, n)
    t[:6] = [0.0, -0.0, 1.0, -1.0, 3.0, 1e300]
    d[:6] = [-6.0, -6.0, 3.0, 3.0, 1.0, 2.0]
    bad = 0
    for a, b in zip(t.tolist(), d.tolist()):
        y = 1.0 / b
        q0 = a * y
        q = fma(fma(-b, q0, a), y, q0)
        bad += struct.pack("<d", q) != struct.pack("<d", a / b)
    assert bad == 0


def test_signed_zero_multipliers_outside_the_band(orc):
    """the reference stores z = (+0)*(1/a_ii) = -0. below the band wherever the
    pivot is negative (interior rows: -6/h^2); the fix-up pass reproduces it"""
    A = orc.coarse_matrix((5, 3, 3), 0.25)
    orc.lu_factor(A)
    n = A.shape[0]
    neg = [struct.pack("<d", A[r, c]) == struct.pack("<d", -0.0)
           for r in range(n) for c in range(r) if r - c > 9]
    assert any(neg), "expected -0. multipliers below the band in the reference factor"
