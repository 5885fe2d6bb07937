"""GPU parity, operator by operator: every CUDA kernel of the hot path against
the CPU oracle on the same seeded inputs, BIT-EXACT for the fields (the
reference's operation order is reproduced without FMA contraction) and 1e-13
relative for norms (only the summation order differs).  All calls go through
the C ABI (include/mgb.h)."""
import math

import numpy as np
import pytest

from oracle_lib import OrcMG, seeded

pytestmark = pytest.mark.gpu

NORM_RTOL = 1e-13

# (coarse extents, levels): cubes like the reference plus boxes
HIERARCHIES = [((3, 3, 3), 4), ((5, 5, 5), 3), ((3, 5, 9), 3), ((9, 3, 5), 2), ((3, 3, 3), 6)]


def _mk(mgb, coarse, levels, gs=2):
    return mgb.Solver(coarse, levels, gs)


@pytest.fixture(params=["tile", "plain"], autouse=True)
def kernel_family(request, mgb):
    """every test runs twice: with the TMA tile kernels (tile.cu) forced onto
    every level they can handle, and with the plain kernels only"""
    if request.param == "tile":
        mgb.set_global(mgb.G_TILE, 1)
        mgb.set_global(mgb.G_TILE_MIN_PLANE, 0)
    else:
        mgb.set_global(mgb.G_TILE, 0)
    yield request.param
    mgb.set_global(mgb.G_TILE, 1)
    mgb.set_global(mgb.G_TILE_MIN_PLANE, 40000)


@pytest.mark.parametrize("coarse,levels", HIERARCHIES)
def test_pack_roundtrip(mgb, coarse, levels):
    with _mk(mgb, coarse, levels) as s:
        for lvl in range(levels):
            a = seeded(s.dims(lvl), 10 + lvl)
            for which in (mgb.MGB_U, mgb.MGB_D, mgb.MGB_R):
                s.upload(lvl, which, a)
                assert np.array_equal(s.download(lvl, which), a)


@pytest.mark.parametrize("coarse,levels", HIERARCHIES)
def test_half_sweeps_bitwise(mgb, orc, coarse, levels):
    with _mk(mgb, coarse, levels) as s:
        for lvl in range(levels):
            shape = s.dims(lvl)
            h = s.spacing(lvl)
            v = seeded(shape, 1)
            d = seeded(shape, 2)
            s.upload(lvl, mgb.MGB_U, v)
            s.upload(lvl, mgb.MGB_D, d)
            for colour in (1, 0, 0, 1):
                orc.half_sweep(v, d, h, colour)
                s.half_sweep(lvl, colour)
                got = s.download(lvl, mgb.MGB_U)
                assert np.array_equal(got, v), f"level {lvl} colour {colour}"


@pytest.mark.parametrize("first_red", [True, False])
@pytest.mark.parametrize("iters", [1, 3])
def test_smoother_bitwise(mgb, orc, first_red, iters):
    with _mk(mgb, (5, 3, 9), 4) as s:
        lvl = 3
        shape, h = s.dims(lvl), s.spacing(lvl)
        v, d = seeded(shape, 3), seeded(shape, 4)
        s.upload(lvl, mgb.MGB_U, v)
        s.upload(lvl, mgb.MGB_D, d)
        orc.smooth(v, d, h, iters, first_red)
        s.smooth(lvl, iters, first_red)
        assert np.array_equal(s.download(lvl, mgb.MGB_U), v)


@pytest.mark.parametrize("coarse,levels", HIERARCHIES)
def test_residual(mgb, orc, coarse, levels):
    with _mk(mgb, coarse, levels) as s:
        for lvl in range(levels):
            shape, h = s.dims(lvl), s.spacing(lvl)
            v, d = seeded(shape, 5), seeded(shape, 6)
            r0 = seeded(shape, 7)  # boundary of r must survive
            s.upload(lvl, mgb.MGB_U, v)
            s.upload(lvl, mgb.MGB_D, d)
            s.upload(lvl, mgb.MGB_R, r0)
            want_r = r0.copy()
            want = orc.residual(v, d, h, want_r)
            got = s.residual(lvl, store_r=True)
            assert np.array_equal(s.download(lvl, mgb.MGB_R), want_r)
            assert got == pytest.approx(want, rel=NORM_RTOL)
            # norm-only form (res == NULL) leaves r alone
            s.upload(lvl, mgb.MGB_R, r0)
            got2 = s.residual(lvl, store_r=False)
            assert got2 == pytest.approx(got, rel=NORM_RTOL)
            assert np.array_equal(s.download(lvl, mgb.MGB_R), r0)


@pytest.mark.parametrize("coarse,levels", HIERARCHIES)
def test_restrict_bitwise(mgb, orc, coarse, levels):
    with _mk(mgb, coarse, levels) as s:
        for lvl in range(1, levels):
            r = seeded(s.dims(lvl), 8)  # non-zero boundary: exercises the injection
            dc = seeded(s.dims(lvl - 1), 9)
            s.upload(lvl, mgb.MGB_R, r)
            s.upload(lvl - 1, mgb.MGB_D, dc)
            orc.restrict(r, dc)
            s.restrict(lvl)
            assert np.array_equal(s.download(lvl - 1, mgb.MGB_D), dc), f"level {lvl}"


@pytest.mark.parametrize("coarse,levels", HIERARCHIES)
def test_residual_restrict_fused_bitwise(mgb, orc, coarse, levels):
    with _mk(mgb, coarse, levels) as s:
        for lvl in range(1, levels):
            shape, h = s.dims(lvl), s.spacing(lvl)
            v, d = seeded(shape, 15), seeded(shape, 16)
            s.upload(lvl, mgb.MGB_U, v)
            s.upload(lvl, mgb.MGB_D, d)
            s.zero(lvl, mgb.MGB_R)
            r = np.zeros(shape)
            orc.residual(v, d, h, r)
            dc = np.zeros(s.dims(lvl - 1))
            orc.restrict(r, dc)
            s.residual_restrict(lvl)
            assert np.array_equal(s.download(lvl - 1, mgb.MGB_D), dc), f"level {lvl}"


# hierarchies whose fine levels span several tiles of the tile kernels (tile.cu)
BIG = [((3, 3, 3), 7), ((3, 5, 9), 5), ((5, 3, 3), 6), ((3, 3, 17), 5)]


@pytest.mark.parametrize("coarse,levels", HIERARCHIES + BIG)
@pytest.mark.parametrize("colour", [0, 1])
def test_sweep_residual_fused_bitwise(mgb, orc, coarse, levels, colour):
    """half-sweep + residual norm in one kernel == the two reference stages"""
    with _mk(mgb, coarse, levels) as s:
        for lvl in range(levels):
            shape, h = s.dims(lvl), s.spacing(lvl)
            v, d = seeded(shape, 51), seeded(shape, 52)
            s.upload(lvl, mgb.MGB_U, v)
            s.upload(lvl, mgb.MGB_D, d)
            orc.half_sweep(v, d, h, colour)
            want = orc.residual(v, d, h, None)
            got = s.sweep_residual(lvl, colour)
            assert np.array_equal(s.download(lvl, mgb.MGB_U), v), f"level {lvl}"
            assert got == pytest.approx(want, rel=NORM_RTOL), f"level {lvl}"
            # and the plain norm on the same state
            assert s.residual(lvl, store_r=False) == pytest.approx(want, rel=NORM_RTOL)


@pytest.mark.parametrize("coarse,levels", HIERARCHIES + BIG)
@pytest.mark.parametrize("colour", [0, 1])
def test_sweep_residual_restrict_fused_bitwise(mgb, orc, coarse, levels, colour):
    """half-sweep + residual + restriction in one kernel == the three stages"""
    with _mk(mgb, coarse, levels) as s:
        for lvl in range(1, levels):
            shape, h = s.dims(lvl), s.spacing(lvl)
            v, d = seeded(shape, 53), seeded(shape, 54)
            s.upload(lvl, mgb.MGB_U, v)
            s.upload(lvl, mgb.MGB_D, d)
            s.upload(lvl - 1, mgb.MGB_D, seeded(s.dims(lvl - 1), 55))  # must be overwritten
            orc.half_sweep(v, d, h, colour)
            r = np.zeros(shape)
            orc.residual(v, d, h, r)
            dc = np.zeros(s.dims(lvl - 1))
            orc.restrict(r, dc)
            s.sweep_residual_restrict(lvl, colour)
            assert np.array_equal(s.download(lvl, mgb.MGB_U), v), f"level {lvl}"
            assert np.array_equal(s.download(lvl - 1, mgb.MGB_D), dc), f"level {lvl}"


@pytest.mark.parametrize("coarse,levels", BIG)
def test_residual_restrict_big_bitwise(mgb, orc, coarse, levels):
    with _mk(mgb, coarse, levels) as s:
        lvl = levels - 1
        shape, h = s.dims(lvl), s.spacing(lvl)
        v, d = seeded(shape, 56), seeded(shape, 57)
        s.upload(lvl, mgb.MGB_U, v)
        s.upload(lvl, mgb.MGB_D, d)
        r = np.zeros(shape)
        want = orc.residual(v, d, h, r)
        dc = np.zeros(s.dims(lvl - 1))
        orc.restrict(r, dc)
        s.residual_restrict(lvl)
        assert np.array_equal(s.download(lvl - 1, mgb.MGB_D), dc)
        assert s.residual(lvl, store_r=False) == pytest.approx(want, rel=NORM_RTOL)


@pytest.mark.parametrize("coarse,levels", HIERARCHIES)
def test_prolong_correct_bitwise(mgb, orc, coarse, levels):
    with _mk(mgb, coarse, levels) as s:
        for lvl in range(1, levels):
            ec = seeded(s.dims(lvl - 1), 11)
            ef = seeded(s.dims(lvl), 12)
            s.upload(lvl - 1, mgb.MGB_U, ec)
            s.upload(lvl, mgb.MGB_U, ef)
            orc.prolong_correct(ec, ef)
            s.prolong_correct(lvl)
            assert np.array_equal(s.download(lvl, mgb.MGB_U), ef), f"level {lvl}"


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


# (17, 9, 9) = the coarsest grid of BASELINE config 5 on 8 GPUs: 1377 unknowns
@pytest.mark.parametrize("coarse", [(3, 3, 3), (5, 5, 5), (3, 5, 9), (9, 9, 9), (3, 9, 9),
                                    (17, 9, 9)])
def test_coarse_lu_bitwise(mgb, orc, coarse):
    """band-limited factor + warp-cooperative band solve (csrc/lu.cu, lu_band.cuh)
    against the oracle's dense loops (gauss_elim.h:9-60): the factor down to the
    sign of its zeros, the solution bit for bit"""
    levels = 2
    with _mk(mgb, coarse, levels) as s:
        hc = s.spacing(0)
        A = orc.coarse_matrix(coarse, hc)
        orc.lu_factor(A)
        n, bw, secs = s.coarse_info()
        assert n == int(np.prod(coarse)) and bw == min(coarse[1] * coarse[2], n - 1) and secs > 0
        assert np.array_equal(_bits(s.coarse_lu()), _bits(A))
        for seed in (13, 14):
            b = seeded(coarse, seed)
            s.upload(0, mgb.MGB_D, b)
            s.coarse_solve()
            want = orc.lu_solve(A, b.reshape(-1)).reshape(coarse)
            assert np.array_equal(_bits(s.download(0, mgb.MGB_U)), _bits(want))


@pytest.mark.parametrize("n,bw", [(3, 1), (40, 39), (150, 70), (257, 5)])
def test_host_lu_generic_matrix(mgb, orc, n, bw):
    """convertToLU_InPlace / solveWithLU on caller-owned matrices (test_lu.c flow):
    any band structure, n not a multiple of the 32-row blocks"""
    A = seeded((n, n), 41)
    r, c = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    A[np.abs(r - c) > bw] = 0.0
    A[np.arange(n), np.arange(n)] = -(np.abs(np.diag(A)) + 2.0 * bw)
    want = A.copy()
    orc.lu_factor(want)
    mgb.host_lu_factor(A)
    assert np.array_equal(_bits(A), _bits(want))
    b = seeded((n,), 42)
    assert np.array_equal(_bits(mgb.host_lu_solve(A, b)), _bits(orc.lu_solve(want, b)))


def test_known_answer_lu(mgb):
    # gauss_elim.h:99-124 (commented mini-test): x = [5.5, 8, 6.5]
    a = np.array([[2., -1, 0], [-1, 2, -1], [0, -1, 2]])
    mgb.host_lu_factor(a)
    x = mgb.host_lu_solve(a, np.array([3., 4, 5]))
    assert np.allclose(x, [5.5, 8, 6.5], rtol=0, atol=1e-14)


def test_dirichlet_and_norms(mgb, orc):
    with _mk(mgb, (3, 3, 3), 5) as s:
        lvl = 4
        shape, h = s.dims(lvl), s.spacing(lvl)
        v = seeded(shape, 14)
        s.upload(lvl, mgb.MGB_U, v)
        s.set_dirichlet(lvl, mgb.MGB_U)
        orc.set_dirichlet(v, h)
        assert np.array_equal(s.download(lvl, mgb.MGB_U), v)
        assert math.sqrt(s.sumsq(lvl, mgb.MGB_U)) == pytest.approx(orc.l2norm(v), rel=NORM_RTOL)
        i, j, k = np.meshgrid(*[np.arange(n) for n in shape], indexing="ij")
        exact = (i * h) ** 2 - 2 * (j * h) ** 2 + (k * h) ** 2
        want = math.sqrt(((v - exact) ** 2).sum())
        assert math.sqrt(s.error_sumsq()) == pytest.approx(want, rel=1e-12)


@pytest.mark.parametrize("shape", [(50, 50, 50), (6, 9, 12), (17, 17, 17)])
def test_stateless_host_entry_points(mgb, orc, shape):
    """the raw-pointer API (test_rb_gs_3d.c flow) incl. even extents (N=50,
    red_black_gs_scalability.txt)"""
    h = 1.0 / (shape[2] - 1)
    v, d = seeded(shape, 21), seeded(shape, 22)
    want = v.copy()
    orc.smooth(want, d, h, 1, True)
    orc.smooth(want, d, h, 1, False)
    mgb.host_smooth(v, d, h, 1, True)
    mgb.host_smooth(v, d, h, 1, False)
    assert np.array_equal(v, want)
    r_want, r_got = seeded(shape, 23), None
    r_got = r_want.copy()
    n_want = orc.residual(want, d, h, r_want)
    n_got = mgb.host_residual(v, d, h, r_got)
    assert np.array_equal(r_got, r_want)
    assert n_got == pytest.approx(n_want, rel=NORM_RTOL)


def test_stateless_transfer_operators(mgb, orc):
    fine, coarse = (17, 9, 33), (9, 5, 17)
    r = seeded(fine, 31)
    dc_w, dc_g = np.zeros(coarse), np.zeros(coarse)
    orc.restrict(r, dc_w)
    mgb.host_restrict(r, dc_g)
    assert np.array_equal(dc_g, dc_w)
    ec = seeded(coarse, 32)
    ef_w = seeded(fine, 33)
    ef_g = ef_w.copy()
    orc.prolong_correct(ec, ef_w)
    mgb.host_prolong_correct(ec, ef_g)
    assert np.array_equal(ef_g, ef_w)
    A_w = orc.coarse_matrix((3, 5, 3), 0.25)
    A_g = mgb.host_coarse_matrix((3, 5, 3), 0.25)
    assert np.array_equal(A_g, A_w)
    orc.lu_factor(A_w)
    mgb.host_lu_factor(A_g)
    assert np.array_equal(A_g, A_w)


@pytest.mark.parametrize("coarse,levels,gs", [((3, 3, 3), 5, 2), ((5, 5, 5), 3, 1), ((3, 5, 9), 4, 3)])
@pytest.mark.parametrize("mode", ["graph", "eager", "profile", "unfused", "fuse1", "fuse2", "notail", "nozeroguess", "noprolongmask"])
def test_vcycle_bitwise_all_levels(mgb, orc, coarse, levels, gs, mode):
    """one and several V-cycles: every level's u and d match the oracle bit
    for bit, the returned norm to 1e-13"""
    from multigrid_parallel_b200.solver import (OPT_FUSE, OPT_GRAPH, OPT_PROFILE, OPT_TAIL,
                                                OPT_ZERO_GUESS)
    mg = OrcMG(orc, coarse, levels, gs)
    with _mk(mgb, coarse, levels, gs) as s:
        if mode == "eager":
            s.set_option(OPT_GRAPH, 0)
        elif mode == "profile":
            s.set_option(OPT_PROFILE, 1)
        elif mode == "unfused":
            s.set_option(OPT_FUSE, 0)
        elif mode == "fuse1":
            s.set_option(OPT_FUSE, 1)
        elif mode == "fuse2":
            s.set_option(OPT_FUSE, 2)
        elif mode == "notail":
            s.set_option(OPT_TAIL, 0)
        elif mode == "nozeroguess":
            s.set_option(OPT_ZERO_GUESS, 0)
        elif mode == "noprolongmask":
            s.set_option(6, 0)
        top = levels - 1
        u0, d0 = seeded(s.dims(top), 41), seeded(s.dims(top), 42)
        mg.u(top)[...] = u0
        mg.d(top)[...] = d0
        s.upload(top, mgb.MGB_U, u0)
        s.upload(top, mgb.MGB_D, d0)
        for cyc in range(3):
            want = mg.vcycle()
            got = s.vcycle()
            assert got == pytest.approx(want, rel=NORM_RTOL), f"cycle {cyc}"
            for lvl in range(levels):
                assert np.array_equal(s.download(lvl, mgb.MGB_U), mg.u(lvl)), (cyc, lvl)
                if lvl < top:
                    assert np.array_equal(s.download(lvl, mgb.MGB_D), mg.d(lvl)), (cyc, lvl)
        if mode == "profile":
            calls, secs = s.timing(top, 0)
            assert calls == 3 and secs > 0
        assert s.launch_count > 0
        mg.close()


@pytest.mark.parametrize("graph", [1, 0])
def test_vcycle_after_coarse_levels_were_overwritten(mgb, orc, graph):
    """the cycle no longer zeroes the coarse levels every time (their first
    half-sweep takes the zero guess as given), so it relies on their faces being
    0; arrays uploaded onto coarse levels in between (non-zero faces) must not
    leak into the next cycle (mg_3d.h:1258-1259 zeroes them, so does the oracle)"""
    from multigrid_parallel_b200.solver import OPT_GRAPH
    coarse, levels, gs = (3, 3, 3), 6, 2
    mg = OrcMG(orc, coarse, levels, gs)
    with _mk(mgb, coarse, levels, gs) as s:
        s.set_option(OPT_GRAPH, graph)
        top = levels - 1
        u0, d0 = seeded(s.dims(top), 61), seeded(s.dims(top), 62)
        mg.u(top)[...] = u0
        mg.d(top)[...] = d0
        s.upload(top, mgb.MGB_U, u0)
        s.upload(top, mgb.MGB_D, d0)
        for cyc in range(3):
            if cyc != 1:  # scribble over every coarse level before cycles 0 and 2
                for lvl in range(top):
                    s.upload(lvl, mgb.MGB_U, seeded(s.dims(lvl), 70 + lvl))
                    s.upload(lvl, mgb.MGB_D, seeded(s.dims(lvl), 80 + lvl))
            want = mg.vcycle()
            got = s.vcycle()
            assert got == pytest.approx(want, rel=NORM_RTOL), cyc
            for lvl in range(levels):
                assert np.array_equal(s.download(lvl, mgb.MGB_U), mg.u(lvl)), (cyc, lvl)
        mg.close()


def test_vcycle_keeps_the_sign_of_zero_on_the_faces(mgb, orc):
    """the reference's prolongation does `ef[p] += 0.` on every face point, which
    turns a -0. boundary value into +0.; the colour-masked prolongation of the
    cycle must do the same (np.array_equal cannot see the difference: signbit)"""
    coarse, levels, gs = (3, 3, 3), 5, 2
    mg = OrcMG(orc, coarse, levels, gs)
    with _mk(mgb, coarse, levels, gs) as s:
        top = levels - 1
        u0, d0 = seeded(s.dims(top), 91), seeded(s.dims(top), 92)
        for a in (u0,):
            a[0, :, :] = -0.0; a[-1, :, :] = -0.0
            a[:, 0, :] = -0.0; a[:, -1, :] = -0.0
            a[:, :, 0] = -0.0; a[:, :, -1] = -0.0
        mg.u(top)[...] = u0
        mg.d(top)[...] = d0
        s.upload(top, mgb.MGB_U, u0)
        s.upload(top, mgb.MGB_D, d0)
        for cyc in range(2):
            mg.vcycle()
            s.vcycle()
            got, want = s.download(top, mgb.MGB_U), mg.u(top)
            assert np.array_equal(got, want)
            assert np.array_equal(np.signbit(got), np.signbit(want)), cyc
        mg.close()
