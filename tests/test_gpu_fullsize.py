"""GPU parity at BASELINE.json's full sizes, against the golden values the
reference itself produced (tests/golden/histories.json, oracle/gen_golden.py)
and through size-independent properties.

Norm tolerance at full size: the GPU iterates are bit-identical to the
reference's (checked through the solution's SHA-256), so the only difference in
the printed norms is the ORDER in which ~1e8 squares are added.  The reference
adds them sequentially per OpenMP thread, which is itself only reproducible to
~1e-11 between team sizes at 513^3; the kernels use a tree whose result is
within 1e-14 of the exactly rounded sum (test_exact_norm_of_identical_residual).
Measured deviation of the reference's own one-thread sum from the exact one:
<1e-12 up to 65^3, 5e-12 at 129^3, 2e-11 at 257^3, 2.6e-10 at 513^3
(oracle/gen_exact_history.py prints them), and at 1025^3 the reference's
1-thread and 8-thread norms differ from each other by 6.5e-10 (first cycle:
golden 309102136.013 vs SURVEY Appendix A 309102136.214).  Hence, against the
reference's PRINTED norms: 1e-12 up to 65^3, 1e-11 at 129^3, 5e-11 at 257^3,
5e-10 at 513^3, 1e-8 at 1025^3 (measured 2.6e-9: the sequential sum's error
grows like the number of points) -- and against "history_exact" (the reference's
own residual field after every cycle, summed in long double by
oracle/gen_exact_history.py) 1e-13 at every size that has it."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def histories():
    return json.load(open(os.path.join(GOLD, "histories.json")))


def _solve(mgb, g, **opts):
    s = mgb.Solver(g["coarse"], g["levels"], g["gs"])
    top = s.levels - 1
    s.set_dirichlet(top, mgb.MGB_D)          # SolverSetupBoundaryConditions
    init = math.sqrt(s.sumsq(top, mgb.MGB_D))  # SolverGetInitialResidual
    s.set_dirichlet(top, mgb.MGB_U)          # test_mg_3d.c:29
    hist = s.solve(g["init_norm"] * g["tol"], 60)
    return s, init, hist


@pytest.mark.parametrize("key,rtol", [("3_5_2", 1e-12), ("3_5_1", 1e-12), ("3_5_3", 1e-12),
                                      ("5_4_2", 1e-12), ("9_3_2", 1e-12), ("3_6_2", 1e-12),
                                      ("3_7_2", 1e-11), ("3_8_2", 5e-11), ("3_9_2", 5e-10),
                                      ("3_10_2", 1e-8)])
def test_solve_matches_reference_golden(mgb, histories, key, rtol):
    if key not in histories:
        pytest.skip(f"golden {key} not generated (oracle/gen_golden.py --huge)")
    g = histories[key]
    s, init, hist = _solve(mgb, g)
    # GetL2NormOfVector is a sequential sum in the reference (6e-12 off the exact
    # value at 513^3); the drop-in header keeps that sum on the host, bit-exact
    assert init == pytest.approx(g["init_norm"], rel=max(rtol, 1e-12))
    assert len(hist) == g["cycles"], "V-cycle count to 1e-8*||d|| differs from the reference"
    dev = np.max(np.abs(hist - np.array(g["history"])) / np.array(g["history"]))
    assert dev <= rtol, dev
    if "history_exact" in g:  # the reference's residual fields, summed exactly
        ex = np.array(g["history_exact"])
        dev_exact = np.max(np.abs(hist - ex) / ex)
        assert dev_exact <= 1e-13, dev_exact
    u = s.download(s.levels - 1, mgb.MGB_U)
    assert hashlib.sha256(u.tobytes()).hexdigest() == g["sha256"], "solution not bit-identical"
    assert float(u[1, 2, 3]) == g["probe_1_2_3"]
    err = math.sqrt(s.error_sumsq())
    assert err == pytest.approx(g["errnorm_np"], rel=1e-6)
    s.close()


def test_exact_norm_of_identical_residual(mgb, histories):
    """the returned norm against an exactly-rounded sum (math.fsum) of the
    residual field the kernel itself stored: isolates the kernel's own
    summation error (<= 1e-14) from the reference's sequential-sum error"""
    g = histories["3_8_2"]
    s = mgb.Solver(g["coarse"], g["levels"], g["gs"])
    top = s.levels - 1
    s.set_dirichlet(top, mgb.MGB_D)
    s.set_dirichlet(top, mgb.MGB_U)
    for _ in range(2):
        s.vcycle()
    got = s.residual(top, store_r=True)
    r = s.download(top, mgb.MGB_R)
    exact = math.sqrt(math.fsum((r * r).reshape(-1).tolist()))
    assert got == pytest.approx(exact, rel=1e-14)
    s.close()


def test_1025_cycle_and_properties(mgb):
    """1025^3 (config 4, single GPU): one V-cycle reduces the residual by the
    reference's first-cycle factor (2513804842.77 ->
    309102136.21), boundary values untouched, coarse boundaries zero"""
    s = mgb.Solver(3, 10, 2)
    top = s.levels - 1
    s.set_dirichlet(top, mgb.MGB_D)
    assert math.sqrt(s.sumsq(top, mgb.MGB_D)) == pytest.approx(2394.206159330749, rel=1e-11)
    s.set_dirichlet(top, mgb.MGB_U)
    before = math.sqrt(s.sumsq(top, mgb.MGB_U))
    assert s.residual(top) == pytest.approx(2513804842.7661724, rel=1e-10)
    assert s.vcycle() == pytest.approx(309102136.21424252, rel=1e-10)
    assert s.vcycle() == pytest.approx(37472190.538855247, rel=1e-10)
    # Dirichlet faces never written: ||u||^2 grew only by interior values
    assert math.sqrt(s.sumsq(top, mgb.MGB_U)) > before
    for lvl in (0, 1, 2):
        a = s.download(lvl, mgb.MGB_D)
        assert not a[0].any() and not a[:, -1].any() and not a[:, :, 0].any()
    s.close()


def test_linearity_of_cycle_operator(mgb):
    """the V-cycle error propagator is linear: with zero rhs and zero boundary
    data, cycle(a*u) == a*cycle(u) for a power of two (exact in fp64)"""
    from oracle_lib import seeded
    with mgb.Solver(3, 6, 2) as s:
        top = s.levels - 1
        u0 = seeded(s.dims(top), 77)
        u0[0] = u0[-1] = 0; u0[:, 0] = u0[:, -1] = 0; u0[:, :, 0] = u0[:, :, -1] = 0
        s.zero(top, mgb.MGB_D)
        s.upload(top, mgb.MGB_U, u0)
        s.vcycle()
        a = s.download(top, mgb.MGB_U)
        s.upload(top, mgb.MGB_U, 4.0 * u0)
        s.vcycle()
        b = s.download(top, mgb.MGB_U)
        assert np.array_equal(b, 4.0 * a)


def test_fixed_point_is_idempotent(mgb):
    """the analytic solution x^2-2y^2+z^2 is reproduced exactly by the 7-point
    stencil: a half-sweep / residual on it changes nothing beyond rounding"""
    with mgb.Solver(3, 7, 2) as s:
        top = s.levels - 1
        N = s.dims(top)[0]
        h = s.spacing(top)
        x = np.arange(N) * h
        exact = (x * x)[:, None, None] - 2 * (x * x)[None, :, None] + (x * x)[None, None, :]
        s.upload(top, mgb.MGB_U, exact)
        s.zero(top, mgb.MGB_D)
        assert s.residual(top) < 1e-6 * math.sqrt(N ** 3) / h ** 2 * 1e-9
        s.smooth(top, 2, True)
        assert np.max(np.abs(s.download(top, mgb.MGB_U) - exact)) < 1e-14


def test_rb_gs_flow_257_matches_reference(mgb):
    """BASELINE config 2 (test_rb_gs_3d.c:56-101 flow, d == 0, 257^3): residual
    norms after iterations 1..5 and 10 as produced by the reference (SURVEY.md
    Appendix A, one thread; 8 threads agree to 1e-13)"""
    want = {0: 39429263.432111837, 1: 26202730.587062154, 2: 18704743.16078392,
            3: 14819928.628106939, 4: 12409130.770850345, 5: 10750238.124142876,
            10: 6718454.1814461211}
    with mgb.Solver(257, 1, 1) as s:
        s.set_dirichlet(0, mgb.MGB_U)
        assert s.residual(0) == pytest.approx(want[0], rel=1e-11)
        for it in range(1, 11):
            s.smooth(0, 1, True)    # preSmoother(v, d, N, h, 1)
            s.smooth(0, 1, False)   # postSmoother(v, d, N, h, 1)
            if it in want:
                assert s.residual(0) == pytest.approx(want[it], rel=1e-11), it
