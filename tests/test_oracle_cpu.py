"""CPU: pins the oracle.  (1) the repo's restatement (oracle/mg_oracle.c)
against the golden vectors produced by the reference itself
(oracle/gen_golden.py -> tests/golden), (2) against the reference compiled
into oracle/_ref, bit for bit, wherever that build exists, (3) the known-answer
vectors the reference ships (gauss_elim.h:99-124, red_black_gs_scalability.txt)."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

from oracle_lib import OrcMG, seeded

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def histories():
    return json.load(open(os.path.join(GOLD, "histories.json")))


@pytest.mark.parametrize("key", ["3_5_2", "3_5_1", "3_5_3", "5_4_2", "9_3_2", "3_6_2"])
def test_oracle_solve_matches_reference_golden(orc, histories, key):
    g = histories[key]
    mg = OrcMG(orc, g["coarse"], g["levels"], g["gs"])
    hist, init = mg.solve(tol=g["tol"], max_cycles=60)
    assert init == g["init_norm"]
    assert len(hist) == g["cycles"]
    # golden small cases were produced with one thread: bitwise
    assert [float(x) for x in hist] == g["history"]
    u = mg.u(g["levels"] - 1)
    assert sha(u) == g["sha256"]
    assert float(u[1, 2, 3]) == g["probe_1_2_3"]
    mg.close()


def test_appendix_a_values(histories):
    """the headline numbers quoted in SURVEY.md Appendix A / BASELINE.md"""
    assert histories["3_5_2"]["history"][0] == 9178.6027816049173
    assert histories["3_5_2"]["history"][-1] == 3.0990167206114415e-07
    assert histories["3_5_2"]["init_norm"] == 74.870102354427829
    assert [histories[k]["cycles"] for k in ("3_5_2", "3_6_2", "3_7_2", "3_8_2", "3_9_2")] == \
        [14, 15, 15, 16, 16]
    # 513^3 norms depend on the OpenMP team size at the 1e-11 level (each thread
    # sums its slab sequentially); the survey's run and the golden run differ so
    assert histories["3_9_2"]["history"][0] == pytest.approx(38614842.175600566, rel=1e-10)
    assert histories["3_9_2"]["history"][15] == pytest.approx(9.8556680734968831e-06, rel=1e-10)


def test_oracle_operators_match_reference_golden(orc):
    ops = json.load(open(os.path.join(GOLD, "operators.json")))
    N, Nc = 17, 9
    h = 1.0 / (N - 1)
    v, d = seeded((N,) * 3, 1), seeded((N,) * 3, 2)
    assert sha(v) == ops["inputs"]["v"] and sha(d) == ops["inputs"]["d"]
    w = v.copy()
    orc.smooth(w, d, h, 2, True)
    assert sha(w) == ops["pre_smooth_2"]
    orc.smooth(w, d, h, 3, False)
    assert sha(w) == ops["then_post_smooth_3"]
    r = np.zeros_like(v)
    n = orc.residual(w, d, h, r)
    assert sha(r) == ops["residual"] and n == ops["residual_norm"]
    dc = np.zeros((Nc,) * 3)
    orc.restrict(r, dc)
    assert sha(dc) == ops["restrict"]
    ec, ef = seeded((Nc,) * 3, 9), seeded((N,) * 3, 11)
    orc.prolong_correct(ec, ef)
    assert sha(ef) == ops["prolong_correct"]
    A = orc.coarse_matrix((5, 5, 5), 0.25)
    assert sha(A) == ops["coarse_matrix_5"]
    orc.lu_factor(A)
    assert sha(A) == ops["lu_5"]
    assert sha(orc.lu_solve(A, seeded((125,), 3))) == ops["lu_solve_5"]


def test_known_answer_lu(orc):
    # gauss_elim.h:99-124
    a = np.array([[2., -1, 0], [-1, 2, -1], [0, -1, 2]])
    orc.lu_factor(a)
    assert np.allclose(orc.lu_solve(a, np.array([3., 4, 5])), [5.5, 8, 6.5], rtol=0, atol=1e-14)


def test_rbgs_50_known_history(orc):
    """red_black_gs_scalability.txt flow at N=50 (shipped code: final norm
    0.28128984869198337 after 1303 iterations) -- checked on a prefix: the
    residual ratio of (pre+post) iterations settles at sqrt(0.983675)"""
    N = 50
    h = 1.0 / (N - 1)
    u, d = np.zeros((N,) * 3), np.zeros((N,) * 3)
    orc.set_dirichlet(u, h)
    prev = orc.residual(u, d, h)
    for it in range(60):
        orc.smooth(u, d, h, 1, True)
        orc.smooth(u, d, h, 1, False)
        cur = orc.residual(u, d, h)
        ratio, prev = cur / prev, cur
    assert 0.98 < ratio < 0.995


def test_oracle_invariants(orc):
    """r and coarse boundaries stay 0; non-cubic boxes reduce to
    the cubic code path"""
    mg = OrcMG(orc, 3, 4, 2)
    mg.solve(tol=1e-8, max_cycles=3)
    for lvl in range(3):
        for arr in (mg.u(lvl), mg.d(lvl)):
            assert not arr[0].any() and not arr[-1].any() and not arr[:, 0].any()
            assert not arr[:, :, -1].any()
    r = mg.r(3)
    assert not r[0].any() and not r[:, 0].any() and not r[:, :, 0].any()
    mg.close()


@pytest.mark.parametrize("threads", [1, 3])
def test_oracle_vs_compiled_reference(orc, ref, threads):
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    ref.set_threads(threads)
    N, Nc = 17, 9
    h = 1.0 / (N - 1)
    v, d = seeded((N,) * 3, 1), seeded((N,) * 3, 2)
    a, b = v.copy(), v.copy()
    orc.smooth(a, d, h, 2, True); ref.smooth(b, d, h, 2, True)
    assert np.array_equal(a, b)
    orc.smooth(a, d, h, 1, False); ref.smooth(b, d, h, 1, False)
    assert np.array_equal(a, b)
    ra, rb = np.zeros_like(v), np.zeros_like(v)
    na, nb = orc.residual(a, d, h, ra), ref.residual(b, d, h, rb)
    assert np.array_equal(ra, rb) and na == pytest.approx(nb, rel=1e-14)
    r = seeded((N,) * 3, 7)
    da, db = seeded((Nc,) * 3, 5), None
    db = da.copy()
    orc.restrict(r, da); ref.restrict(r, db)
    assert np.array_equal(da, db)
    ec = seeded((Nc,) * 3, 9)
    ea = seeded((N,) * 3, 11)
    eb = ea.copy()
    orc.prolong_correct(ec, ea); ref.prolong_correct(ec, eb)
    assert np.array_equal(ea, eb)
    A1, A2 = orc.coarse_matrix((5, 5, 5), 0.25), ref.coarse_matrix(5, 0.25)
    assert np.array_equal(A1, A2)
    orc.lu_factor(A1); ref.lu_factor(A2)
    assert np.array_equal(A1, A2)
    bb = seeded((125,), 3)
    assert np.array_equal(orc.lu_solve(A1, bb), ref.lu_solve(A2, bb))
    ref.set_threads(1)


def test_oracle_full_solve_vs_compiled_reference(orc, ref):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    ref.set_threads(1)
    hist, init, u, _ = ref.solve(5, 3, 2)
    mg = OrcMG(orc, 5, 3, 2)
    h2, i2 = mg.solve()
    assert init == i2 and np.array_equal(hist, h2) and np.array_equal(u, mg.u(2))
    mg.close()
