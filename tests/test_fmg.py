"""FMG initialisation: SolverFMGInitialize of the reference
(mg_3d.h:1364-1404, commented out upstream; restated on the live functions by
oracle/ref_harness.c:ref_fmg_init, whose outputs are tests/golden/fmg.json).

CPU: the oracle's restatement against the goldens and against the compiled
reference.  GPU: mgb_fmg_init against the goldens (cubes) and the oracle
(boxes), every level's u and d bit for bit, then the solve that follows."""
import hashlib
import json
import math
import os

import numpy as np
import pytest

from oracle_lib import OrcMG

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "fmg.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("key", sorted(GOLD))
def test_oracle_fmg_matches_reference_golden(orc, key):
    g = GOLD[key]
    mg = OrcMG(orc, g["coarse"], g["levels"], g["gs"])
    init = mg.setup_problem()
    assert init == g["init_norm"]
    mg.fmg_init()
    for l in range(g["levels"]):
        assert sha(mg.u(l)) == g["u_sha256"][l], f"u level {l}"
        assert sha(mg.d(l)) == g["d_sha256"][l], f"d level {l}"
    # the V-cycle loop that follows (test_mg_3d.c:40-66)
    hist = []
    while (not hist or hist[-1] > init * 1e-8) and len(hist) < 60:
        hist.append(mg.vcycle())
    assert len(hist) == g["cycles_after_fmg"]
    assert np.array_equal(np.array(hist), np.array(g["history"]))
    assert sha(mg.u(g["levels"] - 1)) == g["solution_sha256"]
    mg.close()


def test_fmg_as_written_is_one_vcycle_ahead():
    """what the statement order amounts to: every cycle entered below the
    finest level starts by zeroing its entry level and finds a zero right-hand side,
    so the initialisation ends up as boundary values + ONE finest-level V-cycle"""
    hist = json.load(open(os.path.join(HERE, "golden", "histories.json")))
    for key, g in GOLD.items():
        if key not in hist:
            continue
        plain = hist[key]
        assert g["cycles_after_fmg"] == plain["cycles"] - 1
        assert np.allclose(g["history"], plain["history"][1:], rtol=1e-12, atol=0)


@pytest.mark.parametrize("threads", [1, 3])
def test_oracle_fmg_vs_compiled_reference(orc, ref, threads):
    if ref is None:
        pytest.skip("oracle/_ref not built")
    ref.set_threads(threads)
    state = ref.fmg_state(3, 4, 2)
    mg = OrcMG(orc, 3, 4, 2)
    mg.setup_problem()
    mg.fmg_init()
    for l, (u, d) in enumerate(state):
        assert np.array_equal(bits(u), bits(mg.u(l))) and np.array_equal(bits(d), bits(mg.d(l)))
    ref.set_threads(1)
    mg.close()


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(GOLD))
@pytest.mark.parametrize("tail", [1, 0])
def test_gpu_fmg_matches_reference_golden(mgb, key, tail):
    from multigrid_parallel_b200.solver import OPT_TAIL
    g = GOLD[key]
    with mgb.Solver(g["coarse"], g["levels"], g["gs"]) as s:
        s.set_option(OPT_TAIL, tail)
        top = s.levels - 1
        s.set_dirichlet(top, mgb.MGB_D)
        init = math.sqrt(s.sumsq(top, mgb.MGB_D))
        s.set_dirichlet(top, mgb.MGB_U)
        assert init == pytest.approx(g["init_norm"], rel=1e-13)
        first = s.fmg_init()
        for l in range(g["levels"]):
            assert sha(s.download(l, mgb.MGB_U)) == g["u_sha256"][l], f"u level {l}"
            assert sha(s.download(l, mgb.MGB_D)) == g["d_sha256"][l], f"d level {l}"
        hist = s.solve(g["init_norm"] * 1e-8, 60)
        assert len(hist) == g["cycles_after_fmg"]
        # the reference's printed norms are sequential sums over N^3 squares: they differ
        # from the exactly rounded sum by 5e-14 at 33^3 and 5e-12 at 129^3 (DESIGN.md
        # section 3); the solution below is compared bit for bit
        rtol = 1e-12 if g["levels"] <= 5 else 2e-11
        assert np.allclose(hist, g["history"], rtol=rtol, atol=0)
        assert sha(s.download(top, mgb.MGB_U)) == g["solution_sha256"]
        assert first > hist[0]


@pytest.mark.gpu
@pytest.mark.parametrize("coarse,levels,gs", [((3, 5, 9), 4, 2), ((5, 3, 3), 5, 3)])
def test_gpu_fmg_on_boxes_matches_oracle(mgb, orc, coarse, levels, gs):
    mg = OrcMG(orc, coarse, levels, gs)
    mg.setup_problem()
    mg.fmg_init()
    with mgb.Solver(coarse, levels, gs) as s:
        top = levels - 1
        s.set_dirichlet(top, mgb.MGB_D)
        s.set_dirichlet(top, mgb.MGB_U)
        s.fmg_init()
        for l in range(levels):
            assert np.array_equal(bits(s.download(l, mgb.MGB_U)), bits(mg.u(l))), f"u {l}"
            assert np.array_equal(bits(s.download(l, mgb.MGB_D)), bits(mg.d(l))), f"d {l}"
        for _ in range(2):  # and the cycles after it run on consistent coarse levels
            a, b = s.vcycle(), mg.vcycle()
            assert a == pytest.approx(b, rel=1e-13)
        assert np.array_equal(bits(s.download(top, mgb.MGB_U)), bits(mg.u(top)))
    mg.close()
