"""Generate tests/golden/*.json from the REFERENCE ITSELF (oracle/_ref, built
from /root/reference by oracle/Makefile).  Run in the build container:

    python oracle/gen_golden.py [--big] [--huge]

histories.json : per-cycle residual norms (test_mg_3d.c flow, tol 1e-8), cycle
counts, error norms and checksums of the final solution for the configurations
of SURVEY.md Appendix A.  Everything is run with ONE OpenMP thread, so the
norms are bitwise reproducible by the serial oracle (the solution is
thread-count invariant anyway; the reference's per-thread sequential norm sums
are not: 8 threads differ from 1 thread by 1e-11 at 513^3 and 6e-10 at 1025^3,
cf. SURVEY.md Appendix A).  --big adds 257^3 and 513^3, --huge 1025^3 (30 GB,
~17 min).
operators.json : checksums of every operator's output on seeded inputs.
"""
import argparse
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import Ref, seeded  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def checksums(u, h):
    N = u.shape[0]
    x = np.arange(N) * h
    exact = (x * x)[:, None, None] - 2 * (x * x)[None, :, None] + (x * x)[None, None, :]
    diff = u - exact
    p = np.arange(u.size, dtype=np.uint64)
    w = ((p * np.uint64(2654435761)) % np.uint64(1000)).astype(np.float64) * 1e-3
    return {
        "errnorm_np": float(np.sqrt((diff * diff).sum())),
        "max_abs_err": float(np.abs(diff).max()),
        "sumsq_np": float((u * u).sum()),
        "wsum_np": float((u.reshape(-1) * w).sum()),
        "probe_1_2_3": float(u[1, 2, 3]),
        "sha256": hashlib.sha256(u.tobytes()).hexdigest(),
    }


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true")
    ap.add_argument("--huge", action="store_true")
    a = ap.parse_args()
    ref = Ref()
    assert ref.available, "build oracle/_ref first (make -C oracle)"
    os.makedirs(GOLD, exist_ok=True)
    path = os.path.join(GOLD, "histories.json")
    out = json.load(open(path)) if os.path.exists(path) else {}

    cases = [(3, 5, 2), (3, 5, 1), (3, 5, 3), (5, 4, 2), (9, 3, 2), (3, 6, 2), (3, 7, 2)]
    big = [(3, 8, 2), (3, 9, 2)] if a.big else []
    huge = [(3, 10, 2)] if a.huge else []
    for coarse, levels, gs in cases + big + huge:
        key = f"{coarse}_{levels}_{gs}"
        threads = 1
        ref.set_threads(threads)
        hist, init, u, secs = ref.solve(coarse, levels, gs, tol=1e-8, max_cycles=60)
        N = u.shape[0]
        entry = {"coarse": coarse, "levels": levels, "gs": gs, "N": N, "threads": threads,
                 "tol": 1e-8, "init_norm": init, "cycles": len(hist),
                 "history": [float(x) for x in hist], "seconds": secs}
        entry.update(checksums(u, 1.0 / (N - 1)))
        out.setdefault(key, {}).update(entry)  # keeps history_exact (gen_exact_history.py)
        print(key, "N", N, "cycles", len(hist), "last", hist[-1], "err", entry["errnorm_np"],
              f"{secs:.2f}s", flush=True)
        del u
        json.dump(out, open(path, "w"), indent=1)
    ref.set_threads(1)

    # FMG initialisation (SolverFMGInitialize, mg_3d.h:1364-1404, restated on the live
    # functions by oracle/ref_harness.c:ref_fmg_init): every level right after it, and
    # the solve that follows
    fmg = {}
    for coarse, levels, gs in [(3, 5, 2), (3, 7, 2), (5, 4, 1)]:
        ref.set_threads(1)
        state = ref.fmg_state(coarse, levels, gs)
        hist, init, u, _ = ref.solve(coarse, levels, gs, tol=1e-8, max_cycles=60, fmg=True)
        fmg[f"{coarse}_{levels}_{gs}"] = {
            "coarse": coarse, "levels": levels, "gs": gs,
            "u_sha256": [sha(uu) for uu, dd in state], "d_sha256": [sha(dd) for uu, dd in state],
            "init_norm": init, "cycles_after_fmg": len(hist), "history": [float(x) for x in hist],
            "solution_sha256": sha(u)}
        print("fmg", coarse, levels, gs, "cycles after", len(hist), flush=True)
    json.dump(fmg, open(os.path.join(GOLD, "fmg.json"), "w"), indent=1)

    # operator-level fixtures on seeded inputs
    ops = {}
    N, Nc = 17, 9
    h = 1.0 / (N - 1)
    v, d = seeded((N,) * 3, 1), seeded((N,) * 3, 2)
    ops["inputs"] = {"v": sha(v), "d": sha(d), "generator": "numpy PCG64 uniform(-1,1), seeds 1,2"}
    w = v.copy(); ref.smooth(w, d, h, 2, True); ops["pre_smooth_2"] = sha(w)
    ref.smooth(w, d, h, 3, False); ops["then_post_smooth_3"] = sha(w)
    r = np.zeros_like(v); n = ref.residual(w, d, h, r)
    ops["residual"] = sha(r); ops["residual_norm"] = n
    dc = np.zeros((Nc,) * 3); ref.restrict(r, dc); ops["restrict"] = sha(dc)
    ec = seeded((Nc,) * 3, 9); ef = seeded((N,) * 3, 11); ref.prolong_correct(ec, ef)
    ops["prolong_correct"] = sha(ef)
    A = ref.coarse_matrix(5, 0.25); ops["coarse_matrix_5"] = sha(A)
    ref.lu_factor(A); ops["lu_5"] = sha(A)
    b = seeded((125,), 3); ops["lu_solve_5"] = sha(ref.lu_solve(A, b))
    # GaussSeidelSmoother (mg_3d.h:546-637: lexicographic sweeps + updateEdgeValues), the
    # reference's own symbol, on the same seeded inputs
    w = v.copy(); ref.gs_lex(w, d, h, 2); ops["gauss_seidel_smoother_2"] = sha(w)
    ref.gs_lex(w, d, h, 1); ops["then_gauss_seidel_smoother_1"] = sha(w)
    # writeOutputData (postprocess.h:5-47), the reference's own writer, on a seeded 9^3 grid
    import tempfile
    g = seeded((9,) * 3, 21) * 10.0 ** np.random.default_rng(21).integers(-12, 3, (9,) * 3)
    with tempfile.TemporaryDirectory() as tmp:
        ref.write_vtk(os.path.join(tmp, "g.vtk"), g, 0.125)
        ops["vtk_9_sha256"] = hashlib.sha256(open(os.path.join(tmp, "g.vtk"), "rb").read()).hexdigest()
    json.dump(ops, open(os.path.join(GOLD, "operators.json"), "w"), indent=1)
    print("operators.json written")


if __name__ == "__main__":
    main()
