"""Times the lexicographic Gauss-Seidel wavefront (mgb_gs_lex, csrc/gslex.cu) on one
resident grid: ms per call for 1, 2 and 4 pipelined sweeps.  GaussSeidelSmoother flow of
test_gs_3d.c:56 (rhs 0, Dirichlet faces)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_parallel_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[65, 129, 257, 513])
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
out = {}
for n in a.n:
    with m.Solver((n, n, n), 1, 1) as s:
        s.set_dirichlet(0, m.MGB_U)
        row = {}
        for iters in (1, 2, 4):
            s.gs_lex(0, iters)
            s.sync()
            s.timer_start()
            for _ in range(a.reps):
                s.gs_lex(0, iters)
            dt = s.timer_stop() / a.reps
            row[f"ms_{iters}_sweeps"] = 1e3 * dt
            row[f"us_per_step_{iters}"] = 1e6 * dt / (3 * n - 8 + 2 * (iters - 1))
        out[str(n)] = row
        print(n, json.dumps(row), flush=True)
