"""Times the tile kernels (tile.cu) one by one at one level size (CUDA events on
the solver's stream).  Developer tool; tile shapes come from MGB_TILE_* env."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_parallel_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--levels", type=int, default=9)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--which", default="norm,rr,k3,k1")
a = ap.parse_args()
s = m.Solver(3, a.levels, 2)
top = a.levels - 1
N = s.dims(top)[0]
dof = float(N) ** 3
s.set_dirichlet(top, m.MGB_D)
s.set_dirichlet(top, m.MGB_U)
peak = 6553.6
rows = {
    "norm": ("residual norm", lambda: s.L.mgb_residual(s.h_, top, 0, None), 16),
    "rr": ("residual+restrict", lambda: s.residual_restrict(top), 17),
    "k3": ("sweep+residual norm", lambda: s.L.mgb_sweep_residual(s.h_, top, 1, None), 16),
    "k1": ("sweep+residual+restrict", lambda: s.sweep_residual_restrict(top, 0), 17),
    "hs": ("half sweep", lambda: s.half_sweep(top, 1), 12),
    "pc": ("prolong+correct", lambda: s.prolong_correct(top), 17),
}
for key in a.which.split(","):
    name, fn, bpd = rows[key]
    fn()
    s.sync()
    s.timer_start()
    for _ in range(a.reps):
        fn()
    sec = s.timer_stop() / a.reps
    gbs = bpd * dof / sec / 1e9
    print(f"N={N} {name:26s} {sec*1e6:9.1f} us {gbs:8.1f} GB/s alg. {gbs/peak*100:5.1f}% of measured peak", flush=True)
