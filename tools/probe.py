"""Per-kernel timing probe (CUDA events on the solver's stream).  Not the
bench: a developer tool to see where a V-cycle's time goes."""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_parallel_b200 as m  # noqa: E402
from multigrid_parallel_b200.solver import OPT_FUSE, OPT_GRAPH, OPT_PROFILE  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--levels", type=int, default=9)
ap.add_argument("--coarse", type=int, default=3)
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--cycles", type=int, default=5)
a = ap.parse_args()

s = m.Solver(a.coarse, a.levels, 2)
top = a.levels - 1
N = s.dims(top)[0]
dof = float(N) ** 3
s.set_dirichlet(top, m.MGB_D)
s.set_dirichlet(top, m.MGB_U)
peak = 6553.6


def t(fn, reps=a.reps):
    fn()
    s.sync()
    s.timer_start()
    for _ in range(reps):
        fn()
    return s.timer_stop() / reps


rows = [
    ("half_sweep red", lambda: s.half_sweep(top, 1), 12),
    ("half_sweep black", lambda: s.half_sweep(top, 0), 12),
    ("residual norm", lambda: s.L.mgb_residual(s.h_, top, 0, None), 16),
    ("residual store", lambda: s.L.mgb_residual(s.h_, top, 1, None), 24),
    ("restrict", lambda: s.restrict(top), 9),
    ("residual_restrict", lambda: s.residual_restrict(top), 17),
    ("prolong_correct", lambda: s.prolong_correct(top), 17),
]
print(f"N={N}  DOF={dof:.4g}")
for name, fn, bpd in rows:
    sec = t(fn)
    gbs = bpd * dof / sec / 1e9
    print(f"{name:20s} {sec*1e6:10.1f} us  {gbs:8.1f} GB/s algorithmic  {gbs/peak*100:5.1f}% of measured copy peak")

for mode in ("graph", "eager"):
    s.set_option(OPT_GRAPH, mode == "graph")
    sec = t(lambda: s.vcycle(), a.cycles)
    print(f"vcycle[{mode}] {sec*1e3:8.3f} ms  {dof/sec:.4g} DOF/s")
s.set_option(OPT_PROFILE, 1)
s.timing_reset()
for _ in range(a.cycles):
    s.vcycle()
for lvl in range(a.levels - 1, -1, -1):
    line = " ".join(f"{s.timing(lvl, st)[1]/a.cycles*1e6:9.1f}" for st in range(7))
    print(f"L{lvl} us/cycle: {line}")
