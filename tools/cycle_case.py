"""A short, fixed sequence of V-cycles for profilers (ncu launch lists / --set full
captures): `--cycles` graph-replayed cycles of the test_mg_3d problem after
`--warm` warm-up cycles, then (optionally) a few stand-alone coarse solves of a
config-5 coarse grid.  Prints the last residual norm; not a benchmark."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_parallel_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--levels", type=int, default=9)
ap.add_argument("--coarse", type=int, nargs=3, default=[3, 3, 3])
ap.add_argument("--warm", type=int, default=3)
ap.add_argument("--cycles", type=int, default=2)
ap.add_argument("--graph", type=int, default=1)
ap.add_argument("--batch", type=int, default=0, help="1: the timed cycles are enqueued back to back (mgb_vcycles)")
ap.add_argument("--lu", type=int, default=0, help="N: also time coarse solves of the (2N+1)x9x9 grid")
a = ap.parse_args()
with m.Solver(tuple(a.coarse), a.levels, 2) as s:
    top = s.levels - 1
    s.set_option(m.solver.OPT_GRAPH, a.graph)
    s.set_dirichlet(top, m.MGB_D)
    s.set_dirichlet(top, m.MGB_U)
    for _ in range(a.warm):
        s.vcycle()
    s.sync()
    s.timer_start()
    if a.batch:
        r = s.vcycles(a.cycles)
    else:
        for _ in range(a.cycles):
            r = s.vcycle()
    dt = s.timer_stop()
    print(f"grid {s.dims(top)} {1e3 * dt / a.cycles:.3f} ms/cycle{" (batched)" if a.batch else ""}, residual {r!r}")
if a.lu:
    with m.Solver((2 * a.lu + 1, 9, 9), 2, 2) as s:
        n, bw, fsec = s.coarse_info()
        for _ in range(3):
            s.coarse_solve()
        s.sync()
        s.timer_start()
        for _ in range(20):
            s.coarse_solve()
        print(f"coarse n={n} bw={bw}: factor {fsec * 1e3:.2f} ms, solve {s.timer_stop() / 20 * 1e6:.1f} us")
