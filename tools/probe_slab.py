"""Times a half-sweep over a SLAB of planes of one level on one GPU (what a rank of the
partitioned solver runs): tuning of the chunking for short slabs.  MGB_TILE_* /
MGB_MARCH_* environment variables select the launch plan."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_parallel_b200 as m  # noqa: E402
from multigrid_parallel_b200._lib import check  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--levels", type=int, default=9)
ap.add_argument("--planes", type=int, default=64)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
s = m.Solver(3, a.levels, 2)
top = a.levels - 1
N = s.dims(top)[0]
s.set_dirichlet(top, m.MGB_U)
lo = 1
hi = min(lo + a.planes, N - 1)


def run():
    check(s.L.mgb_debug_half_sweep_range(s.h_, top, 1, lo, hi))
    check(s.L.mgb_debug_half_sweep_range(s.h_, top, 0, lo, hi))


for _ in range(3):
    run()
s.sync()
s.timer_start()
for _ in range(a.reps):
    run()
t = s.timer_stop() / (2 * a.reps)
dof = float(hi - lo) * N * N
print(f"N={N} planes={hi - lo}: {t * 1e6:8.1f} us per half-sweep, {12 * dof / t / 1e9:7.0f} GB/s "
      f"(ideal at 6500 GB/s: {12 * dof / 6.5e12 * 1e6:6.1f} us)  env "
      + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("MGB_")), flush=True)
