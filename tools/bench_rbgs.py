"""BASELINE config 2: red-black Gauss-Seidel sweeps alone on one resident grid
(test_rb_gs_3d.c:56-101 flow: per iteration preSmoother(.,1) then
postSmoother(.,1), i.e. two full sweeps = four colour half-sweeps; the rhs is 0,
the Dirichlet data sits on the faces of v).  Prints the residual norms the flow
prints and the HBM bandwidth of the sweeps: 24 B/DOF per full sweep
 over CUDA-event time on the solver's stream.

    python tools/bench_rbgs.py [--n 257] [--iters 100]
"""
import argparse
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_parallel_b200 as m  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=257)
ap.add_argument("--iters", type=int, default=100)
a = ap.parse_args()
peak = 6553.6
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(
        os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
s = m.Solver(a.n, 1, 1)  # one level: the grid itself, no hierarchy
dof = float(a.n) ** 3
s.set_dirichlet(0, m.MGB_U)
init = s.residual(0)
hist = []
for it in range(10):
    s.smooth(0, 1, True)
    s.smooth(0, 1, False)
    hist.append(s.residual(0))
s.sync()
s.timer_start()
for it in range(a.iters):
    s.smooth(0, 1, True)
    s.smooth(0, 1, False)
sec = s.timer_stop()
gbs = 24.0 * dof * 2 * a.iters / sec / 1e9
print(json.dumps({
    "workload": f"test_rb_gs_3d flow, {a.n}^3 fp64, {a.iters} iterations x 2 full sweeps",
    "initial_residual": init, "residual_after_iterations_1_to_10": hist,
    "us_per_full_sweep": sec / (2 * a.iters) * 1e6, "rbgs_gbs": gbs, "frac_of_measured_peak": gbs / peak,
    "frac_of_8TBs_nominal": gbs / 8000.0, "bytes_per_dof_per_full_sweep": 24.0,
    "note": "257^3: the two arrays (272 MB) are close to L2-sized (126 MB); 513^3 is the DRAM-bound case"}))
