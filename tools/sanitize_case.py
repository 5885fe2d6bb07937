"""Small V-cycle + operator calls with the TMA tile kernels forced onto every
level: the workload for compute-sanitizer (memcheck / racecheck) runs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_parallel_b200 as m  # noqa: E402

m.set_global(m.G_TILE, 1)
m.set_global(m.G_TILE_MIN_PLANE, 0)
for coarse, levels in (((3, 3, 3), 6), ((3, 5, 9), 4)):
    s = m.Solver(coarse, levels, 2)
    top = levels - 1
    s.set_dirichlet(top, m.MGB_D)
    s.set_dirichlet(top, m.MGB_U)
    for fuse in (1, 2):
        s.set_option(2, fuse)
        print(coarse, levels, "fuse", fuse, [s.vcycle() for _ in range(2)], flush=True)
    s.half_sweep(top, 1)
    s.residual_restrict(top)
    s.prolong_correct(top)
    print("norm", s.residual(top), s.sweep_residual(top, 0), flush=True)
    s.close()
print("SANITIZE_CASE_OK")
