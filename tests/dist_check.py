"""Run under torchrun with one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
      --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank builds its slab of a partitioned hierarchy and, on the same GPU, the
whole single-GPU hierarchy; after each stage the slab (owned planes AND halos)
must equal the corresponding planes of the single-GPU arrays BIT FOR BIT, and
the norms must agree to 1e-13."""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import multigrid_parallel_b200 as m  # noqa: E402
from multigrid_parallel_b200 import dist as D  # noqa: E402
from oracle_lib import seeded  # noqa: E402


def check_case(coarse, levels, gs, min_planes, min_points, cycles, random_rhs, rank, world,
               local_rank):
    s = D.make_solver(coarse, levels, gs, min_planes=min_planes, min_points=min_points)
    one = m.Solver(coarse, levels, gs, device=local_rank)
    top = levels - 1
    i0, li, own_lo, own_hi = s.local_range(top)
    assert s.first_dist_level >= 1
    if random_rhs:
        shape = one.dims(top)
        u0, d0 = seeded(shape, 91), seeded(shape, 92)
        one.upload(top, m.MGB_U, u0)
        one.upload(top, m.MGB_D, d0)
        s.upload(top, m.MGB_U, u0[i0:i0 + li])
        s.upload(top, m.MGB_D, d0[i0:i0 + li])
    else:
        for solver in (one, s):
            solver.set_dirichlet(top, m.MGB_D)
            solver.set_dirichlet(top, m.MGB_U)
    n_one = math.sqrt(one.sumsq(top, m.MGB_D))
    n_s = math.sqrt(s.sumsq(top, m.MGB_D))
    assert abs(n_one - n_s) <= 1e-13 * n_one, (n_one, n_s)
    r_one, r_s = one.residual(top), s.residual(top)
    assert abs(r_one - r_s) <= 1e-13 * r_one, (r_one, r_s)
    for c in range(cycles):
        a, b = one.vcycle(), s.vcycle()
        assert abs(a - b) <= 1e-13 * a, (c, a, b)
        full = one.download(top, m.MGB_U)
        mine = s.download(top, m.MGB_U)
        lo = own_lo - (1 if rank > 0 else 0)      # owned planes + nearest halos
        hi = own_hi + (1 if rank < world - 1 else 0)
        assert np.array_equal(mine[lo - i0:hi - i0], full[lo:hi]), f"cycle {c} rank {rank}"
        # every partitioned level, owned planes
        for lvl in range(s.first_dist_level, top):
            j0, lj, olo, ohi = s.local_range(lvl)
            fu = one.download(lvl, m.MGB_U)
            mu = s.download(lvl, m.MGB_U)
            assert np.array_equal(mu[olo - j0:ohi - j0], fu[olo:ohi]), (c, lvl, rank)
            fd = one.download(lvl, m.MGB_D)
            md = s.download(lvl, m.MGB_D)
            assert np.array_equal(md[olo - j0:ohi - j0], fd[olo:ohi]), (c, lvl, rank, "d")
    e_one, e_s = one.error_sumsq(), s.error_sumsq()
    assert abs(e_one - e_s) <= 1e-12 * max(e_one, 1e-300), (e_one, e_s)
    info = (coarse, levels, gs, "LD", s.first_dist_level, "own", own_lo, own_hi)
    s.close()
    one.close()
    return info


def main():
    rank, world, local_rank = D.init_process_group("gloo")
    assert world > 1, "run under torchrun with >= 2 processes"
    cases = [
        # coarse, levels, gs, min_planes, min_points, cycles, random rhs
        ((3, 3, 3), 6, 2, 2, 0, 3, False),       # cube 65^3, partitioned down to tiny levels
        ((3, 3, 3), 6, 2, 16, 0, 2, True),       # same, agglomerated below 33 planes/rank
        ((2 * world + 1, 3, 3), 6, 1, 2, 0, 2, True),   # weak-scaling box
        ((5, 3, 5), 5, 3, 4, 0, 2, True),
        ((3, 3, 3), 8, 2, 16, -1, 2, False),     # 257^3 with the default thresholds
    ]
    for min_plane in (40000, 0):  # default kernel choice, then TMA tile kernels on every level
        m.set_global(m.G_TILE_MIN_PLANE, min_plane)
        for case in cases:
            if (case[0][0] - 1) * 2 % world:
                continue
            info = check_case(*case, rank, world, local_rank)
            D.barrier()
            if rank == 0:
                print("ok", info, "tile_min_plane", min_plane, flush=True)
    print(f"DIST_CHECK_OK rank {rank}/{world}", flush=True)


if __name__ == "__main__":
    main()
