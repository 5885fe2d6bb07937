"""Multi-GPU parity: the slab-partitioned solver against the single-GPU solver,
bit for bit, both running the CUDA kernels of libmgb (no CPU code involved).

Every rank builds its slab of a partitioned hierarchy (mgb_create_dist) and, on
the same GPU, the whole single-GPU hierarchy (mgb_create); after every V-cycle
the rank's planes of every partitioned level must equal the corresponding
planes of the single-GPU arrays BIT FOR BIT, and the norms agree to 1e-13
(their summation order over the ranks differs).  Used by tests/dist_check.py
(under torchrun) and by bench.py --gpus N, which refuses to print a number for
a partitioned run that is not bitwise equal.

The slab decomposition mirrors the reference's `omp for schedule(static)` over i
(mg_3d.h:658,681,729,753,807,962,1006): the result must not depend on it.
"""
import math

import numpy as np

from . import dist as D
from .solver import MGB_D, MGB_U, Solver


def seeded(shape, seed):
    """uniform(-1,1) float64 array from a fixed seed (PCG64)"""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.uniform(-1.0, 1.0, size=shape)


def bits_equal(a, b):
    """equal as bit patterns (np.array_equal would call -0. and +0. equal)"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))


def feasible(coarse, levels, world, min_planes=2, min_points=0):
    """can this hierarchy be cut into `world` slabs at all (mgb_create_dist's own rule)?"""
    from ._lib import load_library
    mp = min_planes if min_planes > 0 else 16
    pts = min_points if min_points >= 0 else (1 << 20)
    return load_library().mgb_plan_first_dist_level(*coarse, levels, world, mp, pts) < levels


def check_case(coarse, levels, gs, min_planes=2, min_points=0, cycles=2, random_rhs=True,
               check_halos=True):
    """collective; returns a dict describing the case; raises AssertionError on
    the first difference"""
    rank, world, local_rank = D.env_ranks()
    s = D.make_solver(coarse, levels, gs, min_planes=min_planes, min_points=min_points)
    one = Solver(coarse, levels, gs, device=local_rank)
    try:
        top = levels - 1
        i0, li, own_lo, own_hi = s.local_range(top)
        assert s.first_dist_level >= 1
        if random_rhs:
            shape = one.dims(top)
            u0, d0 = seeded(shape, 91), seeded(shape, 92)
            one.upload(top, MGB_U, u0)
            one.upload(top, MGB_D, d0)
            s.upload(top, MGB_U, u0[i0:i0 + li])
            s.upload(top, MGB_D, d0[i0:i0 + li])
        else:
            for solver in (one, s):
                solver.set_dirichlet(top, MGB_D)
                solver.set_dirichlet(top, MGB_U)
        n_one = math.sqrt(one.sumsq(top, MGB_D))
        n_s = math.sqrt(s.sumsq(top, MGB_D))
        assert abs(n_one - n_s) <= 1e-13 * n_one, (n_one, n_s)
        r_one, r_s = one.residual(top), s.residual(top)
        assert abs(r_one - r_s) <= 1e-13 * r_one, (r_one, r_s)
        norm_dev = 0.0
        for c in range(cycles):
            a, b = one.vcycle(), s.vcycle()
            bad = None  # differences are agreed on by all ranks before anybody raises:
            #             a rank that left early would leave the others in a halo wait
            if not abs(a - b) <= 1e-13 * a:
                bad = f"cycle {c} rank {rank}: norm {a!r} vs {b!r}"
            norm_dev = max(norm_dev, abs(a - b) / a)
            full = one.download(top, MGB_U)
            mine = s.download(top, MGB_U)
            # owned planes, and the nearest halo plane on each side (the halo planes are
            # only defined between collective calls: nobody starts the next cycle
            # before everybody has looked -- barrier below)
            halo = 1 if check_halos else 0
            lo = own_lo - (halo if rank > 0 else 0)
            hi = own_hi + (halo if rank < world - 1 else 0)
            if not bits_equal(mine[lo - i0:hi - i0], full[lo:hi]):
                bad = bad or f"cycle {c} rank {rank}: finest u differs"
            for lvl in range(s.first_dist_level, top):
                j0, lj, olo, ohi = s.local_range(lvl)
                for which, name in ((MGB_U, "u"), (MGB_D, "d")):
                    f = one.download(lvl, which)
                    m_ = s.download(lvl, which)
                    if not bits_equal(m_[olo - j0:ohi - j0], f[olo:ohi]):
                        bad = bad or f"cycle {c} rank {rank}: level {lvl} {name} differs"
            if D.max_over_ranks(1.0 if bad else 0.0) != 0.0:  # also the barrier
                raise AssertionError(bad or "another rank's slab differs")
        e_one, e_s = one.error_sumsq(), s.error_sumsq()
        assert abs(e_one - e_s) <= 1e-12 * max(e_one, 1e-300), (e_one, e_s)
        ni, nj, nk = one.dims(top)
        return {"grid": f"{ni}x{nj}x{nk}", "coarse": list(coarse), "levels": levels, "gs": gs,
                "first_partitioned_level": s.first_dist_level, "cycles": cycles,
                "rhs": "random" if random_rhs else "dirichlet", "norm_max_rel_dev": norm_dev}
    finally:
        s.close()
        one.close()


def bench_cases(world):
    """what bench.py --gpus N checks before it times anything: one cube, one
    weak-scaling box (coarse (2N+1) x 3 x 3) and one config-5 box (coarse
    (2N+1) x 9 x 9: the dense-LU coarse grid), default kernel choice, the first
    cycle eager, the others graph replays"""
    return [
        dict(coarse=(3, 3, 3), levels=6, gs=2, cycles=3, random_rhs=False),
        dict(coarse=(2 * world + 1, 3, 3), levels=6, gs=2, cycles=3, random_rhs=True),
        dict(coarse=(2 * world + 1, 9, 9), levels=3, gs=2, cycles=3, random_rhs=True),
    ]


def run_bench_parity(world):
    """collective; {"bitwise": bool, "cases": [...]} -- never raises"""
    rank = D.env_ranks()[0]
    out, ok = [], True
    for case in bench_cases(world):
        if not feasible(case["coarse"], case["levels"], world):
            continue  # this grid cannot be cut into `world` slabs
        try:
            out.append(check_case(**case))
        except Exception as e:  # AssertionError or a libmgb failure
            ok = False
            out.append({"coarse": list(case["coarse"]), "levels": case["levels"],
                        "failed": f"rank {rank}: {type(e).__name__}: {e}"[:300]})
    # one rank's failure is everybody's
    ok = D.max_over_ranks(0.0 if ok else 1.0) == 0.0
    return {"bitwise": ok, "against": "the single-GPU solver on the same GPUs (all levels, every "
                                      "cycle; norms to 1e-13)", "cases": out}
