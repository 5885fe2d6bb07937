"""bench.py's N>1 arm: weak scaling of the V-cycle over the slab axis.

Workload at N ranks: (512*N+1) x 513 x 513 fp64 Laplace problem (the 513^3
headline problem stretched along i, so every GPU keeps a 512-plane slab of 513^2
points: BASELINE config 5's shape), coarse grid (2N+1) x 3 x 3, 9 levels,
V(2,2).  One process per GPU; halo planes, the norm all-reduce and the
coarse-level gather/broadcast go over NCCL inside libmgb."""
import json
import math
import time

LEVELS, GS, TOL = 9, 2, 1e-8


def run(args, rank, world, local_rank):
    import torch

    import multigrid_parallel_b200 as m
    from multigrid_parallel_b200 import dist as D

    D.init_process_group("gloo")
    strong = getattr(args, "problem", "weak") == "strong1025"
    # weak: 513^3 per GPU (config 5's shape); strong: the SAME 1025^3 cube on any
    # number of GPUs (BASELINE config 4), 10 levels
    coarse = (3, 3, 3) if strong else (2 * world + 1, 3, 3)
    levels = 10 if strong else LEVELS
    s = D.make_solver(coarse, levels, GS)
    top = s.levels - 1
    ni, nj, nk = s.dims(top)
    dof = float(ni) * nj * nk
    i0, li, own_lo, own_hi = s.local_range(top)

    def fresh():
        s.zero(top, m.MGB_U)
        s.zero(top, m.MGB_D)
        s.set_dirichlet(top, m.MGB_D)
        s.set_dirichlet(top, m.MGB_U)

    fresh()
    init = math.sqrt(s.sumsq(top, m.MGB_D))
    for _ in range(max(args.warmup, 3)):
        s.vcycle()
    s.sync()
    sampler = None
    if rank == 0:
        from bench import ClockSampler
        sampler = ClockSampler(local_rank)
        sampler.start()
    l0 = s.launch_count
    D.barrier()
    s.sync()
    s.timer_start()
    for _ in range(args.steps):
        s.vcycle()
    dt_local = s.timer_stop()
    D.barrier()
    dt = D.max_over_ranks(dt_local)
    launches = s.launch_count - l0
    clocks = sampler.finish() if sampler else None
    value = dof * args.steps / dt

    # dominant kernel per GPU: the RB-GS half-sweep on this rank's slab of the finest
    # level, timed with its halo step (push + wait) as the cycle runs it
    import json as _json
    import os as _os
    try:
        peak = float(_json.load(open(_os.path.join(_os.path.dirname(_os.path.dirname(
            _os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    for _ in range(2):
        s.half_sweep(top, 1)
        s.half_sweep(top, 0)
    s.sync()
    D.barrier()
    s.timer_start()
    nrep = 10
    for _ in range(nrep):
        s.half_sweep(top, 1)
        s.half_sweep(top, 0)
    t_half = D.max_over_ranks(s.timer_stop()) / (2 * nrep)
    own_planes = own_hi - own_lo
    local_dof = float(own_planes) * nj * nk
    achieved = 12.0 * local_dof / t_half / 1e9
    roofline = {"bound": "hbm", "kernel": "k_tile_sweep (RB-GS half-sweep, TMA ring) + halo step",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src, "bytes_per_dof": 12.0,
                "avg_launch_us": t_half * 1e6,
                "note": f"per GPU, rank {rank}'s slab of {own_planes} planes; time = slowest rank, "
                        "half-sweep kernel plus its P2P halo push + wait"}

    if strong:
        fresh()
        hist = s.solve(init * TOL, 100)
        import os
        if os.environ.get("MGB_BENCH_STAGES"):
            # per level and stage device times (CUDA events) of rank 0 and the last rank
            from multigrid_parallel_b200.solver import OPT_PROFILE
            s.set_option(OPT_PROFILE, 1)
            s.vcycle()
            s.timing_reset()
            for _ in range(3):
                s.vcycle()
            if rank in (0, world - 1):
                for lvl in range(s.levels - 1, -1, -1):
                    row = " ".join(f"{s.timing(lvl, st)[1] / 3 * 1e6:9.1f}" for st in range(7))
                    print(f"[rank {rank}] L{lvl} us/cycle: {row}", flush=True)
            s.set_option(OPT_PROFILE, 0)
        if rank == 0:
            print(json.dumps({
                "metric": "vcycle_dof_per_s", "value": value, "unit": "DOF*cycles/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"{ni}x{nj}x{nk} fp64 Laplace V(2,2)-cycle, coarse 3^3 LU, "
                                       f"{levels} levels (BASELINE config 4: strong scaling)",
                           "parallelism": f"i-slabs over {world} GPUs, NCCL halo exchange per "
                                          f"half-sweep, levels < {s.first_dist_level} on rank 0"
                                          if world > 1 else "single GPU",
                           "l2": "inputs larger than L2", "cycles_to_1e-8": len(hist),
                           "final_residual": float(hist[-1])},
                "clocks": clocks, "gpu_launches": int(launches), "e2e": None, "roofline": roofline,
                "cpu_baseline": None}), flush=True)
        D.barrier()
        s.close()
        return 0

    # end to end: every rank uploads its slab from pinned host memory, the
    # ranks solve to 1e-8*||d|| together, every rank downloads its slab
    shape = s.local_shape(top)
    hu = torch.zeros(shape, dtype=torch.float64).pin_memory()
    hd = torch.zeros(shape, dtype=torch.float64).pin_memory()
    fresh()
    s.download_ptr(top, m.MGB_U, hu.data_ptr())
    s.download_ptr(top, m.MGB_D, hd.data_ptr())
    u0 = hu.clone().pin_memory()
    times, cycles, hist = [], 0, [0.0]
    for rep in range(3):
        hu.copy_(u0)
        D.barrier()
        t0 = time.perf_counter()
        s.upload_ptr(top, m.MGB_U, hu.data_ptr())
        s.upload_ptr(top, m.MGB_D, hd.data_ptr())
        hist = s.solve(init * TOL, 100)
        s.download_ptr(top, m.MGB_U, hu.data_ptr())
        times.append(D.max_over_ranks(time.perf_counter() - t0))
        cycles = len(hist)
    t_e2e = min(times[1:])
    slab_bytes = float(shape[0]) * shape[1] * shape[2] * 8

    if rank == 0:
        line = {
            "metric": "vcycle_dof_per_s", "value": value, "unit": "DOF*cycles/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"{ni}x{nj}x{nk} fp64 Laplace V(2,2)-cycle (513^3 per-GPU slab "
                                   f"stretched along i), coarse {coarse[0]}x3x3 LU, {levels} levels",
                       "parallelism": f"i-slabs over {world} GPUs, NCCL halo exchange per half-sweep, "
                                      f"levels < {s.first_dist_level} agglomerated on rank 0",
                       "l2": "inputs larger than L2", "cycles_to_1e-8": cycles,
                       "final_residual": float(hist[-1])},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": dof * cycles / t_e2e, "unit": "DOF*cycles/s",
                    "h2d_bytes_per_step": int(2 * slab_bytes * world),
                    "d2h_bytes_per_step": int(slab_bytes * world), "seconds_per_solve": t_e2e,
                    "cycles": cycles,
                    "step": "one full solve: every rank uploads its grid+rhs slab from pinned host "
                            "memory, V-cycles to 1e-8*||d||, every rank downloads its slab"},
            "roofline": roofline, "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    D.barrier()
    s.close()
    return 0
