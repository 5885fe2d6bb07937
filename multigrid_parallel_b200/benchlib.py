"""Measurement legs of bench.py, for one GPU and for N GPUs (one process per GPU).

Every leg goes through the C ABI (libmgb.so via the ctypes mirror); times are
CUDA events on the solver's own stream (mgb_timer_start/stop), max over ranks.
No CPU checker and no reference code is involved here (bench.py times those itself).

  main_problem   the headline V-cycle problem (513^3 per GPU)
  stage_roofline per-stage CUDA-event times of the finest level (1 GPU) or the
                 half-sweep incl. its halo exchange per GPU (N GPUs)
  e2e_solve      whole solve from pinned host buffers and back
  rbgs           BASELINE config 2: RB-GS sweeps alone, 257^3 and 513^3
  strong1025     BASELINE config 4: the 1025^3 cube on N GPUs (+ 1 GPU for the ratio)
  config5        BASELINE config 5: 1025 x 1025 x (256N+1), coarse 9 x 9 x (2N+1) dense LU
"""
import json
import math
import os
import time

from . import dist as D
from .solver import MGB_D, MGB_U, OPT_PROFILE, STAGE_NAMES, Solver

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOL = 1e-8                       # test_mg_3d.c:19
HALF_SWEEP_BYTES_PER_DOF = 12.0  # read 1/2 v, read 1/2 d, write 1/2 v
# SURVEY 8(d) algorithmic bytes per DOF of the other finest-level stages:
# stage -> (kernel as named in profiles/traffic.json, what it is, bytes per DOF)
STAGE_BYTES_PER_DOF = {
    "CalcResidual1": ("k_tile<-1,1,2,11,34>", "residual+restrict (TMA tile kernel)", 17.0),
    # inside the cycle only the red points are corrected (the post-smoother's first
    # half-sweep overwrites the black ones): 1 + 4 + 4 B/DOF
    "Prolongate&Correct": ("k_tile_prolong_one<1>", "prolongation+correction of the red points "
                           "(TMA ring)", 9.0),
    "CalcResidual2": ("k_tile<-1,0,2,5,43>", "residual norm (TMA tile kernel)", 16.0),
}


def measured_peak_gbs():
    try:
        return (float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]),
                "measured (MEASURED_PEAKS.json)")
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the
    committed `ncu --set full` capture (profiles/traffic.json), or None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def fresh_problem(s):
    """the test_mg_3d.c problem on the finest level: u = d = 0 inside, BCFunc on the faces"""
    top = s.levels - 1
    s.zero(top, MGB_U)
    s.zero(top, MGB_D)
    s.set_dirichlet(top, MGB_D)
    s.set_dirichlet(top, MGB_U)
    return math.sqrt(s.sumsq(top, MGB_D))


def time_cycles(s, steps, warmup):
    """`steps` V-cycles after `warmup`; device seconds (max over ranks), launches"""
    for _ in range(warmup):
        s.vcycle()
    s.sync()
    l0 = s.launch_count
    D.barrier()
    s.sync()
    s.timer_start()
    for _ in range(steps):
        s.vcycle()
    dt = D.max_over_ranks(s.timer_stop())
    D.barrier()
    return dt, s.launch_count - l0


def stage_roofline_single(s, prof_cycles):
    """1 GPU: the same cycles with per-stage CUDA events (eager launches); the smoother
    stages of the finest level are 2*gs half-sweep launches each"""
    peak, peak_src = measured_peak_gbs()
    top = s.levels - 1
    ni, nj, nk = s.dims(top)
    dof = float(ni) * nj * nk
    s.set_option(OPT_PROFILE, 1)
    s.vcycle()
    s.timing_reset()
    for _ in range(prof_cycles):
        s.vcycle()
    stage = {st: s.timing(top, st)[1] / prof_cycles for st in range(7)}
    s.set_option(OPT_PROFILE, 0)
    n_half = 2 * s.gs * 2  # launches per cycle on the finest level (pre + post)
    t_half = (stage[0] + stage[5]) / n_half
    achieved = HALF_SWEEP_BYTES_PER_DOF * dof / t_half / 1e9
    share = (stage[0] + stage[5]) / sum(stage.values())
    others = {}
    for st in range(7):
        name = STAGE_NAMES[st]
        if name in STAGE_BYTES_PER_DOF and stage[st] > 0:
            kern, what, bpd = STAGE_BYTES_PER_DOF[name]
            gbs = bpd * dof / stage[st] / 1e9
            others[kern] = {"what": what, "achieved": gbs, "frac": gbs / peak, "bytes_per_dof": bpd,
                            "launch_us": stage[st] * 1e6, "traffic": ncu_traffic(kern),
                            "traffic_source": "static ncu capture (profiles/traffic.json)"}
    roofline = {"bound": "hbm", "kernel": "k_tile_sweep<c,6,43> (RB-GS half-sweep, TMA ring)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic("k_tile_sweep<1,6,43>"),
                "traffic_source": "static ncu capture (profiles/traffic.json)",
                "peak_source": peak_src, "bytes_per_dof": HALF_SWEEP_BYTES_PER_DOF,
                "avg_launch_us": t_half * 1e6, "share_of_finest_level": share,
                "frac_of_8TBs_nominal": achieved / 8000.0, "other_kernels": others}
    return roofline, {STAGE_NAMES[st]: stage[st] * 1e6 for st in range(7)}


def stage_roofline_dist(s, rank):
    """N GPUs: the half-sweep on this rank's slab of the finest level, timed with its halo
    exchange as the cycle runs it"""
    peak, peak_src = measured_peak_gbs()
    top = s.levels - 1
    ni, nj, nk = s.dims(top)
    i0, li, own_lo, own_hi = s.local_range(top)
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
    achieved = HALF_SWEEP_BYTES_PER_DOF * local_dof / t_half / 1e9
    return {"bound": "hbm", "kernel": "k_tile_sweep<c,6,43> (RB-GS half-sweep, TMA ring) + its halo "
                                      "exchange over NVLink peer memory",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "peak_source": peak_src, "bytes_per_dof": HALF_SWEEP_BYTES_PER_DOF,
            "avg_launch_us": t_half * 1e6,
            "note": f"per GPU, rank {rank}'s slab of {own_planes} planes; time = slowest rank, "
                    "half-sweep incl. pushing its boundary planes and waiting for the neighbours'"}


def e2e_solve(s, init, reps=3):
    """whole solve through the ABI from pinned host arrays: upload grid + rhs slab,
    V-cycles to 1e-8*||d||, download the grid slab; wall clock, max over ranks"""
    import torch
    top = s.levels - 1
    shape = s.local_shape(top)
    hu = torch.zeros(shape, dtype=torch.float64).pin_memory()
    hd = torch.zeros(shape, dtype=torch.float64).pin_memory()
    fresh_problem(s)
    s.download_ptr(top, MGB_U, hu.data_ptr())
    s.download_ptr(top, MGB_D, hd.data_ptr())
    u0 = hu.clone().pin_memory()
    times, hist = [], [0.0]
    for _ in range(reps):
        hu.copy_(u0)
        D.barrier()
        t0 = time.perf_counter()
        s.upload_ptr(top, MGB_U, hu.data_ptr())
        s.upload_ptr(top, MGB_D, hd.data_ptr())
        hist = s.solve(init * TOL, 100)
        s.download_ptr(top, MGB_U, hu.data_ptr())
        times.append(D.max_over_ranks(time.perf_counter() - t0))
    slab_bytes = float(shape[0]) * shape[1] * shape[2] * 8
    return min(times[1:]), hist, slab_bytes


def rbgs(n, iters, device):
    """BASELINE config 2 (test_rb_gs_3d.c:56-101 flow): preSmoother(.,1) + postSmoother(.,1)
    per iteration on ONE resident grid = 2 full sweeps = 48 B/DOF"""
    peak, _ = measured_peak_gbs()
    with Solver(n, 1, 1, device=device) as s:
        dof = float(n) ** 3
        s.set_dirichlet(0, MGB_U)
        for _ in range(5):
            s.smooth(0, 1, True)
            s.smooth(0, 1, False)
        s.sync()
        s.timer_start()
        for _ in range(iters):
            s.smooth(0, 1, True)
            s.smooth(0, 1, False)
        sec = s.timer_stop()
        res = s.residual(0)
    gbs = 24.0 * dof * 2 * iters / sec / 1e9
    return {"grid": f"{n}^3", "full_sweeps": 2 * iters, "us_per_full_sweep": sec / (2 * iters) * 1e6,
            "gbs": gbs, "frac_of_measured_peak": gbs / peak, "frac_of_8TBs_nominal": gbs / 8000.0,
            "bytes_per_dof_per_full_sweep": 24.0, "residual_after": res}


def gs_lex(n, device, reps=10):
    """GaussSeidelSmoother's lexicographic sweep (mg_3d.h:546-637,
    test_gs_3d.c:56 flow: one sweep per call on a resident grid) as a tile wavefront"""
    with Solver(n, 1, 1, device=device) as s:
        s.set_dirichlet(0, MGB_U)
        for _ in range(10):  # also brings the clocks up after the host-bound legs before it
            s.gs_lex(0, 1)
        s.sync()
        s.timer_start()
        for _ in range(reps):
            s.gs_lex(0, 1)
        sec = s.timer_stop() / reps
    return {"grid": f"{n}^3", "ms_per_sweep": 1e3 * sec, "dof_per_s": float(n) ** 3 / sec,
            "note": "bit-identical to the reference's serial (i,j,k) loop; shared-memory tiles "
                    "of 16x16x32 points, hyperplanes inside a tile, tiles ordered by flags; "
                    "bound by tile latency along the dependence chain, not by HBM"}


def fmg_cycles(device):
    """cycles to 1e-8*||d|| at 513^3 with and without SolverFMGInitialize
    (mg_3d.h:1364-1404 as written: it amounts to one V-cycle ahead)"""
    out = {}
    for use in (False, True):
        with Solver(3, 9, 2, device=device) as s:
            init = fresh_problem(s)
            n = 0
            r = None
            if use:
                r = s.fmg_init()
            while (r is None or r > 1e-8 * init) and n < 60:
                r = s.vcycle()
                n += 1
            out["with_fmg_init" if use else "plain"] = {"cycles_to_1e-8": n, "final_residual": r}
    return out


def vtk_stream(n, device):
    """writeOutputData's file for an n^3 grid produced on the GPU and
    streamed to the host (mgb_vtk_*): seconds and bytes; the chunks are only counted here
    (a C caller fwrite()s them), so this is the rate the library delivers text at"""
    import ctypes as C

    import numpy as np

    from ._lib import check, load_library
    L = load_library()
    rng = np.random.default_rng(0)
    v = rng.uniform(-1e-6, 1e-6, (n, n, n))
    nxt = L.mgb_vtk_next
    nxt.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong)]
    t0 = time.perf_counter()
    w = C.c_void_p()
    check(L.mgb_vtk_open(C.byref(w), v.ctypes.data_as(C.POINTER(C.c_double)), n, n, n,
                         1.0 / (n - 1), device))
    nbytes = 0
    p, cnt, hc = C.c_void_p(), C.c_longlong(), C.c_longlong()
    while True:
        check(nxt(w, C.byref(p), C.byref(cnt)))
        if cnt.value == 0:
            break
        nbytes += cnt.value
    check(L.mgb_vtk_host_chunks(w, C.byref(hc)))
    check(L.mgb_vtk_close(w))
    sec = time.perf_counter() - t0
    return {"grid": f"{n}^3", "bytes": nbytes, "seconds": sec, "gb_per_s": nbytes / sec / 1e9,
            "chunks_formatted_by_host_snprintf": hc.value,
            "note": "text produced on the device, byte-identical to postprocess.h:5-47; open "
                    "(pinned buffers, tables) + H2D of the values from pageable memory + D2H of "
                    "the text + close; the reference's fprintf loop writes ~50 MB/s"}


def _cycle_ms(coarse, levels, gs, steps, warmup, solve=True, device=None, single=False):
    """ms per V-cycle of a fresh test_mg_3d-type problem on the current ranks (or, with
    single=True, on this process's GPU alone)"""
    s = Solver(coarse, levels, gs, device=device) if single else D.make_solver(coarse, levels, gs)
    try:
        init = fresh_problem(s)
        if single:
            for _ in range(warmup):
                s.vcycle()
            s.sync()
            s.timer_start()
            for _ in range(steps):
                s.vcycle()
            dt = s.timer_stop()
        else:
            dt, _ = time_cycles(s, steps, warmup)
        out = {"ms_per_cycle": 1e3 * dt / steps, "first_partitioned_level":
               (None if single or s.nranks == 1 else s.first_dist_level)}
        ni, nj, nk = s.dims(levels - 1)
        out["grid"] = f"{ni}x{nj}x{nk}"
        out["dof_cycles_per_s"] = float(ni) * nj * nk * steps / dt
        if solve:
            fresh_problem(s)
            hist = s.solve(init * TOL, 100)
            out["cycles_to_1e-8"] = len(hist)
            out["final_residual"] = float(hist[-1])
        n, bw, fsec = s.coarse_info()
        out["coarse_unknowns"], out["coarse_half_bandwidth"] = n, bw
        out["lu_factor_ms"] = fsec * 1e3
        # the coarsest solve alone (one launch per cycle, on the rank(s) that own level 0)
        for _ in range(3):
            s.coarse_solve()
        s.sync()
        s.timer_start()
        for _ in range(20):
            s.coarse_solve()
        out["lu_solve_us"] = s.timer_stop() / 20 * 1e6
        return out
    finally:
        s.close()


def strong1025(world, rank, local_rank, steps=10, warmup=3):
    """BASELINE config 4: the SAME 1025^3 cube (3 10 2) on `world` GPUs; the 1-GPU time the
    speed-up is quoted against is measured in the same run (on rank 0's GPU)"""
    out = _cycle_ms((3, 3, 3), 10, 2, steps, warmup)
    out["workload"] = "test_mg_3d 3 10 2: 1025^3 fp64 Laplace V(2,2)-cycle, strong scaling"
    out["n_gpus"] = world
    if world > 1:
        one = None
        if rank == 0:
            one = _cycle_ms((3, 3, 3), 10, 2, 5, 2, solve=True, device=local_rank, single=True)
        D.barrier()
        one = D.broadcast_bytes(one, 0)
        out["one_gpu_ms_per_cycle"] = one["ms_per_cycle"]
        out["one_gpu_final_residual"] = one.get("final_residual")
        out["speedup_vs_1gpu"] = one["ms_per_cycle"] / out["ms_per_cycle"]
        out["target_speedup_at_8"] = 6.0
    return out


def config5(world, steps=10, warmup=3):
    """BASELINE config 5: weak scaling 1025 x 1025 x (256*N+1) per N GPUs (the slab axis is
    the reference's slowest index i, so the extents read (256N+1) x 1025 x 1025), 8 levels,
    coarsest grid (2N+1) x 9 x 9 = 81(2N+1) unknowns solved by the dense-LU path of
    gauss_elim.h every cycle"""
    out = _cycle_ms((2 * world + 1, 9, 9), 8, 2, steps, warmup)
    out["workload"] = (f"{256 * world + 1}x1025x1025 fp64 Laplace V(2,2)-cycle, coarse "
                       f"{2 * world + 1}x9x9 LU ({81 * (2 * world + 1)} unknowns), 8 levels")
    out["n_gpus"] = world
    out["lu_solve_share_of_cycle"] = out["lu_solve_us"] * 1e-3 / out["ms_per_cycle"]
    return out


def dropin_e2e(device):
    """BASELINE config 3 as the reference's user runs it: the UNMODIFIED test_mg_3d.c
    compiled against the drop-in headers (compat/_build/test_mg_3d_gpu, built where the
    reference's sources are mounted), `3 9 2`, default settings; the ASCII VTK (8 GB) goes
    to /dev/null.  Returns the driver's own `Overall time for solving` (its timed region,
    test_mg_3d.c:36-68) and the wall time of the whole program."""
    import re
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "multigrid_parallel_b200", "compat", "_build", "test_mg_3d_gpu")
    if not os.path.exists(exe):
        return {"unavailable": "compat/_build/test_mg_3d_gpu not prebuilt (needs the reference's "
                               "test_mg_3d.c at build time)"}
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    with tempfile.TemporaryDirectory() as tmp:
        os.symlink("/dev/null", os.path.join(tmp, "diff2.vtk"))
        env = dict(os.environ, OMP_NUM_THREADS=str(threads), MGB_VTK_TIMING="1",
                   MGB_DEVICE=str(device))
        t0 = time.perf_counter()
        p = subprocess.run([exe, "3", "9", "2"], cwd=tmp, env=env, capture_output=True, text=True,
                           timeout=900)
        wall = time.perf_counter() - t0
    if p.returncode != 0:
        return {"error": (p.stdout[-300:] + p.stderr[-300:])}
    secs = float(re.search(r"^Overall time for solving:\s*(\S+)", p.stdout, re.M).group(1))
    cycles = len(re.findall(r"Residual Norm:", p.stdout))
    err = re.search(r"^Error norm:\s*(\S+)", p.stdout, re.M).group(1)
    vtk = re.search(r"writeOutputData .*: 513\^3 points, formatted on the GPU \((\d+) chunks by the "
                    r"host's snprintf\), ([0-9.]+) s", p.stderr)
    dof = 513.0 ** 3
    return {"program": "the reference's unmodified test_mg_3d.c on compat/mg_3d.h, args 3 9 2, "
                       f"{threads} OpenMP threads, MGB_PROFILE=1 (per-stage timing, eager launches), "
                       "lazy page-protection coherence",
            "overall_time_for_solving_s": secs, "cycles": cycles,
            "value": dof * cycles / secs, "unit": "DOF*cycles/s",
            "error_norm_printed": err, "whole_program_wall_s": wall,
            "vtk_formatting_s": float(vtk.group(2)) if vtk else None,
            "vtk_chunks_formatted_by_host_snprintf": int(vtk.group(1)) if vtk else None,
            "vtk_note": "ASCII legacy VTK of 513^3 points (8 GB): text produced on the GPU "
                        "(mgb_vtk_*, exact decimal conversion), streamed through pinned buffers, "
                        "byte-identical to the reference's writer, written to /dev/null; the "
                        "reference's fprintf loop takes ~2.5 min for the same file"}
