#!/usr/bin/env python
"""bench.py -- headline benchmark: V-cycle DOF/s at 513^3 (test_mg_3d `3 9 2`).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one V(2,2)-cycle of the fp64 Laplace problem of test_mg_3d.c
(Dirichlet x^2-2y^2+z^2 on the unit cube, coarse grid 3^3, 9 levels -> 513^3).
`value` = DOF*cycles/s with everything resident in HBM, CUDA-event timed on the
solver's stream; `e2e` = the same metric for the whole reference-facing solve
(host grid + rhs in pinned memory -> upload -> V-cycles to 1e-8*||d|| ->
download).  `roofline` is the RB-GS half-sweep kernel (12 B/DOF algorithmic)
against the measured HBM copy peak; `cpu_baseline` is the reference itself
(oracle/_ref, built from /root/reference) on this box's host cores.

--impl reference times the reference's own OpenMP implementation on the host.
Nothing here reads /root/reference at run time.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

COARSE, LEVELS, GS = 3, 9, 2          # test_mg_3d 3 9 2 -> 513^3
TOL = 1e-8                            # test_mg_3d.c:19
METRIC = "vcycle_dof_per_s"
UNIT = "DOF*cycles/s"
HALF_SWEEP_BYTES_PER_DOF = 12.0       # read 1/2 v, read 1/2 d, write 1/2 v
# SURVEY 8(d) algorithmic bytes per DOF of the other finest-level stages
# stage -> (kernel as named in profiles/traffic.json, what it is, bytes per DOF)
STAGE_BYTES_PER_DOF = {"CalcResidual1": ("k_tile<-1,1,2,11,34>", "residual+restrict (TMA tile kernel)", 17.0),
                       # inside the cycle only the red points are corrected (the post-smoother's
                       # first half-sweep overwrites the black ones): 1 + 4 + 4 B/DOF
                       "Prolongate&Correct": ("k_tile_prolong_one<1>", "prolongation+correction of the red "
                                              "points (TMA ring) + face fix-up kernel", 9.0),
                       "CalcResidual2": ("k_tile<-1,0,2,5,43>", "residual norm (TMA tile kernel)", 16.0)}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` from the
    committed `ncu --set full` summary (profiles/traffic.json), or None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[kernel]["dram_bytes_per_launch"]
    except Exception:
        return None


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock + throttle reasons while the timed region runs"""

    def __init__(self, index=0, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def finish(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        if not self.samples:
            return self._smi()
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}

    def _smi(self):
        try:
            out = subprocess.run(
                ["nvidia-smi", f"--id={self.index}",
                 "--query-gpu=clocks.sm,clocks.max.sm,clocks_event_reasons.active",
                 "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10)
            a, b, c = [x.strip() for x in out.stdout.strip().split(",")[:3]]
            return {"sm_mhz": int(a), "sm_max_mhz": int(b), "reasons": [c], "samples": 1,
                    "note": "single idle nvidia-smi sample (NVML sampling unavailable)"}
        except Exception as e:  # pragma: no cover
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "note": str(e)}


# --------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own OpenMP code on host cores
# --------------------------------------------------------------------------
def cpu_reference_cycles(coarse, levels, gs, warm, timed):
    """`warm`+`timed` V-cycles of the reference (oracle/_ref/libmg_ref.so);
    returns (seconds for the timed cycles, threads, N)"""
    import ctypes as C

    from oracle_lib import Ref, c_dp
    ref = Ref()
    if not ref.available:
        raise RuntimeError("oracle/_ref/libmg_ref.so missing (built by __graft_entry__.build() "
                           "where /root/reference exists)")
    L = ref.L
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1,
    # which would silently turn the OpenMP reference into a serial run)
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    ref.set_threads(ncores)
    grid, rhs, h = c_dp(), c_dp(), C.c_double()
    N = L.ref_solver_open(coarse, levels, gs, C.byref(grid), C.byref(rhs), C.byref(h))
    # test_mg_3d.c:17-29 set-up
    L.SolverSetupBoundaryConditions()
    L.setupBoundaryConditions.argtypes = [c_dp, C.c_int, C.c_double]
    L.setupBoundaryConditions(grid, N, h.value)
    for _ in range(warm):
        L.ref_vcycle()
    t0 = time.perf_counter()
    for _ in range(timed):
        L.ref_vcycle()
    dt = time.perf_counter() - t0
    threads = L.ref_max_threads()
    L.ref_solver_close()
    return dt, threads, N


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ.setdefault("OMP_PLACES", "cores")
    dt, threads, N = cpu_reference_cycles(COARSE, LEVELS, GS, args.warmup, args.steps)
    dof = float(N) ** 3
    val = dof * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"test_mg_3d {COARSE} {LEVELS} {GS}: {N}^3 fp64 Laplace V(2,2)-cycle, "
                               "reference OpenMP code on host cores"
                               + ("" if args.gpus == 1 else f" (the reference only has cubes: this is one "
                                  f"rank's 513^3 share of the {args.gpus}-GPU arm's (512*N+1)x513x513 box; "
                                  "the metric is per DOF)")},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "reference",
                         "sample": f"{args.steps} V-cycles of the {N}^3 problem after {args.warmup} warm-up"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def run_ours(args):
    import numpy as np

    import multigrid_parallel_b200 as m
    from multigrid_parallel_b200.solver import OPT_PROFILE

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if world > 1 or args.problem != "weak":
        from multigrid_parallel_b200 import dist_bench
        return dist_bench.run(args, rank, world, local_rank)

    peak, peak_src = measured_peak_gbs()
    s = m.Solver(COARSE, LEVELS, GS, device=local_rank)
    top = s.levels - 1
    N = s.dims(top)[2]
    dof = float(N) ** 3

    def fresh_problem():
        s.zero(top, m.MGB_U)
        s.zero(top, m.MGB_D)
        s.set_dirichlet(top, m.MGB_D)
        s.set_dirichlet(top, m.MGB_U)

    fresh_problem()
    init = math.sqrt(s.sumsq(top, m.MGB_D))
    for _ in range(max(args.warmup, 3)):
        s.vcycle()
    s.sync()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = s.launch_count
    s.timer_start()
    for _ in range(args.steps):
        s.vcycle()
    dt = s.timer_stop()
    launches = s.launch_count - l0
    clocks = sampler.finish()
    value = dof * args.steps / dt

    # dominant kernel, live: the same cycles with per-stage CUDA events; the
    # smoother stages of the finest level are 2*gs half-sweep launches each
    s.set_option(OPT_PROFILE, 1)
    s.vcycle()
    s.timing_reset()
    prof_cycles = max(3, min(args.steps, 10))
    for _ in range(prof_cycles):
        s.vcycle()
    stage = {st: s.timing(top, st)[1] / prof_cycles for st in range(7)}
    s.set_option(OPT_PROFILE, 0)
    n_half = 2 * GS * 2  # launches per cycle on the finest level (pre + post)
    t_half = (stage[0] + stage[5]) / n_half
    achieved = HALF_SWEEP_BYTES_PER_DOF * dof / t_half / 1e9
    share = (stage[0] + stage[5]) / (sum(stage[st] for st in range(7)))
    others = {}
    for st in range(7):
        name = m.STAGE_NAMES[st]
        if name in STAGE_BYTES_PER_DOF and stage[st] > 0:
            kern, what, bpd = STAGE_BYTES_PER_DOF[name]
            gbs = bpd * dof / stage[st] / 1e9
            others[kern] = {"what": what, "achieved": gbs, "frac": gbs / peak, "bytes_per_dof": bpd,
                            "launch_us": stage[st] * 1e6, "traffic": ncu_traffic(kern)}

    # end to end through the C ABI with host buffers (pinned), whole solve
    import torch
    hu = torch.zeros((N, N, N), dtype=torch.float64).pin_memory()
    hd = torch.zeros((N, N, N), dtype=torch.float64).pin_memory()
    # host copies of the problem (faces = BCFunc, interior 0), produced once
    fresh_problem()
    s.download_ptr(top, m.MGB_U, hu.data_ptr())
    s.download_ptr(top, m.MGB_D, hd.data_ptr())
    u0 = hu.clone().pin_memory()
    thr = init * TOL
    e2e_times, cycles = [], 0
    for rep in range(3):
        hu.copy_(u0)
        t0 = time.perf_counter()
        s.upload_ptr(top, m.MGB_U, hu.data_ptr())
        s.upload_ptr(top, m.MGB_D, hd.data_ptr())
        hist = s.solve(thr, 100)
        s.download_ptr(top, m.MGB_U, hu.data_ptr())
        e2e_times.append(time.perf_counter() - t0)
        cycles = len(hist)
    t_e2e = min(e2e_times[1:])
    e2e_val = dof * cycles / t_e2e

    cpu = None
    if not args.no_cpu_baseline:
        try:
            cdt, threads, _ = cpu_reference_cycles(COARSE, LEVELS, GS, 1, 3)
            cpu = {"value": dof * 3 / cdt, "unit": UNIT, "cores": threads, "kind": "reference",
                   "sample": f"3 V-cycles of the same {N}^3 problem after 1 warm-up, reference OpenMP "
                             "code (oracle/_ref) on this box's host cores"}
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                   "sample": f"unavailable: {e}"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"test_mg_3d {COARSE} {LEVELS} {GS}: {N}^3 fp64 Laplace V(2,2)-cycle, "
                               "Dirichlet x^2-2y^2+z^2, coarse 3^3 LU",
                   "l2": "inputs larger than L2 (3 x 1.1 GB level arrays vs 126 MB)",
                   "cycles_to_1e-8": cycles, "final_residual": float(hist[-1])},
        "clocks": clocks,
        "gpu_launches": int(launches),
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(2 * dof * 8),
                "d2h_bytes_per_step": int(dof * 8), "seconds_per_solve": t_e2e, "cycles": cycles,
                "step": "one full solve: upload grid+rhs from pinned host memory, V-cycles to "
                        "1e-8*||d||, download grid"},
        "roofline": {"bound": "hbm", "kernel": "k_tile_sweep (RB-GS half-sweep, TMA ring)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic("k_tile_sweep<1,6,43>"),
                     "peak_source": peak_src, "bytes_per_dof": HALF_SWEEP_BYTES_PER_DOF,
                     "avg_launch_us": t_half * 1e6, "share_of_finest_level": share,
                     "frac_of_8TBs_nominal": achieved / 8000.0,
                     "other_kernels": others},
        "stage_us_finest": {m.STAGE_NAMES[st]: stage[st] * 1e6 for st in range(7)},
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    s.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--problem", default="weak", choices=["weak", "strong1025"],
                    help="weak (default): 513^3 per GPU; strong1025: the 1025^3 cube of BASELINE "
                         "config 4 on --gpus GPUs (extra mode, no e2e leg)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
