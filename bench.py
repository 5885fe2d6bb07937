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

The line also carries the other BASELINE configs as objects: `rbgs` (config 2:
RB-GS sweeps alone at 257^3 / 513^3), `strong1025` (config 4), `config5` (1025 x
1025 x (256N+1) with the dense-LU coarse grid 9 x 9 x (2N+1)), and for N > 1
`parity`: the partitioned solver against the single-GPU one, bit for bit, checked
BEFORE anything is timed -- the run fails (rc 1) if it is not bitwise.

--impl reference times the reference's own OpenMP implementation on the host
(all host cores, in a fresh process so libgomp sees the thread settings).
Nothing here reads /root/reference at run time.
"""
import os

# Read BEFORE anything can load libgomp: with OMP_PROC_BIND set, libgomp pins the
# main thread to one place when it is loaded, after which the affinity mask says
# "1 CPU" (that is how round 1's reference arm came to run on one thread).


def _host_cpu_set():
    try:
        cpus = set(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return set(range(os.cpu_count() or 1))
    if len(cpus) > 1 or not os.environ.get("OMP_PROC_BIND"):
        return cpus
    # one CPU AND binding requested: a libgomp loaded earlier (by an import before this
    # module, or by the launcher we inherited the mask from) has most likely pinned this
    # thread.  Ask the processes that were not pinned, and the cgroup.
    for pid in (os.getppid(), 1):
        try:
            other = set(os.sched_getaffinity(pid))
            if len(other) > len(cpus):
                cpus = other
        except OSError:
            pass
    if len(cpus) == 1:
        try:
            txt = open("/sys/fs/cgroup/cpuset.cpus.effective").read().strip()
            eff = set()
            for p in txt.split(","):
                if "-" in p:
                    eff.update(range(int(p.split("-")[0]), int(p.split("-")[1]) + 1))
                elif p:
                    eff.add(int(p))
            if len(eff) > 1:
                cpus = eff
        except (OSError, ValueError):
            pass
    return cpus


HOST_CPU_SET = _host_cpu_set()
HOST_CPUS = len(HOST_CPU_SET)

import argparse  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import sys  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

COARSE, LEVELS, GS = 3, 9, 2          # test_mg_3d 3 9 2 -> 513^3
METRIC = "vcycle_dof_per_s"
UNIT = "DOF*cycles/s"


def workload_config(n_gpus):
    """`config` of the JSON line -- the SAME dict in both arms (ours / --impl reference)"""
    if n_gpus == 1:
        return {"workload": f"test_mg_3d {COARSE} {LEVELS} {GS}: 513^3 fp64 Laplace V(2,2)-cycle, "
                            "Dirichlet x^2-2y^2+z^2, coarse 3^3 LU, stopping rule 1e-8*||d||",
                "parallelism": "single GPU",
                "l2": "inputs larger than L2 (3 x 1.1 GB level arrays vs 126 MB)"}
    return {"workload": f"{512 * n_gpus + 1}x513x513 fp64 Laplace V(2,2)-cycle (the 513^3 problem of "
                        f"test_mg_3d 3 9 2 stretched along i: 512 planes of 513^2 per GPU), Dirichlet "
                        f"x^2-2y^2+z^2, coarse {2 * n_gpus + 1}x3x3 LU, {LEVELS} levels, stopping rule "
                        "1e-8*||d||",
            "parallelism": f"i-slabs over {n_gpus} GPUs (one process per GPU), halo planes stored into "
                           "the neighbours' memory over NVLink by the compute kernels, coarse levels "
                           "below the partitioning threshold replicated on every GPU (P2P all-gather)",
            "l2": "inputs larger than L2 (3 x 1.1 GB level arrays per GPU vs 126 MB)"}


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
def cpu_worker(argv):
    """child process: `warm` + `timed` V-cycles of the reference (oracle/_ref/libmg_ref.so,
    the unmodified mg_3d.h behind oracle/ref_harness.c), the timed ones inside the
    driver's own timed region (test_mg_3d.c:36-68: one parallel region around the loop,
    omp_get_wtime on both sides).  Thread count and binding come from the environment the
    parent set BEFORE this process started."""
    import ctypes as C

    from oracle_lib import Ref, c_dp
    coarse, levels, gs, warm, timed, want_threads = [int(x) for x in argv]
    ref = Ref()
    if not ref.available:
        print(json.dumps({"error": "oracle/_ref/libmg_ref.so missing (built by __graft_entry__."
                                   "build() where /root/reference exists)"}))
        return 0
    L = ref.L
    L.ref_timed_cycles.restype = C.c_double
    L.ref_timed_cycles.argtypes = [C.c_int, c_dp]
    threads = L.ref_max_threads()
    if threads != want_threads:
        print(json.dumps({"error": f"OpenMP offers {threads} threads, wanted {want_threads}"}))
        return 0
    grid, rhs, h = c_dp(), c_dp(), C.c_double()
    N = L.ref_solver_open(coarse, levels, gs, C.byref(grid), C.byref(rhs), C.byref(h))
    # test_mg_3d.c:17-29 set-up
    L.SolverSetupBoundaryConditions()
    L.setupBoundaryConditions.argtypes = [c_dp, C.c_int, C.c_double]
    L.setupBoundaryConditions(grid, N, h.value)
    last = C.c_double()
    if warm:
        L.ref_timed_cycles(warm, C.byref(last))
    secs = L.ref_timed_cycles(timed, C.byref(last))
    L.ref_solver_close()
    print(json.dumps({"seconds": secs, "threads": threads, "N": N, "last_norm": last.value}))
    return 0


def cpu_reference_cycles(coarse, levels, gs, warm, timed):
    """(seconds for the timed cycles, threads, N, host cpus) -- in a fresh process with
    OMP_NUM_THREADS = every CPU this process may run on (torchrun exports
    OMP_NUM_THREADS=1, which would silently make the OpenMP reference serial)"""
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(HOST_CPUS)
    env.setdefault("OMP_PROC_BIND", "close")   # CPU baseline plan
    env.setdefault("OMP_PLACES", "cores")
    env["OMP_DYNAMIC"] = "false"
    cmd = [sys.executable, os.path.abspath(__file__), "--cpu-worker", str(coarse), str(levels),
           str(gs), str(warm), str(timed), str(HOST_CPUS)]

    def widen():  # the child must not inherit a mask some libgomp narrowed in this process
        try:
            os.sched_setaffinity(0, HOST_CPU_SET)
        except OSError:
            pass

    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=3600,
                       preexec_fn=widen)
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    if p.returncode != 0 or not line:
        raise RuntimeError(f"cpu worker failed (rc {p.returncode}): {p.stderr[-500:]}")
    r = json.loads(line[-1])
    if "error" in r:
        raise RuntimeError(r["error"])
    if HOST_CPUS > 1 and r["threads"] != HOST_CPUS:
        raise RuntimeError(f"reference ran on {r['threads']} threads, box offers {HOST_CPUS}")
    return r["seconds"], r["threads"], r["N"]


def cpu_baseline_dict(threads, value, sample):
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
            "host_cpus_available": HOST_CPUS, "is_full_node": threads == HOST_CPUS,
            "sample": sample}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    dt, threads, N = cpu_reference_cycles(COARSE, LEVELS, GS, args.warmup, args.steps)
    dof = float(N) ** 3
    val = dof * args.steps / dt
    sample = (f"{args.steps} V-cycles of the {N}^3 problem (test_mg_3d {COARSE} {LEVELS} {GS}) after "
              f"{args.warmup} warm-up, unmodified reference OpenMP code on {threads} host threads, "
              "timed like test_mg_3d.c:36-68")
    if args.gpus > 1:
        sample += (f"; the reference only has cubes: this is one GPU's 513^3 share of the "
                   f"{args.gpus}-GPU workload, and the metric is per DOF")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus),
        "cpu_baseline": cpu_baseline_dict(threads, val, sample),
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def _leg(fn, *a, **kw):
    """an extra leg must never cost the headline line"""
    try:
        return fn(*a, **kw)
    except Exception as e:  # pragma: no cover
        return {"error": f"{type(e).__name__}: {e}"[:400]}


def run_ours(args):
    import multigrid_parallel_b200 as m
    from multigrid_parallel_b200 import benchlib as B
    from multigrid_parallel_b200 import dist as D
    from multigrid_parallel_b200 import dist_parity

    rank, world, local_rank = D.env_ranks()
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with "
                         "python -m torch.distributed.run --nproc-per-node N bench.py --gpus N")
    D.init_process_group("gloo")
    warmup = max(args.warmup, 3)

    if args.problem == "strong1025":  # stand-alone mode of config 4 (also part of the default line)
        out = B.strong1025(world, rank, local_rank, steps=args.steps, warmup=warmup)
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": out["dof_cycles_per_s"], "unit": UNIT,
                              "n_gpus": world, "steps": args.steps, "warmup": warmup,
                              "ms_per_step": out["ms_per_cycle"], "higher_is_better": True,
                              "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                              "data": "synthetic", "config": {"workload": out["workload"]},
                              "strong1025": out}), flush=True)
        return 0

    # ---- N > 1: parity first.  No number is printed for a partitioned run that does
    # not reproduce the single-GPU solver bit for bit.
    parity = None
    if world > 1:
        parity = dist_parity.run_bench_parity(world)
        if not parity["bitwise"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world,
                                  "parity": parity,
                                  "error": "partitioned solver differs from the single-GPU solver"}),
                      flush=True)
            return 1

    # ---- the headline problem
    coarse = (COARSE,) * 3 if world == 1 else (2 * world + 1, COARSE, COARSE)
    s = D.make_solver(coarse, LEVELS, GS)
    top = s.levels - 1
    ni, nj, nk = s.dims(top)
    dof = float(ni) * nj * nk
    init = B.fresh_problem(s)
    sampler = None
    if rank == 0:
        sampler = ClockSampler(local_rank)
        sampler.start()
    dt, launches = B.time_cycles(s, args.steps, warmup)
    clocks = sampler.finish() if sampler else None
    value = dof * args.steps / dt

    stage_us = None
    if world == 1:
        roofline, stage_us = B.stage_roofline_single(s, max(3, min(args.steps, 10)))
    else:
        roofline = B.stage_roofline_dist(s, rank)

    t_e2e, hist, slab_bytes = B.e2e_solve(s, init)
    cycles = len(hist)
    first_dist = s.first_dist_level if world > 1 else None
    s.close()

    # ---- the other BASELINE configs
    rbgs = None
    if world == 1:
        rbgs = {"flow": "test_rb_gs_3d.c:56-101: preSmoother(.,1) + postSmoother(.,1) per iteration "
                        "on one resident grid, rhs 0, Dirichlet faces",
                "257": _leg(B.rbgs, 257, 200, local_rank), "513": _leg(B.rbgs, 513, 40, local_rank),
                "target": ">= 0.70 of the 8 TB/s nominal HBM peak (BASELINE north_star)"}
    strong = None if args.no_extras else _leg(B.strong1025, world, rank, local_rank)
    c5 = None if args.no_extras else _leg(B.config5, world)

    dropin = next_rows = None
    if world == 1 and not args.no_extras:
        dropin = _leg(B.dropin_e2e, local_rank)
        # SURVEY 8(f) rows built after the hot path: measured, not part of the metric
        next_rows = {"f1_fmg": _leg(B.fmg_cycles, local_rank),
                     "f2_vtk_gpu": _leg(B.vtk_stream, 513, local_rank),
                     "f4_gs_lex": {"257": _leg(B.gs_lex, 257, local_rank),
                                   "513": _leg(B.gs_lex, 513, local_rank)}}

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            cdt, threads, _ = cpu_reference_cycles(COARSE, LEVELS, GS, 1, 3)
            cpu = cpu_baseline_dict(threads, 513.0 ** 3 * 3 / cdt,
                                    "3 V-cycles of the 513^3 problem (test_mg_3d 3 9 2) after 1 warm-up, "
                                    "unmodified reference OpenMP code (oracle/_ref) on this box's host "
                                    "threads, timed like test_mg_3d.c:36-68")
        except Exception as e:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference",
                   "sample": f"unavailable: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world),
            "convergence": {"cycles_to_1e-8": cycles, "final_residual": float(hist[-1]),
                            "grid": f"{ni}x{nj}x{nk}", "first_partitioned_level": first_dist},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": dof * cycles / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(2 * slab_bytes * world),
                    "d2h_bytes_per_step": int(slab_bytes * world), "seconds_per_solve": t_e2e,
                    "cycles": cycles,
                    "step": "one full solve through the C ABI: every rank uploads its grid+rhs slab "
                            "from pinned host memory, V-cycles to 1e-8*||d||, every rank downloads "
                            "its grid slab"},
            "roofline": roofline, "cpu_baseline": cpu,
        }
        if stage_us is not None:
            line["stage_us_finest"] = stage_us
        if parity is not None:
            line["parity"] = parity
        if rbgs is not None:
            line["rbgs"] = rbgs
        if strong is not None:
            line["strong1025"] = strong
        if c5 is not None:
            line["config5"] = c5
        if dropin is not None:
            line["e2e_dropin"] = dropin
        if next_rows is not None:
            line["next_rows"] = next_rows
        print(json.dumps(line), flush=True)
    D.barrier()
    return 0


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--cpu-worker":
        return cpu_worker(sys.argv[2:])
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the strong1025 / config5 legs (quick runs, profiling)")
    ap.add_argument("--problem", default="weak", choices=["weak", "strong1025"],
                    help="weak (default): 513^3 per GPU; strong1025: only the 1025^3 cube of "
                         "BASELINE config 4 on --gpus GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
