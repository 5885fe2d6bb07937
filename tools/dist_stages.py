"""Per-level, per-stage device times of the partitioned V-cycle (CUDA events, eager
launches, MGB_OPT_PROFILE) on rank 0 and the last rank, next to the graph-replayed
cycle time.  Developer tool; run under torchrun:

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29520 tools/dist_stages.py [--problem weak|strong1025|config5]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multigrid_parallel_b200 as m  # noqa: E402
from multigrid_parallel_b200 import benchlib as B  # noqa: E402
from multigrid_parallel_b200 import dist as D  # noqa: E402
from multigrid_parallel_b200.solver import OPT_PROFILE  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--problem", default="weak")
a = ap.parse_args()
rank, world, local_rank = D.init_process_group("gloo")
coarse, levels = {"weak": ((2 * world + 1, 3, 3), 9), "strong1025": ((3, 3, 3), 10),
                  "config5": ((2 * world + 1, 9, 9), 8)}[a.problem]
if world == 1 and a.problem == "weak":
    coarse = (3, 3, 3)
s = D.make_solver(coarse, levels, 2)
B.fresh_problem(s)
dt, _ = B.time_cycles(s, 10, 3)
if rank == 0:
    print(f"{a.problem} {s.dims(levels - 1)} on {world} GPU(s): {1e3 * dt / 10:.3f} ms/cycle (graph), "
          f"first partitioned level {s.first_dist_level if world > 1 else None}", flush=True)
s.set_option(OPT_PROFILE, 1)
s.vcycle()
s.timing_reset()
for _ in range(5):
    s.vcycle()
D.barrier()
for r in (0, world - 1):
    if rank == r:
        print(f"[rank {rank}] us/cycle: " + " ".join(f"{n[:9]:>9s}" for n in m.STAGE_NAMES))
        for lvl in range(levels - 1, -1, -1):
            row = " ".join(f"{s.timing(lvl, st)[1] / 5 * 1e6:9.1f}" for st in range(7))
            print(f"[rank {rank}] L{lvl} {s.dims(lvl)}: {row}", flush=True)
    D.barrier()
s.close()
