"""Graph-replayed cycle time of a partitioned problem for several partitioning thresholds
(min planes per rank / min points per rank below which a level is replicated instead of
partitioned).  Developer tool; run under torchrun."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multigrid_parallel_b200 import benchlib as B  # noqa: E402
from multigrid_parallel_b200 import dist as D  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--problem", default="weak")
ap.add_argument("--cases", default="16:-1,32:-1,64:-1")
a = ap.parse_args()
rank, world, local_rank = D.init_process_group("gloo")
coarse, levels = {"weak": ((2 * world + 1, 3, 3), 9), "strong1025": ((3, 3, 3), 10),
                  "config5": ((2 * world + 1, 9, 9), 8)}[a.problem]
for case in a.cases.split(","):
    mp, pts = (int(x) for x in case.split(":"))
    s = D.make_solver(coarse, levels, 2, min_planes=mp, min_points=pts)
    B.fresh_problem(s)
    dt, _ = B.time_cycles(s, 20, 5)
    if rank == 0:
        print(f"{a.problem} {s.dims(levels - 1)} on {world} GPU(s), min_planes {mp} min_points/rank {pts}: "
              f"{1e3 * dt / 20:.3f} ms/cycle, first partitioned level {s.first_dist_level}", flush=True)
    s.close()
    D.barrier()
