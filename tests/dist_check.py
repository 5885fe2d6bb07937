"""Run under torchrun with one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
      --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Every rank builds its slab of a partitioned hierarchy and, on the same GPU, the
whole single-GPU hierarchy; after each cycle the slab (owned planes AND nearest
halos, every partitioned level) must equal the corresponding planes of the
single-GPU arrays BIT FOR BIT, and the norms must agree to 1e-13
(multigrid_parallel_b200/dist_parity.py; bench.py --gpus N runs a subset of the
same checks before it times anything)."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import multigrid_parallel_b200 as m  # noqa: E402
from multigrid_parallel_b200 import dist as D  # noqa: E402
from multigrid_parallel_b200.dist_parity import check_case, feasible  # noqa: E402


def main():
    rank, world, local_rank = D.init_process_group("gloo")
    assert world > 1, "run under torchrun with >= 2 processes"
    cases = [
        # coarse, levels, gs, min_planes, min_points, cycles, random rhs
        ((3, 3, 3), 6, 2, 2, 0, 3, False),       # cube 65^3, partitioned down to tiny levels
        ((3, 3, 3), 6, 2, 16, 0, 2, True),       # same, coarse levels from 33 planes/rank on one GPU
        ((2 * world + 1, 3, 3), 6, 1, 2, 0, 2, True),   # weak-scaling box
        ((2 * world + 1, 9, 9), 3, 2, 2, 0, 2, True),   # config-5 box: dense-LU coarse grid
        ((5, 3, 5), 5, 3, 4, 0, 2, True),
        ((3, 3, 3), 8, 2, 16, -1, 2, False),     # 257^3 with the default thresholds
        ((3, 3, 3), 8, 2, 0, -1, 3, False),      # 257^3, everything default
    ]
    # DIST_CHECK_CASES=0,2 / DIST_CHECK_REPEAT=5: a subset, several times (hunting a race)
    if os.environ.get("DIST_CHECK_CASES"):
        cases = [cases[int(x)] for x in os.environ["DIST_CHECK_CASES"].split(",")]
    cases = cases * int(os.environ.get("DIST_CHECK_REPEAT", "1"))
    for min_plane in (40000, 0):  # default kernel choice, then TMA tile kernels on every level
        m.set_global(m.G_TILE_MIN_PLANE, min_plane)
        for case in cases:
            if not feasible(case[0], case[1], world, case[3], case[4]):
                continue
            info = check_case(*case)
            D.barrier()
            if rank == 0:
                print("ok", info, "tile_min_plane", min_plane, flush=True)
    print(f"DIST_CHECK_OK rank {rank}/{world}", flush=True)


if __name__ == "__main__":
    main()
