"""CPU, world_size 2 (gloo): the host-side logic of the multi-GPU path.

* the slab planner of the C ABI (mgb_plan_slab, mgb_plan_first_dist_level)
  tiles every partitioned level exactly and nests across levels;
* the rendezvous helper ships the NCCL id from rank 0 to all ranks;
* the partitioned V-cycle SCHEDULE (tests/dist_model.py, a gloo mirror of what
  libmgb does over NCCL) reproduces the serial oracle bit for bit;
* without a GPU the partitioned solver fails loudly (no fallback)."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_slab_planner_tiles_and_nests(mgb):
    lib = mgb.load_library()
    a, b = C.c_int(), C.c_int()
    for P in (1, 2, 4, 8):
        for ci in (3, 5, 2 * P + 1):
            for level in range(1, 8):
                ni = (ci - 1) * (1 << level) + 1
                if (ni - 1) % P:
                    assert lib.mgb_plan_slab(ni, P, 0, a, b) != 0
                    continue
                cover = []
                for r in range(P):
                    assert lib.mgb_plan_slab(ni, P, r, a, b) == 0
                    cover.append((a.value, b.value))
                assert cover[0][0] == 0 and cover[-1][1] == ni
                assert all(cover[r][1] == cover[r + 1][0] for r in range(P - 1))
                if level >= 2 and ((ni - 1) // 2) % P == 0:  # nesting with the coarser level
                    nic = (ni + 1) // 2
                    for r in range(P):
                        lib.mgb_plan_slab(nic, P, r, a, b)
                        assert cover[r][0] == 2 * a.value
    # thresholds: 1025^3 over 8 ranks partitions 257^3 and finer by default
    assert lib.mgb_plan_first_dist_level(3, 3, 3, 10, 8, 16, 1 << 20) == 7
    assert lib.mgb_plan_first_dist_level(3, 3, 3, 10, 1, 16, 1 << 20) == 0
    assert lib.mgb_plan_first_dist_level(17, 3, 3, 9, 8, 16, 1 << 20) == 6
    assert lib.mgb_plan_first_dist_level(3, 3, 3, 3, 8, 16, 0) == 3  # nothing splittable


def _worker(rank, world, port, coarse, levels, gs, min_planes, cycles, q, shortcuts=False,
            fused=False):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, HERE)
        os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                          MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        import torch.distributed as dist

        import multigrid_parallel_b200 as m
        from dist_model import SlabMG
        from multigrid_parallel_b200 import dist as D
        from oracle_lib import Orc, OrcMG, seeded

        r, w, _ = D.init_process_group("gloo")
        assert (r, w) == (rank, world)
        lib = m.load_library()
        # rendezvous helper: rank 0's payload reaches everybody
        payload = D.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 0)
        assert payload == bytes(range(128))
        assert D.max_over_ranks(float(rank)) == world - 1
        # no GPU here: the partitioned solver must refuse, not fall back
        try:
            import torch
            if not torch.cuda.is_available():
                try:
                    m.Solver(coarse, levels, gs, device=0, rank=rank, nranks=world,
                             nccl_uid=bytes(128))
                    raise AssertionError("partitioned solver ran without a GPU")
                except m.MgbError:
                    pass
        finally:
            pass
        orc = Orc()
        slab = SlabMG(lib, orc, coarse, levels, gs, rank, world, min_planes=min_planes,
                      shortcuts=shortcuts, fused=fused)
        serial = OrcMG(orc, coarse, levels, gs)
        top = levels - 1
        shape = serial.dims(top)
        u0, d0 = seeded(shape, 61), seeded(shape, 62)
        serial.u(top)[...] = u0
        serial.d(top)[...] = d0
        lt = slab.lv[top]
        lt.u[...] = u0[lt.i0:lt.i0 + lt.li]
        lt.d[...] = d0[lt.i0:lt.i0 + lt.li]
        for c in range(cycles):
            want = serial.vcycle()
            got = slab.vcycle()
            assert abs(want - got) <= 1e-13 * want, (c, want, got)
            for lvl in range(slab.LD, levels):
                lv = slab.lv[lvl]
                lo = lv.own_lo - (1 if rank > 0 else 0)
                hi = lv.own_hi + (1 if rank < world - 1 else 0)
                assert np.array_equal(lv.u[lv.loc(lo):lv.loc(hi)], serial.u(lvl)[lo:hi]), (c, lvl)
                if lvl < top:
                    assert np.array_equal(lv.d[lv.loc(lv.own_lo):lv.loc(lv.own_hi)],
                                          serial.d(lvl)[lv.own_lo:lv.own_hi]), (c, lvl, "d")
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok", slab.LD))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail", traceback.format_exc()))


@pytest.mark.parametrize("shortcuts,fused", [(False, False), (True, False), (True, True)])
@pytest.mark.parametrize("coarse,levels,gs,min_planes", [
    ((3, 3, 3), 5, 2, 2),      # cube 33^3, partitioned down to level 1
    ((3, 3, 3), 5, 2, 8),      # levels < 3 unpartitioned (on rank 0 / replicated)
    ((5, 3, 3), 4, 1, 2),      # weak-scaling box (2P+1) x 3 x 3
])
def test_partitioned_schedule_matches_serial_oracle(coarse, levels, gs, min_planes, shortcuts,
                                                    fused):
    """shortcuts=True: the schedule as libmgb runs it -- coarse levels never zeroed (stale
    halos must all be refreshed by the exchanges before they are read), RED-only
    prolongation without a halo fence -- over 3 cycles.  fused=True: the P2P path's
    schedule (csrc/halo.cuh): only the swept colour of the boundary planes travels with
    each half-sweep (two planes up, one down), no exchange in front of the restriction,
    the unpartitioned levels all-gathered and computed redundantly on every rank."""
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker,
                         args=(r, world, port, coarse, levels, gs, min_planes, 3 if shortcuts else 2,
                               q, shortcuts, fused))
             for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", f"rank {rank}:\n{info}"
    assert len({info for _, _, info in results}) == 1
