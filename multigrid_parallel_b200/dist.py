"""Multi-process plumbing: one process per GPU (torchrun), torch.distributed
for the rendezvous and host-side barriers, NCCL inside libmgb for the data
path (halo planes, norm all-reduce, coarse-level gather/broadcast)."""
import os

from .solver import Solver


def env_ranks():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend="gloo"):
    """rendezvous from the torchrun environment (MASTER_ADDR/MASTER_PORT)"""
    import torch.distributed as dist
    rank, world, local_rank = env_ranks()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local_rank


def broadcast_bytes(payload, src=0):
    import torch.distributed as dist
    box = [payload]
    if dist.is_initialized():
        dist.broadcast_object_list(box, src=src)
    return box[0]


def make_solver(coarse, levels, gs, min_planes=16, min_points=-1, device=None):
    """collective: every rank gets its slab of one partitioned hierarchy"""
    rank, world, local_rank = env_ranks()
    device = local_rank if device is None else device
    if world == 1:
        return Solver(coarse, levels, gs, device=device)
    uid = Solver.nccl_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, 0)
    return Solver(coarse, levels, gs, device=device, rank=rank, nranks=world, nccl_uid=uid,
                  min_planes=min_planes, min_points=min_points)


def barrier():
    import torch.distributed as dist
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(x):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])
