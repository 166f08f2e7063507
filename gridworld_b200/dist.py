"""Multi-GPU plumbing: envs shard independently across ranks (one process per GPU), so the step path has
no collective. torch.distributed is only used for the start/stop barrier and the max-over-ranks timing
of bench.py (NCCL on GPUs, gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def rank_info():
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


def shard_range(total_envs, rank, world):
    """Contiguous block of env indices owned by `rank`: sizes differ by at most one."""
    base, rem = divmod(int(total_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def init(backend, device=None):
    rank, world, _ = rank_info()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29512')
        kw = {'device_id': device} if (backend == 'nccl' and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world


def barrier(device=None):
    if device is not None and device.type == 'cuda':
        torch.cuda.synchronize(device)
    if dist.is_initialized():
        dist.barrier()
    if device is not None and device.type == 'cuda':
        torch.cuda.synchronize(device)


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (the bench reports the slowest rank's time)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else 'cpu')
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else 'cpu')
    if dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def shutdown():
    if dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
