"""Process-group bootstrap and autograd-aware collectives (src/utils/distributed.py).

One process per GPU, NCCL over NVLink 5 / NVSwitch.  `init_distributed` keeps the reference's
behaviour (early return if a group exists, SLURM variables, port 37123, fall back to (1, 0) on
failure); the backend is NCCL when CUDA is present and gloo otherwise so the host logic can be
exercised on CPU-only machines.
"""
import os
from logging import getLogger

import torch
import torch.distributed as dist

logger = getLogger()


def _active():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def init_distributed(port=37123, rank_and_world_size=(None, None)):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    rank, world_size = rank_and_world_size
    os.environ.setdefault('MASTER_ADDR', 'localhost')
    if (rank is None) or (world_size is None):
        try:
            world_size = int(os.environ['SLURM_NTASKS'])
            rank = int(os.environ['SLURM_PROCID'])
            os.environ['MASTER_ADDR'] = os.environ['HOSTNAME']
        except Exception:
            logger.info('SLURM vars not set (distributed training not available)')
            return 1, 0
    try:
        os.environ['MASTER_PORT'] = str(port)
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        dist.init_process_group(backend=backend, world_size=world_size, rank=rank)
    except Exception as e:
        world_size, rank = 1, 0
        logger.info(f'Rank: {rank}. Distributed training not available {e}')
    return world_size, rank


class AllGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if _active():
            x = x.contiguous()
            parts = [torch.zeros_like(x) for _ in range(dist.get_world_size())]
            dist.all_gather(parts, x)
            return torch.cat(parts, 0)
        return x

    @staticmethod
    def backward(ctx, grads):
        if _active():
            per = grads.shape[0] // dist.get_world_size()
            grads = grads.contiguous()
            dist.all_reduce(grads)
            return grads[per * dist.get_rank(): per * (dist.get_rank() + 1)]
        return grads


class AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        if _active():
            x = x.contiguous()
            dist.all_reduce(x)
        return x

    @staticmethod
    def backward(ctx, grads):
        return grads


class AllReduce(torch.autograd.Function):
    """Average across ranks; identity gradient."""

    @staticmethod
    def forward(ctx, x):
        if _active():
            x = x.contiguous() / dist.get_world_size()
            dist.all_reduce(x)
        return x

    @staticmethod
    def backward(ctx, grads):
        return grads
