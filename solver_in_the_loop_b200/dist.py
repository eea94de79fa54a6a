"""Data parallelism over independent simulations.

The reference has no multi-GPU code (one process, CUDA_VISIBLE_DEVICES=<one id>;
karman-2d/karman_train.py:22,49).  Simulations of a batch are independent in the forward and the
adjoint sweep; they couple only through the shared CNN weights, whose gradients are SUMMED over the
batch (tf.nn.l2_loss sums over the batch, karman_train.py:430).  So: contiguous split of the
simulations over ranks, replicated weights, one all-reduce(SUM) per optimiser step on one flat
bucket [grad(260,354) | per-step losses(m)], then the identical Adam update on every rank.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = "nccl"):
    """torch.distributed.run exports RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, local, world


def shard_range(n_sims: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n_sims simulations; the remainder goes to the lowest ranks."""
    base, rem = divmod(n_sims, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def make_bucket(n_params: int, msteps: int, device) -> torch.Tensor:
    return torch.zeros(n_params + msteps, dtype=torch.float32, device=device)


def allreduce_bucket(bucket: torch.Tensor, group=None) -> torch.Tensor:
    """SUM over ranks, in place: gradients of the summed loss and the global per-step losses."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    return bucket
