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


class DataParallelStep:
    """The rank-independent half of one optimiser step (karman_train.py:449-457): the flat bucket
    [gradients | per-step losses] of this rank's simulations -> ONE all-reduce(SUM) -> optional per-variable
    clip_by_norm -> the identical TF1-Adam update on every rank.  `SolTrainer` fills the bucket with the CUDA engine
    and updates with the `sol_adam_tf1` kernel; the world-size-2 gloo test fills it on CPU and drives this same code.

    Subclasses provide `_adam(lr)` (update self.weights from self.grad) and `layer_shapes` ((Cin, Cout) per Conv2D).
    """

    def _init_bucket(self, n_params: int, msteps: int, device, process_group=None):
        self.pg = process_group
        self.world = 1
        if process_group is not None or (dist.is_available() and dist.is_initialized()):
            self.world = dist.get_world_size(process_group)
        self.n_params, self.msteps = int(n_params), int(msteps)
        self.bucket = make_bucket(n_params, msteps, device)
        self.grad = self.bucket[:n_params]
        self.loss_steps = self.bucket[n_params:]
        self.t = 0

    def _clip_by_norm(self, clip: float):
        """tf.clip_by_norm(grad, 1e-3) per variable (karman_train.py:452-454)."""
        o = 0
        for ci, co in self.layer_shapes:
            for n in (25 * ci * co, co):
                g = self.grad[o:o + n]
                nrm = g.norm()
                g.mul_(torch.clamp(clip / (nrm + 1e-30), max=1.0))
                o += n

    def reduce_and_update(self, lr: float, clip_grad: bool = False) -> torch.Tensor:
        """All-reduce the bucket, clip, Adam.  Returns the global total loss (0-d tensor): sum_i loss_i / msteps
        over all simulations of all ranks (karman_train.py:436)."""
        if self.world > 1:
            allreduce_bucket(self.bucket, self.pg)
        if clip_grad:
            self._clip_by_norm(1e-3)
        self.t += 1
        self._adam(lr)
        return self.loss_steps.sum() / self.msteps
