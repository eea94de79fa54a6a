"""N>1 host logic on CPU: world_size-2 gloo.  Each rank computes the weight gradient of ITS shard of
the simulations (with the CPU oracle standing in for the GPU engine), the flat bucket is
all-reduced, and every rank must hold the gradient of the full batch and take the same Adam step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sol_oracle as so


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _grad(case, lo, hi, m):
    geom, rho, vy, vx, re, gty, gtx, sig = case
    params = [p.clone().requires_grad_() for p in so.init_params(seed=0)]
    loss, losses = so.unrolled_loss(params, rho[lo:hi], vy[lo:hi], vx[lo:hi], re[lo:hi], gty[:, lo:hi], gtx[:, lo:hi], geom, sig, m)
    loss.backward()
    return so.flatten_params([p.grad for p in params]), torch.stack([l.detach() for l in losses])


class _OracleShardStep:
    """dist.DataParallelStep (the class SolTrainer inherits its all-reduce / clip / Adam sequence from) with the CPU
    oracle filling the bucket instead of the CUDA engine."""

    def __new__(cls, *a, **k):
        from solver_in_the_loop_b200.dist import DataParallelStep

        class Impl(DataParallelStep):
            def __init__(self, n, m):
                self._init_bucket(n, m, "cpu")
                from solver_in_the_loop_b200.trainer import model_layers
                self.layer_shapes = model_layers("mars_moon", 3)
                self.weights = so.flatten_params(so.init_params(seed=0)).float()
                self.adam_m = torch.zeros_like(self.weights); self.adam_v = torch.zeros_like(self.weights)

            def _adam(self, lr):
                self.weights, self.adam_m, self.adam_v = so.adam_tf1_step(self.weights, self.grad, self.adam_m, self.adam_v, self.t, lr)

        return Impl(*a, **k)


def _worker(rank, world, port, out, clip):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from solver_in_the_loop_b200 import dist as sd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    m, B = 2, 4
    case = so.make_case(Y=32, X=32, B=B, msteps=m, spin=6)
    lo, hi = sd.shard_range(B, rank, world)
    g, l = _grad(case, lo, hi, m)
    step = _OracleShardStep(g.numel(), m)
    assert step.world == world
    step.grad.copy_(g.float()); step.loss_steps.copy_(l.float())
    total = step.reduce_and_update(1e-4, clip_grad=clip)
    if rank == 0:
        gfull, lfull = _grad(case, 0, B, m)
        if clip:      # tf.clip_by_norm per variable (karman_train.py:452-454)
            o = 0
            for ci, co in so.model_layers():
                for n in (25 * ci * co, co):
                    nrm = gfull[o:o + n].norm()
                    gfull[o:o + n] *= min(1.0, 1e-3 / float(nrm))
                    o += n
        theta, _, _ = so.adam_tf1_step(so.flatten_params(so.init_params(seed=0)), gfull, torch.zeros_like(gfull), torch.zeros_like(gfull), 1, 1e-4)
        out.put((float((step.grad.double() - gfull).norm() / gfull.norm()),
                 float((step.loss_steps.double() - lfull).norm() / lfull.norm()),
                 float((step.weights.double() - theta).norm() / theta.norm()),
                 abs(float(total) - float(lfull.sum()) / m) / abs(float(lfull.sum()) / m)))
    gathered = [torch.zeros_like(step.weights) for _ in range(world)]
    dist.all_gather(gathered, step.weights)
    assert all(torch.equal(gathered[0], t) for t in gathered)      # identical update on every rank
    dist.destroy_process_group()


def test_shard_range_partition():
    from solver_in_the_loop_b200.dist import shard_range
    for n in (1, 3, 6, 7, 32):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("clip", [False, True], ids=["plain", "clip-grad"])
def test_two_rank_gradient_allreduce_matches_single_process(clip):
    """The trainer's own data-parallel step (dist.DataParallelStep.reduce_and_update) on 2 gloo ranks == the
    single-process gradient / loss / Adam update of the full batch; with --clip-grad the per-variable clip_by_norm."""
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, clip)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    eg, el, ew, et = out.get(timeout=10)
    assert eg < 1e-6 and el < 1e-6 and ew < 1e-6 and et < 1e-6, (eg, el, ew, et)
