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


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from solver_in_the_loop_b200 import dist as sd
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    m, B = 2, 4
    case = so.make_case(Y=32, X=32, B=B, msteps=m, spin=6)
    lo, hi = sd.shard_range(B, rank, world)
    g, l = _grad(case, lo, hi, m)
    bucket = sd.make_bucket(g.numel(), m, "cpu")
    bucket[:g.numel()] = g.float(); bucket[g.numel():] = l.float()
    sd.allreduce_bucket(bucket)
    theta = so.flatten_params(so.init_params(seed=0)).float()
    new, _, _ = so.adam_tf1_step(theta, bucket[:g.numel()], torch.zeros_like(theta), torch.zeros_like(theta), 1, 1e-4)
    if rank == 0:
        gfull, lfull = _grad(case, 0, B, m)
        out.put((float((bucket[:g.numel()].double() - gfull).norm() / gfull.norm()),
                 float((bucket[g.numel():].double() - lfull).norm() / lfull.norm())))
    gathered = [torch.zeros_like(new) for _ in range(world)]
    dist.all_gather(gathered, new)
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


def test_two_rank_gradient_allreduce_matches_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    eg, el = out.get(timeout=10)
    assert eg < 1e-6 and el < 1e-6, (eg, el)
