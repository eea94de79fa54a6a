"""Experiment: the SOL-32 iteration as independent LANES (sub-batches of simulations on their own streams, each with its own
unroll object and CUDA graph) vs one unroll over the whole batch.  Lanes are de-phased by a device-side delay."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from solver_in_the_loop_b200 import engine  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
Y, X, B, m = 128, 64, 3, 32
plan = engine.Plan.karman(Y, X, B)
plan.set_cg(1e-7, 1e-6, 4000, 0)
re, vy0, vx0, gy, gx, sig = bench.synth_batch(plan, engine, torch, B, m, 0, 100)
plan.set_cg(1e-5, 0.0, 2000, 0)
w = torch.randn(260354, device=dev) * 0.005


def timed(fn, n=10):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


un = engine.Unroll(plan, m, B, sig, use_graph=True)
g = torch.zeros(un.nparams, device=dev)
t_one = timed(lambda: un.train_iter(w, re, vy0, vx0, gy, gx, g))
print("one unroll, B=3: %.2f ms" % t_one, flush=True)
gref = g.clone()

for split in ([2, 1], [1, 1, 1]):
    lanes = []
    b0 = 0
    for nb in split:
        p = engine.Plan.karman(Y, X, nb)
        p.set_cg(1e-5, 0.0, 2000, 0)
        u = engine.Unroll(p, m, nb, sig, use_graph=True)
        sl = slice(b0, b0 + nb)
        args = (re[sl].contiguous(), vy0[sl].contiguous(), vx0[sl].contiguous(), gy[:, sl].contiguous(), gx[:, sl].contiguous())
        lanes.append((u, args, torch.zeros(un.nparams, device=dev), torch.cuda.Stream()))
        b0 += nb
    main = torch.cuda.current_stream()
    for delay_us in (0, 60, 120, 180):
        def run():
            ev = torch.cuda.Event(); ev.record(main)
            for k, (u, a, gk, s) in enumerate(lanes):
                s.wait_event(ev)
                with torch.cuda.stream(s):
                    if k and delay_us:
                        torch.cuda._sleep(int(delay_us * k * 1965))      # ~cycles at 1965 MHz
                    u.train_iter(w, *a, gk)
                e = torch.cuda.Event(); e.record(s); main.wait_event(e)
        t = timed(run)
        gs = sum(l[2] for l in lanes)
        err = float((gs - gref).norm() / gref.norm())
        print("lanes %s, stagger %3d us: %.2f ms  (grad vs one-unroll %.1e)" % (split, delay_us, t, err), flush=True)
