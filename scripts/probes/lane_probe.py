"""Probe for the per-simulation-lane schedule: period of a dependent chain of tensor-core conv layers for B = 1, 2, 3
images on one stream, and of three B = 1 chains on three streams (with and without a pressure solve per 12 layers)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from solver_in_the_loop_b200 import engine  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
Y, X = 128, 64
wl = torch.randn(5, 5, 32, 32, device=dev) * 0.03
bl = torch.randn(32, device=dev) * 0.1
ws = engine.conv5x5_split_weights(wl)
torch.cuda.synchronize()


def chain(a, b, n):
    for _ in range(n // 2):
        engine.conv5x5_c32_presplit(a, ws, bl, act=1, out=b, weights_settled=True)
        engine.conv5x5_c32_presplit(b, ws, bl, act=1, out=a, weights_settled=True)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3


N = 120
for B in (1, 2, 3, 6):
    a = torch.randn(B, Y, X, 32, device=dev); b = torch.empty_like(a)
    us = timed(lambda: chain(a, b, N))
    print("one stream, B=%d: %.2f us per layer" % (B, us / N), flush=True)

# three lanes of B = 1 on three streams
plan = engine.Plan.karman(Y, X, 3)
plan.set_cg(1e-7, 1e-6, 4000, 0)
re, vy0, vx0, gy, gx, sig = bench.synth_batch(plan, engine, torch, 3, 1, 0, 30)
plan.set_cg(1e-5, 0.0, 2000, 0)
plans = [engine.Plan.karman(Y, X, 1) for _ in range(3)]
for p in plans:
    p.set_cg(1e-5, 0.0, 2000, 0)
o = plan.step_fwd(re, vy0, vx0)
ay, ax = plan.advect(o["vy1"], o["vx1"])
lanes = [(torch.randn(1, Y, X, 32, device=dev), torch.empty(1, Y, X, 32, device=dev), torch.cuda.Stream()) for _ in range(3)]
main = torch.cuda.current_stream()


def three(with_solve, steps=10):
    ev = torch.cuda.Event()
    ev.record(main)
    for k, (a, b, s) in enumerate(lanes):
        s.wait_event(ev)
        with torch.cuda.stream(s):
            for _ in range(steps):
                if with_solve:
                    plans[k].project(ay[k:k + 1].contiguous(), ax[k:k + 1].contiguous())
                chain(a, b, 12)
        e = torch.cuda.Event()
        e.record(s)
        main.wait_event(e)


def one(with_solve, steps=10):
    a = lanes[0][0].repeat(3, 1, 1, 1).contiguous(); b = torch.empty_like(a)
    for _ in range(steps):
        if with_solve:
            plan.project(ay, ax)
        chain(a, b, 12)


for ws_ in (False, True):
    t3 = timed(lambda: three(ws_))
    t1 = timed(lambda: one(ws_))
    print("10 steps x (%s12 conv layers): 3 lanes of B=1 on 3 streams %.1f us | one stream B=3 %.1f us" %
          ("solve + " if ws_ else "", t3, t1), flush=True)
