"""One training iteration of the bench workload between cudaProfilerStart/Stop (for ncu
--profile-from-start off).  Usage: python scripts/profile_iter.py [--msteps 32] [--graph]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from solver_in_the_loop_b200 import engine  # noqa: E402
from solver_in_the_loop_b200.trainer import SolTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--msteps", type=int, default=32)
ap.add_argument("--batch", type=int, default=3)
ap.add_argument("--Y", type=int, default=128)
ap.add_argument("--X", type=int, default=64)
ap.add_argument("--spin", type=int, default=30)
ap.add_argument("--graph", action="store_true")
ap.add_argument("--cluster", type=int, default=0)
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--conv-path", type=int, default=0)
ap.add_argument("--wgrad-path", type=int, default=0)
ap.add_argument("--cg-rows", type=int, default=0)
ap.add_argument("--cg-precond", type=int, default=1)
ap.add_argument("--direct-solve", type=int, default=1)
a = ap.parse_args()
torch.cuda.set_device(0)
engine.set_option("conv_path", a.conv_path)
engine.set_option("wgrad_path", a.wgrad_path)
plan = engine.Plan.karman(a.Y, a.X, a.batch)
plan.set_option("cg_rows", a.cg_rows)
plan.set_option("cg_precond", a.cg_precond)
plan.set_option("direct_solve", a.direct_solve)
plan.set_cg(1e-7, 1e-6, 4000, a.cluster)
re, vy0, vx0, gy, gx, sig = bench.synth_batch(plan, engine, torch, a.batch, a.msteps, 0, a.spin)
plan.set_cg(1e-5, 0.0, 2000, a.cluster)
tr = SolTrainer(plan, a.msteps, a.batch, sig, use_graph=a.graph)
tr.weights.mul_(0.1)
for _ in range(3 if a.graph else 1):
    tr.train_step(re, vy0, vx0, gy, gx)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.iters):
    tr.train_step(re, vy0, vx0, gy, gx)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("cg iters", tr.unroll.cg_iters().float().mean(dim=(1, 2)).tolist())
