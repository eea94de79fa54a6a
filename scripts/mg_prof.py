import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from solver_in_the_loop_b200 import engine
torch.cuda.set_device(0)
Y, X, B = 128, 64, 3
plan = engine.Plan.karman(Y, X, B)
plan.set_cg(1e-7, 1e-6, 4000, 0)
re, vy0, vx0, gy, gx, sig = bench.synth_batch(plan, engine, torch, B, 1, 0, 30)
o = plan.step_fwd(re, vy0, vx0)
ay, ax = plan.advect(o["vy1"], o["vx1"])
plan.set_cg(1e-5, 0.0, 2000, 0)
for _ in range(3):
    py, px, it = plan.project(ay, ax)
torch.cuda.synchronize()
torch.cuda.profiler.start()
py, px, it = plan.project(ay, ax)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(it.tolist())
