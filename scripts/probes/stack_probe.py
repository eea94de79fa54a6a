import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from solver_in_the_loop_b200 import _lib, engine
import bench
torch.cuda.set_device(0)
lib = _lib.load()
lib.sol_debug_conv_stack_capacity.restype = ctypes.c_int
print("capacity(3,128,64) =", lib.sol_debug_conv_stack_capacity(3, 128, 64))
plan = engine.Plan.karman(128, 64, 3)
plan.set_cg(1e-5, 0.0, 2000, 0)
re, vy0, vx0, gy, gx, sig = bench.synth_batch(plan, engine, torch, 3, 2, 0, 10)
for stack in (0, 1):
    engine.set_option("conv_stack", stack)
    un = engine.Unroll(plan, 2, 3, sig)
    w = torch.randn(un.nparams, device="cuda") * 0.01
    g = torch.zeros(un.nparams, device="cuda")
    l0 = lib.sol_launch_count()
    ls = un.train_iter(w, re, vy0, vx0, gy, gx, g)
    torch.cuda.synchronize()
    print("stack", stack, "launches", lib.sol_launch_count() - l0, "loss", ls.tolist(), "gnorm", float(g.norm()))
