"""How much of the multigrid solve time is the irregular (obstacle-adjacent) code path?  Times the fused projection on the
karman scene and on the same open domain without the obstacle (all warps take the regular path)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from solver_in_the_loop_b200 import engine  # noqa: E402

torch.cuda.set_device(0)
Y, X, B = 128, 64, 3
g = torch.Generator().manual_seed(0)
for name, plan in (("karman (obstacle)", engine.Plan.karman(Y, X, B)),
                   ("open box, no obstacle", engine.Plan(Y, X, B, 100.0 / X, solid=np.zeros((Y, X), np.uint8), inflow=np.zeros((Y, X), np.float32),
                                                          bc_mask_y=np.zeros((Y + 1, X), np.float32), bc_val_y=np.zeros((Y + 1, X), np.float32)))):
    vy = torch.randn(B, Y + 1, X, generator=g).cuda() * 0.1 + 1.0
    vx = torch.randn(B, Y, X + 1, generator=g).cuda() * 0.1
    plan.set_cg(1e-5, 0.0, 2000, 0)
    for _ in range(3):
        py, px, it = plan.project(vy, vx)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        py, px, it = plan.project(vy, vx)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    K = float(it.float().mean())
    print("%-24s %7.1f us  K=%.1f  %.2f us/iter" % (name, us, K, us / K), flush=True)
