"""Time the pressure projection (CUDA events, 50 launches) for the direct solver vs the multigrid-preconditioned CG, and check
the direct result against the iterative one."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from solver_in_the_loop_b200 import engine  # noqa: E402

torch.cuda.set_device(0)
for (Y, X) in [(128, 64), (64, 32)]:
    for B in (3, 148):
        plan = engine.Plan.karman(Y, X, B)
        plan.set_option("direct_solve", 0)
        plan.set_cg(1e-7, 1e-6, 4000, 0)
        re, vy0, vx0, gy, gx, sig = bench.synth_batch(plan, engine, torch, B, 1, 0, 30)
        o = plan.step_fwd(re, vy0, vx0)
        ay, ax = plan.advect(o["vy1"], o["vx1"])
        ref = None
        for direct in (0, 1):
            plan.set_option("direct_solve", direct)
            plan.set_cg(1e-7, 1e-6, 4000, 0) if not direct else None
            for _ in range(3):
                py, px, pp, it = plan.project(ay, ax, return_pressure=True)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                py, px, it = plan.project(ay, ax)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 50
            if ref is None:
                ref = (py.clone(), pp.clone())
                print("grid %dx%d B=%3d mgpcg(tight): %7.1f us  iters %.0f" % (Y, X, B, us, float(it.float().mean())), flush=True)
            else:
                d = float((py - ref[0]).norm() / ref[0].norm()); dp = float((pp - ref[1]).norm() / ref[1].norm())
                print("grid %dx%d B=%3d direct      : %7.1f us  vy vs mgpcg %.1e  p vs mgpcg %.1e" % (Y, X, B, us, d, dp), flush=True)
        plan.close()
