"""Time the fused projection kernel for (cluster, rows/thread) variants.  CUDA events, 20 reps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from solver_in_the_loop_b200 import engine  # noqa: E402

torch.cuda.set_device(0)
for (Y, X) in [(128, 64), (256, 128)]:
    for B in (3, 148):
        plan = engine.Plan.karman(Y, X, B)
        plan.set_option("direct_solve", 0)      # this script times the iterative solvers (scripts/direct_bench.py: the direct one)
        plan.set_cg(1e-7, 1e-6, 4000, 0)
        re, vy0, vx0, gy, gx, sig = bench.synth_batch(plan, engine, torch, B, 1, 0, 30)
        o = plan.step_fwd(re, vy0, vx0)
        ay, ax = plan.advect(o["vy1"], o["vx1"])
        for cl, rows, pre in [(1, 8, 0), (1, 16, 0), (1, 8, 1), (1, 16, 1), (1, 8, 2), (1, 16, 2), (4, 8, 0), (8, 4, 0)]:
            if True:
                try:
                    plan.set_cg(1e-5, 0.0, 2000, cl)
                    plan.set_option("cg_rows", rows)
                    plan.set_option("cg_precond", 1 if pre else 0)
                    plan.set_option("mg_variant", 2 if pre == 2 else 0)
                    for _ in range(3):
                        py, px, it = plan.project(ay, ax)
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(20):
                        py, px, it = plan.project(ay, ax)
                    e1.record()
                    torch.cuda.synchronize()
                    us = e0.elapsed_time(e1) * 1e3 / 20
                    K = float(it.float().mean())
                    alg = (40 * K + 8) * Y * X * B
                    print("grid %dx%d B=%3d cluster=%d rows=%2d precond=%d: %8.1f us  K=%.0f  %.3f us/iter  alg(40K+8) %.0f GB/s" %
                          (Y, X, B, cl, rows, pre, us, K, us / max(K, 1), alg / us / 1e3), flush=True)
                except Exception as e:   # unsupported combination
                    print("grid %dx%d B=%3d cluster=%d rows=%2d precond=%d: unsupported (%s)" % (Y, X, B, cl, rows, pre, str(e)[:60]), flush=True)
        plan.close()
