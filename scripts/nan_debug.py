"""Hunt the NaN in the 64x32 m=2 backward: poison the workspace, vary CG geometry, repeat."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import sol_oracle as so  # noqa: E402
from solver_in_the_loop_b200 import engine  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
Y, X, B, m = 64, 32, 2, 2
geom, rho, vy, vx, re, gty, gtx, sig = so.make_case(Y=Y, X=X, B=B, msteps=m, spin=25)
params = so.init_params(seed=0)
f = lambda t: t.to(dev, torch.float32).contiguous()
w = f(so.flatten_params(params))
for poison in (False, True):
    for rows in (0, 2, 4, 8, 16):
        plan = engine.Plan.karman(Y, X, B)
        plan.set_cg(1e-7, 1e-6, 4000, 0)
        plan.set_option("cg_rows", rows)
        un = engine.Unroll(plan, m, B, sig)
        if poison:
            un.workspace.view(torch.float32)[:un.workspace.numel() // 4].fill_(float("nan"))
        for rep in range(2):
            ls = un.forward(w, f(re), f(vy), f(vx), f(gty), f(gtx))
            g, gy0, gx0 = un.backward(w, want_input_grad=True)
            torch.cuda.synchronize()
            print("poison", poison, "rows", rows, "rep", rep, "loss", ls.tolist(), "nan in grad", bool(torch.isnan(g).any()),
                  "gnorm", float(g.norm()), "iters", un.cg_iters().tolist(), flush=True)
        # stage-level replay of the adjoint of one step, checking each stage for NaN
        out = plan.step_fwd(f(re), f(vy), f(vx))
        gyo = torch.randn_like(out["vy"]) * 100; gxo = torch.randn_like(out["vx"]) * 100
        py, px, it = plan.project(gyo, gxo)
        ay, ax = plan.advect_bwd(out["vy1"], out["vx1"], py, px)
        dy_, dx_ = plan.diffuse_bc_bwd(f(re), ay, ax)
        print("   stage nan: project", bool(torch.isnan(py).any() or torch.isnan(px).any()), "advect_bwd", bool(torch.isnan(ay).any()),
              "diffuse_bwd", bool(torch.isnan(dy_).any()), "iters", it.tolist(), flush=True)
        un.close(); plan.close()
