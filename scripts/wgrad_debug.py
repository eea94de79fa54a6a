"""Diagnose the tcgen05 weight-gradient kernel against torch autograd (fp64)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from oracle import sol_oracle as so
from solver_in_the_loop_b200 import engine
dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
g = torch.Generator().manual_seed(0)
for (B, Y, X) in [(1, 16, 8), (1, 32, 16), (2, 32, 32)]:
    x = torch.randn(B, Y, X, 32, generator=g, dtype=torch.float64)
    go = torch.randn(B, Y, X, 32, generator=g, dtype=torch.float64)
    w = torch.zeros(5, 5, 32, 32, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(32, dtype=torch.float64, requires_grad=True)
    (so._conv(x, w, b) * go).sum().backward()
    ref = w.grad
    f = lambda t: t.to(dev, torch.float32).contiguous()
    engine.set_option("wgrad_path", 2)
    dW, db = engine.conv5x5_wgrad(f(x), f(go))
    torch.cuda.synchronize()
    a = dW.double().cpu()
    print((B, Y, X), "||a|| %.4e ||ref|| %.4e  slope <a,ref>/<ref,ref> %.4f  rel %.3e  db rel %.3e" % (
        a.norm(), ref.norm(), float((a * ref).sum() / (ref * ref).sum()), float((a - ref).norm() / ref.norm()),
        float((db.double().cpu() - b.grad).norm() / b.grad.norm())), flush=True)
    # per-tap diagnosis: which reference tap does each computed tap correlate with?
    for dy in range(5):
        row = []
        for dx in range(5):
            t = a[dy, dx].reshape(-1)
            best = max(((float((t * ref[ey, ex].reshape(-1)).sum() / (ref[ey, ex].norm() * t.norm() + 1e-30)), ey, ex)
                        for ey in range(5) for ex in range(5)), key=lambda z: abs(z[0]))
            tt = max(((float((a[dy, dx].t().reshape(-1) * ref[ey, ex].reshape(-1)).sum() / (ref[ey, ex].norm() * t.norm() + 1e-30)), ey, ex)
                      for ey in range(5) for ex in range(5)), key=lambda z: abs(z[0]))
            row.append("%+.2f@(%d,%d)|T%+.2f@(%d,%d) n=%.1e" % (best[0], best[1], best[2], tt[0], tt[1], tt[2], float(t.norm())))
        print("   dy", dy, "  ".join(row), flush=True)
    engine.set_option("wgrad_path", 1)
    dW1, db1 = engine.conv5x5_wgrad(f(x), f(go))
    print("   simt rel %.3e" % float((dW1.double().cpu() - ref).norm() / ref.norm()), flush=True)
