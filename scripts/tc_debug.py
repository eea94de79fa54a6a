"""Diagnose the tcgen05 conv kernel against the fp32 SIMT kernel and the fp64 oracle; time both."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import sol_oracle as so  # noqa: E402
from solver_in_the_loop_b200 import engine  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)


def rel(a, b):
    a = a.double().cpu(); b = b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


g = torch.Generator().manual_seed(0)
for (B, Y, X) in [(1, 16, 8), (1, 32, 16), (2, 24, 32), (3, 128, 64)]:
    x = torch.randn(B, Y, X, 32, generator=g, dtype=torch.float64)
    w = torch.randn(5, 5, 32, 32, generator=g, dtype=torch.float64) * 0.05
    b = torch.randn(32, generator=g, dtype=torch.float64)
    add = torch.randn(B, Y, X, 32, generator=g, dtype=torch.float64)
    ref = so._conv(x, w, b) if Y * X * B <= 2048 else None
    f = lambda t: t.to(dev, torch.float32).contiguous()
    engine.set_option("conv_path", 1)
    o_simt = engine.conv5x5(f(x), f(w), f(b))
    for mode in (1, 0):
        engine.set_option("conv_path", 2)
        engine.set_option("tc_base_offset_mode", mode)
        try:
            o_tc = engine.conv5x5(f(x), f(w), f(b))
            torch.cuda.synchronize()
            msg = "tc(mode=%d) vs simt %.3e" % (mode, rel(o_tc, o_simt))
            if ref is not None:
                msg += "  tc vs fp64 %.3e  simt vs fp64 %.3e" % (rel(o_tc, ref), rel(o_simt, ref))
            print((B, Y, X), msg, flush=True)
            if mode == 1:
                o2 = engine.conv5x5(f(x), f(w), f(b), addend=f(add), act=1)
                engine.set_option("conv_path", 1)
                o2s = engine.conv5x5(f(x), f(w), f(b), addend=f(add), act=1)
                print("   epilogue(addend+lrelu) tc vs simt %.3e" % rel(o2, o2s), flush=True)
        except Exception as e:
            print((B, Y, X), "mode", mode, "FAILED", e, flush=True)
# timing at the bench shape
B, Y, X = 3, 128, 64
x = torch.randn(B, Y, X, 32, device=dev); w = torch.randn(5, 5, 32, 32, device=dev) * 0.05; b = torch.randn(32, device=dev)
for path in (1, 2):
    engine.set_option("conv_path", path)
    engine.set_option("tc_base_offset_mode", 1)
    for _ in range(3):
        engine.conv5x5(x, w, b, act=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        engine.conv5x5(x, w, b, act=1)
    e1.record(); torch.cuda.synchronize()
    print("path", path, "us per conv (incl. weight prep for tc)", e0.elapsed_time(e1) * 1e3 / 50, flush=True)
