#!/usr/bin/env python
"""Phase timeline of the 3xFP16 tcgen05 convolution kernel inside a graph-replayed dependent chain (clock64 / globaltimer
stamps per CTA through the diagnostics hook).  Usage: python scripts/conv_trace.py [variant] [B]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from solver_in_the_loop_b200 import _lib, engine  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()
lib.sol_debug_conv_h_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.sol_debug_conv_h_trace.restype = None
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B, Y, X = (int(sys.argv[2]) if len(sys.argv) > 2 else 3), 128, 64
engine.set_option("conv_path", 2); engine.set_option("conv_variant", variant)
a0 = torch.randn(B, Y, X, 32, device=dev); a1 = torch.empty_like(a0)
w = torch.randn(5, 5, 32, 32, device=dev) * 0.01; b = torch.randn(32, device=dev) * 0.1
ws = engine.conv5x5_split_weights(w)
torch.cuda.synchronize()
nct = (X // 8) * (Y // 16) * B
NL = 12
for _ in range(3):
    engine.conv5x5_c32_presplit(a0, ws, b, act=1, out=a1, weights_settled=True)
    engine.conv5x5_c32_presplit(a1, ws, b, act=1, out=a0, weights_settled=True)
torch.cuda.synchronize()
tr = torch.zeros(NL, nct, 16, dtype=torch.int64, device=dev)
lib.sol_debug_conv_h_trace(ctypes.c_void_p(tr.data_ptr()), NL)      # before capture: the pointers are baked into the graph
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    for _ in range(NL // 2):
        engine.conv5x5_c32_presplit(a0, ws, b, act=1, out=a1, weights_settled=True)
        engine.conv5x5_c32_presplit(a1, ws, b, act=1, out=a0, weights_settled=True)
lib.sol_debug_conv_h_trace(None, 0)
for _ in range(3):
    graph.replay()
torch.cuda.synchronize()
t = tr.cpu().numpy()
names = {3: "setup", 4: "halo landed", 5: "converted", 6: "mma may start", 7: "first weights", 8: "mmas issued",
         9: "acc complete", 10: "store issued"}
print("variant %d, B=%d: CTAs %d" % (variant, B, nct))
ends = [t[l][:, 12].max() for l in range(NL)]
print("launch period (last CTA end -> last CTA end), us:", " ".join("%.2f" % ((ends[l] - ends[l - 1]) / 1e3) for l in range(1, NL)))
for l in (NL - 4, NL - 3):
    tl = t[l]
    prev_end = ends[l - 1]
    print("== launch %d: first CTA start %.2f us / last CTA start %.2f us relative to the END of launch %d; own end %.2f us"
          % (l, (tl[:, 1].min() - prev_end) / 1e3, (tl[:, 1].max() - prev_end) / 1e3, l - 1, (ends[l] - prev_end) / 1e3))
    cnt = np.bincount(tl[:, 0].astype(np.int64), minlength=200)
    print("   SMs used %d, CTAs/SM max %d" % ((cnt > 0).sum(), cnt.max()))
    for multi in (1, 2, 3):
        sel = np.array([cnt[int(s)] == multi for s in tl[:, 0]])
        if not sel.any():
            continue
        print("   -- SMs with %d CTA(s) of this launch (%d CTAs)" % (multi, sel.sum()))
        for k in range(3, 11):
            d = (tl[sel, k] - tl[sel, 2]) / 1.965e3
            print("      %-16s mean %6.2f us  min %6.2f  max %6.2f" % (names[k], d.mean(), d.min(), d.max()))
        d = (tl[sel, 12] - tl[sel, 1]) / 1e3
        print("      CTA lifetime (globaltimer) mean %.2f us max %.2f;  start->end of launch window: CTA start rel. prev end mean %.2f"
              % (d.mean(), d.max(), ((tl[sel, 1] - prev_end) / 1e3).mean()))
