#!/usr/bin/env python
"""Dependent ping-pong chain of 32->32 conv launches (exactly like consecutive layers of the sweep) for every tensor-core
kernel variant: us per launch (CUDA events) and error vs the fp64 oracle conv.  Usage: python scripts/conv_bench.py [B Y X]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from solver_in_the_loop_b200 import engine  # noqa: E402

B, Y, X = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (3, 128, 64)
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(3)
x64 = torch.randn(B, Y, X, 32, generator=g, dtype=torch.float64)
w64 = torch.randn(5, 5, 32, 32, generator=g, dtype=torch.float64) * 0.03
b64 = torch.randn(32, generator=g, dtype=torch.float64) * 0.1
ref = torch.nn.functional.conv2d(x64.permute(0, 3, 1, 2), w64.permute(3, 2, 0, 1), b64, padding=2).permute(0, 2, 3, 1)
x = x64.float().to(dev); w = w64.float().to(dev); b = b64.float().to(dev)
for name, path, variant in (("fp16x3 merged sets", 2, 0), ("fp16x3 two sets", 2, 1), ("fp16x3 one set", 2, 2), ("tf32x3 (round 1)", 3, 0)):
    engine.set_option("conv_path", path); engine.set_option("conv_variant", variant)
    ws = engine.conv5x5_split_weights(w)
    torch.cuda.synchronize()
    o = engine.conv5x5_c32_presplit(x, ws, b, act=0, weights_settled=True)
    err = float((o.double().cpu() - ref).norm() / ref.norm())
    a0 = x.clone(); a1 = torch.empty_like(a0)
    wl = engine.conv5x5_split_weights(w * 0.3)      # contraction: the chain stays bounded
    torch.cuda.synchronize()
    for _ in range(4):
        engine.conv5x5_c32_presplit(a0, wl, b, act=1, out=a1, weights_settled=True)
        engine.conv5x5_c32_presplit(a1, wl, b, act=1, out=a0, weights_settled=True)
    n = 200
    # the chain is replayed from a CUDA graph (as in the engine): the Python / ctypes launch cost stays out of the figure
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for _ in range(n // 2):
            engine.conv5x5_c32_presplit(a0, wl, b, act=1, out=a1, weights_settled=True)
            engine.conv5x5_c32_presplit(a1, wl, b, act=1, out=a0, weights_settled=True)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for _ in range(reps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    n *= reps
    print("%-20s %dx%dx%d: %.2f us/launch   rel err vs fp64 %.2e" % (name, B, Y, X, e0.elapsed_time(e1) * 1e3 / n, err), flush=True)
engine.set_option("conv_path", 0); engine.set_option("conv_variant", 0)
