"""Timeline of the tensor-core convolution launches inside the replayed CUDA graph of one training iteration:
per launch start/end (globaltimer) -> kernel durations and the gaps between consecutive launches."""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from solver_in_the_loop_b200 import _lib, engine  # noqa: E402
from solver_in_the_loop_b200.trainer import SolTrainer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--msteps", type=int, default=4)
ap.add_argument("--pdl", type=int, default=1)
ap.add_argument("--graph", type=int, default=1)
ap.add_argument("--chain", type=int, default=1)
a = ap.parse_args()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.sol_debug_conv_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.sol_debug_conv_trace.restype = None
engine.set_option("pdl", a.pdl)
engine.set_option("conv_chain", a.chain)
B, Y, X, m = 3, 128, 64, a.msteps
plan = engine.Plan.karman(Y, X, B)
plan.set_cg(1e-7, 1e-6, 4000, 0)
re, vy0, vx0, gy, gx, sig = bench.synth_batch(plan, engine, torch, B, m, 0, 30)
plan.set_cg(1e-5, 0.0, 2000, 0)
nl = 20 * m
nct = (X // 8) * (Y // 16) * B
tr = torch.zeros(nl, nct, 16, dtype=torch.int64, device=dev)
lib.sol_debug_conv_trace(ctypes.c_void_p(tr.data_ptr()), nl)     # before capture: the pointers are baked into the graph
t = SolTrainer(plan, m, B, sig, use_graph=bool(a.graph))
t.weights.mul_(0.1)
for _ in range(4):
    t.train_step(re, vy0, vx0, gy, gx)
torch.cuda.synchronize()
lib.sol_debug_conv_trace(None, 0)
d = tr.cpu().numpy()
start = d[:, :, 1].min(axis=1); end = d[:, :, 12].max(axis=1)
rel_first = d[:, :, 13].min(axis=1); rel_last = d[:, :, 13].max(axis=1)      # dependency resolved (after griddepcontrol.wait)
print("pdl %d chain %d graph %d: %d conv launches" % (a.pdl, a.chain, a.graph, nl))
period = (end[1:] - end[:-1]) / 1e3
adj = period < 22.0              # directly consecutive conv layers (other kernels in between give longer periods)
print("end(i+1) - end(i) between directly consecutive conv launches: n=%d mean %.2f us  min %.2f  max %.2f" %
      (adj.sum(), period[adj].mean(), period[adj].min(), period[adj].max()))
g1 = ((rel_first[1:] - end[:-1]) / 1e3)[adj]; g2 = ((rel_last[1:] - end[:-1]) / 1e3)[adj]
print("last CTA end of launch i -> first / last CTA of launch i+1 released: mean %.2f / %.2f us" % (g1.mean(), g2.mean()))
work = ((end - rel_first) / 1e3)
print("first release -> last CTA end (useful span of a launch): mean %.2f us" % work.mean())
print("CTA start -> release (time parked in griddepcontrol.wait): mean %.2f us" % (((d[:, :, 13] - d[:, :, 1]) / 1e3).mean()))
# phase stamps relative to CTA start (clock64 at 1.965 GHz assumed)
names = {3: "setup", 4: "halo landed", 5: "split done", 7: "first weights", 8: "mmas issued", 9: "acc complete", 10: "stores issued"}
for k, nme in names.items():
    v = (d[:, :, k] - d[:, :, 2]) / 1.965e3
    print("  %-14s mean %6.2f us  p90 %6.2f" % (nme, v.mean(), np.percentile(v, 90)))
