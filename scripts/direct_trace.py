#!/usr/bin/env python
"""Phase timeline of k_direct_solve (clock64 stamps of thread 0 of every CTA through the diagnostics hook)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from solver_in_the_loop_b200 import _lib, engine  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
lib = _lib.load()
lib.sol_debug_direct_trace.argtypes = [ctypes.c_void_p]
lib.sol_debug_direct_trace.restype = None
B, Y, X = 3, 128, 64
plan = engine.Plan.karman(Y, X, B)
vy = torch.randn(B, Y + 1, X, device=dev); vx = torch.randn(B, Y, X + 1, device=dev)
for _ in range(5):
    plan.project(vy, vx)
torch.cuda.synchronize()
tr = torch.zeros(B * 4, 16, dtype=torch.int64, device=dev)
lib.sol_debug_direct_trace(ctypes.c_void_p(tr.data_ptr()))
for _ in range(3):
    plan.project(vy, vx)
torch.cuda.synchronize()
lib.sol_debug_direct_trace(None)
t = tr.cpu().numpy().astype(float)
names = ["start", "constants issued", "pdl wait done", "cluster sync 1", "D pushed", "cluster sync 2", "stage 1 (Sy D)", "stage 2 (U Sx)", "stage 3 (V Sx)",
         "Z pushed + sync", "stage 4 + store"]
print("k_direct_solve phases, mean over %d CTAs, us at 1.965 GHz (cumulative / delta)" % t.shape[0])
prev = 0.0
for k in range(1, 11):
    d = ((t[:, k] - t[:, 0]) / 1965.0).mean()
    print("  %-22s %7.2f  %+6.2f" % (names[k], d, d - prev))
    prev = d
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
g = torch.cuda.CUDAGraph()
oy, ox = plan.faces(B)
it = torch.zeros(B, dtype=torch.int32, device=dev)
def proj():
    engine.check(plan.lib.sol_project(plan.handle, torch.cuda.current_stream().cuda_stream, B, vy.data_ptr(), vx.data_ptr(), oy.data_ptr(), ox.data_ptr(), None, it.data_ptr()))
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    proj()
    with torch.cuda.graph(g, stream=st):
        for _ in range(20):
            proj()
    g.replay(); st.synchronize()
    e0.record(st)
    for _ in range(5):
        g.replay()
    e1.record(st)
    st.synchronize()
print("projection (solve + apply) in a graph-replayed chain: %.2f us, changed rows kp = %d" % (e0.elapsed_time(e1) * 1e3 / 100, plan.direct_rows()))
