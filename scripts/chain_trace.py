#!/usr/bin/env python
"""Where the time of ONE graph-replayed training iteration goes, kernel by kernel, without a profiler in the way.

Thread 0 of CTA 0 of every kernel stamps %globaltimer right after its griddepcontrol.wait (sol_debug_chain_trace), i.e. when
its predecessor in the stream has completed.  The kernels of an iteration form one serial chain, so stamp[i+1] - stamp[i] is
the in-chain cost of kernel i (its own critical path + the programmatic hand-over to the next kernel).  Names come from the
host-side launch order of the eager first call (same order as the captured graph).

    python scripts/chain_trace.py [--config sol32|c2|c4] [--msteps M] [--batch B] > profiles/rNN_chain_trace.txt
"""
import argparse
import collections
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from solver_in_the_loop_b200 import _lib, engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--Y", type=int, default=128)
ap.add_argument("--X", type=int, default=64)
ap.add_argument("--batch", type=int, default=3)
ap.add_argument("--msteps", type=int, default=32)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--opt", action="append", default=[], help="name=value engine option")
a = ap.parse_args()
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
lib = _lib.load()
lib.sol_debug_chain_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint]
lib.sol_debug_chain_trace.restype = None
lib.sol_debug_chain_names.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_int]
lib.sol_debug_chain_names.restype = ctypes.c_int
for o in a.opt:
    k, v = o.split("=")
    engine.set_option(k, int(v))
Y, X, B, m = a.Y, a.X, a.batch, a.msteps
plan = engine.Plan.karman(Y, X, B)
plan.set_cg(1e-7, 1e-6, 4000, 0)
re, vy0, vx0, gy, gx, sig = bench.synth_batch(plan, engine, torch, B, m, 0, 30)
plan.set_cg(1e-5, 0.0, 2000, 0)
un = engine.Unroll(plan, m, B, sig, use_graph=True)
w = torch.randn(un.nparams, device=dev) * 0.005
g = torch.zeros(un.nparams, device=dev)
lib.sol_debug_chain_names(1, None, 0)
un.train_iter(w, re, vy0, vx0, gy, gx, g)          # eager first call: the launch order is recorded
torch.cuda.synchronize()
buf = ctypes.create_string_buffer(1 << 22)
n_launch = lib.sol_debug_chain_names(0, buf, len(buf))
names = buf.value.decode().split("\n")[:-1]
if any(nm.startswith("_Z") for nm in names):        # cudaFuncGetName returns the mangled symbol
    import subprocess
    uniq = sorted(set(names))
    dem = subprocess.run(["c++filt"], input="\n".join(uniq), capture_output=True, text=True).stdout.split("\n")
    names = [dict(zip(uniq, dem))[nm] for nm in names]
for _ in range(3):                                  # capture + replays
    un.train_iter(w, re, vy0, vx0, gy, gx, g)
torch.cuda.synchronize()
cap = n_launch + 64
stamps = torch.zeros(cap, dtype=torch.int64, device=dev)
count = torch.zeros(1, dtype=torch.int32, device=dev)
agg = collections.OrderedDict()
tot = []
for rep in range(a.reps):
    count.zero_(); stamps.zero_()
    lib.sol_debug_chain_trace(ctypes.c_void_p(stamps.data_ptr()), ctypes.c_void_p(count.data_ptr()), cap)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    un.train_iter(w, re, vy0, vx0, gy, gx, g)
    e1.record()
    torch.cuda.synchronize()
    lib.sol_debug_chain_trace(None, None, 0)
    n = int(count.item())
    t = stamps[:n].cpu().numpy().astype(np.int64)
    if n != n_launch:
        print("# rep %d: %d stamps for %d launches (kernels without a griddepcontrol.wait?)" % (rep, n, n_launch))
        continue
    order = np.argsort(t, kind="stable")
    if not (order == np.arange(n)).all():
        print("# rep %d: %d stamps out of launch order (concurrent streams)" % (rep, int((order != np.arange(n)).sum())))
    dt = np.diff(t) / 1000.0                        # us; the last kernel has no successor
    tot.append((e0.elapsed_time(e1), (t[-1] - t[0]) / 1000.0))
    if rep == 0:
        continue                                    # first traced replay: cold constants
    for i in range(n - 1):
        nm = names[i].replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "").replace("sol::", "")
        agg.setdefault(nm, []).append(dt[i])
print("one iteration %dx%d B=%d msteps=%d: %d launches; CUDA-event ms / first-to-last-stamp ms per traced replay: %s"
      % (Y, X, B, m, n_launch, " ".join("%.3f/%.3f" % (x, y / 1000.0) for x, y in tot)))
reps = max(1, len(tot) - 1)
rows = sorted(agg.items(), key=lambda kv: -sum(kv[1]))
total = sum(sum(v) for _, v in rows) / reps
print("%-52s %6s %10s %9s %7s" % ("kernel (stamp-to-next-stamp = in-chain cost)", "n/iter", "us/launch", "ms/iter", "share"))
for nm, v in rows:
    print("%-52s %6d %10.2f %9.3f %6.1f%%" % (nm[:52], len(v) // reps, float(np.mean(v)), sum(v) / reps / 1000.0, 100.0 * sum(v) / reps / total))
print("%-52s %6s %10s %9.3f" % ("sum", "", "", total / 1000.0))
