"""Phase timeline of the tcgen05 convolution kernel (clock64 stamps per CTA, diagnostics hook)."""
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
lib.sol_debug_conv_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
lib.sol_debug_conv_trace.restype = None
B, Y, X = (int(sys.argv[1]) if len(sys.argv) > 1 else 3), 128, 64
x = torch.randn(B, Y, X, 32, device=dev); w = torch.randn(5, 5, 32, 32, device=dev) * 0.05; b = torch.randn(32, device=dev)
engine.set_option("conv_path", 2)
nct = (X // 8) * (Y // 16) * B
for _ in range(5):
    y = engine.conv5x5(x, w, b, act=1)
torch.cuda.synchronize()
tr = torch.zeros(nct, 16, dtype=torch.int64, device=dev)
lib.sol_debug_conv_trace(ctypes.c_void_p(tr.data_ptr()), 1)
x2 = engine.conv5x5(y, w, b, act=1)     # chained like the network
torch.cuda.synchronize()
lib.sol_debug_conv_trace(None, 0)
t = tr.cpu().numpy()
names = {2: "start", 3: "setup", 4: "halo landed", 5: "split done", 6: "mma may start", 7: "first weights", 8: "mmas issued",
         9: "acc complete", 10: "stores issued"}
gt0 = t[:, 1].min()
print("CTAs %d, SMs used %d, CTAs/SM max %d" % (nct, len(set(t[:, 0])), np.bincount(t[:, 0]).max()))
print("kernel span by globaltimer (start of first CTA -> start of last): %.2f us" % ((t[:, 1].max() - gt0) / 1e3))
for multi in (False, True):
    cnt = np.bincount(t[:, 0], minlength=200)
    sel = np.array([(cnt[s] > 1) == multi for s in t[:, 0]])
    if not sel.any():
        continue
    print("== SMs with %s CTA (%d CTAs)" % ("2" if multi else "1", sel.sum()))
    for k in range(3, 11):
        d = (t[sel, k] - t[sel, 2]) / 1.965e3
        print("  %-16s mean %6.2f us  min %6.2f  max %6.2f" % (names[k], d.mean(), d.min(), d.max()))
