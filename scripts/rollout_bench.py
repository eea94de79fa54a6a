"""Inference rollout (karman_apply.py:138-158, 500 frames at the reference's 64x32 apply grid): frames/s of the reference's own
eager loop on the CUDA engine (phi_compat) and of the same rollout in one library call (sol_unroll_rollout)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from solver_in_the_loop_b200 import engine  # noqa: E402
from solver_in_the_loop_b200.phi_compat import (OPEN, CorrectionModel, Domain, Fluid, KarmanFlow, StaggeredGrid, box, to_feature,  # noqa: E402
                                                to_staggered, unstack_staggered_tensor)

torch.cuda.set_device(0)
res, L, Re, n = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 100, 1.0e6, 500
st = Fluid(Domain(resolution=[res * 2, res], box=box[0:L * 2, 0:L], boundaries=OPEN), buoyancy_factor=0)
vn = st.velocity.staggered_tensor()
vn[..., 0] = 1.0
vn[..., vn.shape[1] // 2 + 10:vn.shape[1] // 2 + 20, vn.shape[2] // 2 - 2:vn.shape[2] // 2 + 2, 1] = 1.0
st0 = st.copied_with(velocity=StaggeredGrid(unstack_staggered_tensor(vn), st.velocity.box))
bc = np.zeros(tuple(st0.velocity.data[0].data.shape))
bc[..., 0:2, 0:bc.shape[2] - 1, 0] = 1.0
bc[..., 0:bc.shape[1], 0:1, 0] = 1.0
bc[..., 0:bc.shape[1], -1:, 0] = 1.0
model = CorrectionModel(seed=0)
model.flat.mul_(0.05)
std_in = torch.tensor([0.4, 0.1, 1.7e6], device="cuda"); std_out = std_in[:2]
sim = KarmanFlow()


def eager():
    s = st0
    for _ in range(n):
        s = sim.step(s, re=Re, res=res, velBCy=bc, velBCyMask=bc)
        cv = to_staggered(model.predict(to_feature([s], Re) / std_in) * std_out, s.velocity.box)
        s = s.copied_with(velocity=s.velocity + cv)
    return s


plan = sim._plan(st0, bc, bc)
un = engine.Unroll(plan, 1, 1, (0.4, 0.1, 1.7e6), with_density=True)
re_t = torch.full((1,), Re, device="cuda")


def fused():
    return un.rollout(model.flat, re_t, st0.velocity._vy.contiguous(), st0.velocity._vx.contiguous(), n, rho0=st0.density._t.contiguous())


for name, fn in (("eager phi_compat loop", eager), ("sol_unroll_rollout", fused)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-24s %dx%d: %d frames in %.1f ms = %.0f frames/s" % (name, 2 * res, res, n, dt * 1e3, n / dt), flush=True)
a = eager().velocity._vy; b = fused()[0][-1]
print("final-frame difference eager vs fused: %.2e" % float((a - b).norm() / a.norm()))
