"""Generates tests/golden/karman_32x32_m2.npz and tests/golden/burgers_32x32_m2.npz from the CPU oracle (float64).

The reference itself cannot run offline (PhiFlow/TensorFlow absent), so these are
REGRESSION PINS of the oracle — they keep the restated semantics from drifting — not reference
outputs.  Re-run only when a `Switches` default is deliberately changed:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sol_oracle as so  # noqa: E402


def main():
    Y, X, B, m = 32, 32, 1, 2
    geom, rho, vy, vx, re, gty, gtx, sig = so.make_case(Y=Y, X=X, B=B, msteps=m, spin=10, seed=0)
    params = [p.requires_grad_() for p in so.init_params(seed=0)]
    loss, losses, states = so.unrolled_loss(params, rho, vy, vx, re, gty, gtx, geom, sig, m, return_states=True)
    loss.backward()
    r1, y1, x1, aux = so.karman_step(rho, vy, vx, re, geom, return_aux=True)
    out = dict(
        Y=Y, X=X, B=B, m=m, sig=np.array(sig), re=re.numpy(), rho0=rho.numpy(), vy0=vy.numpy(), vx0=vx.numpy(),
        gt_vy=gty.numpy(), gt_vx=gtx.numpy(),
        step_vy=y1.numpy(), step_vx=x1.numpy(), step_rho=r1.numpy(), step_p=aux["p"].numpy(), step_div=aux["d"].numpy(),
        losses=np.array([float(l) for l in losses]),
        pred_vy=np.stack([s[1].detach().numpy() for s in states]), pred_vx=np.stack([s[2].detach().numpy() for s in states]),
        grad_b0=params[1].grad.numpy(), grad_w11=params[22].grad.numpy(), grad_b11=params[23].grad.numpy(),
        grad_w5_slice=params[10].grad[:, :, :4, :4].numpy(),
    )
    out = {k: (v.astype(np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 and k not in ("losses", "sig", "re") else v)
           for k, v in out.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "karman_32x32_m2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    burgers()


def burgers():
    """BASELINE config 1 grid (burgers 32x32, dt = 0.1, forced; burgers/Makefile:75-77), msteps = 2, batch 1."""
    R, B, m, dt, nu = 32, 1, 2, 0.1, 0.1
    dx, vy, vx, fy, fx, gty, gtx, sv, sf = so.make_burgers_case(R=R, B=B, msteps=m, dt=dt, seed=0)
    params = [(0.3 * p).requires_grad_() for p in so.init_params(cin0=4, seed=0)]
    loss, losses, states = so.burgers_unrolled_loss(params, vy, vx, fy, fx, gty, gtx, dx, dt, sv, sf, m, nu, return_states=True)
    loss.backward()
    y1, x1 = so.burgers_step(vy, vx, dt, dx, nu, fy[0], fx[0])
    out = dict(R=R, B=B, m=m, dt=dt, nu=nu, dx=dx, sig_v=np.array(sv), sig_f=np.array(sf), vy0=vy.numpy(), vx0=vx.numpy(), f_vy=fy.numpy(),
               f_vx=fx.numpy(), gt_vy=gty.numpy(), gt_vx=gtx.numpy(), step_vy=y1.numpy(), step_vx=x1.numpy(),
               losses=np.array([float(l) for l in losses]),
               pred_vy=np.stack([s[0].detach().numpy() for s in states]), pred_vx=np.stack([s[1].detach().numpy() for s in states]),
               grad_w0=params[0].grad.numpy(), grad_b0=params[1].grad.numpy(), grad_w11=params[22].grad.numpy())
    out = {k: (v.astype(np.float32) if isinstance(v, np.ndarray) and v.dtype == np.float64 and k not in ("losses", "sig_v", "sig_f") else v)
           for k, v in out.items()}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "burgers_32x32_m2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
