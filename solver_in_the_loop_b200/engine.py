"""Torch-facing wrappers of the C ABI: device memory and streams come from PyTorch (plumbing),
every computation is a kernel of libsol_b200.so.  Tensors are fp32, contiguous, CUDA:

    vy [B, Y+1, X]   vx [B, Y, X+1]   rho / p / div [B, Y, X]   CNN tensors NHWC

Reference call sites mirrored here: KarmanFlow.step (karman-2d/karman_train.py:173-185), the CNN
(:101-138), the msteps unroll + loss + optimiser (:393-457).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import SolError, UnrollCfg, check

LEAKY_ALPHA = 0.3


def _ptr(t: Optional[torch.Tensor], dtype=torch.float32) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda:
        raise SolError("libsol_b200 operates on CUDA tensors only (no CPU fallback)")
    if t.dtype != dtype or not t.is_contiguous():
        raise SolError("expected a contiguous %s tensor, got %s contiguous=%s" % (dtype, t.dtype, t.is_contiguous()))
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _host(a, dtype):
    if a is None:
        return None, None
    arr = np.ascontiguousarray(np.asarray(a), dtype=dtype)
    return arr, arr.ctypes.data_as(C.c_void_p)


def set_option(name: str, value: int):
    """Process-wide knob, e.g. set_option("conv_path", 2) selects the tcgen05 convolution kernels."""
    check(_lib.load().sol_set_option(name.encode(), int(value)))


class Plan:
    """Scene geometry (reference: KarmanFlow.__init__, Domain/Fluid set-up, velBCy/velBCyMask)."""

    def __init__(self, Y: int, X: int, B_max: int, dx: float, boundary: int = _lib.SOL_BOUNDARY_OPEN,
                 solid=None, inflow=None, bc_mask_y=None, bc_val_y=None, device: Optional[torch.device] = None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise SolError("no CUDA device: the solver-in-the-loop engine has no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.Y, self.X, self.B_max, self.dx, self.boundary = int(Y), int(X), int(B_max), float(dx), int(boundary)
        keep = [_host(solid, np.uint8), _host(inflow, np.float32), _host(bc_mask_y, np.float32), _host(bc_val_y, np.float32)]
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.sol_plan_create(self.Y, self.X, self.B_max, self.dx, self.boundary,
                                           keep[0][1], keep[1][1], keep[2][1], keep[3][1], C.byref(h)))
        self.handle = h
        self.NY, self.NX, self.NC = (self.Y + 1) * self.X, self.Y * (self.X + 1), self.Y * self.X

    @classmethod
    def karman(cls, Y: int, X: int, B_max: int, L: float = 100.0, device=None) -> "Plan":
        """The karman-2d scene (karman_train.py:166-171, 363-372): Sphere([50,50],10) obstacle,
        Inflow box[5:10,25:75], free-stream BC on the y component; masks built in index space."""
        dx = L / X
        cy = (np.arange(Y) + 0.5) * dx
        cx = (np.arange(X) + 0.5) * dx
        CY, CX = np.meshgrid(cy, cx, indexing="ij")
        solid = ((CY - 50.0) ** 2 + (CX - 50.0) ** 2) <= 10.0 ** 2
        inflow = ((CY >= 5.0) & (CY <= 10.0) & (CX >= 25.0) & (CX <= 75.0)).astype(np.float32)
        vn = np.zeros((Y + 1, X), dtype=np.float32)
        vn[0:2, 0:X - 1] = 1.0
        vn[:, 0:1] = 1.0
        vn[:, -1:] = 1.0
        return cls(Y, X, B_max, dx, _lib.SOL_BOUNDARY_OPEN, solid, inflow, vn, vn.copy(), device)

    @classmethod
    def periodic(cls, Y: int, X: int, B_max: int, dx: float, device=None) -> "Plan":
        return cls(Y, X, B_max, dx, _lib.SOL_BOUNDARY_PERIODIC, device=device)

    def set_cg(self, tol_abs: float = 1e-5, tol_rel: float = 0.0, max_it: int = 2000, cluster: int = 0):
        check(self.lib.sol_plan_set_cg(self.handle, tol_abs, tol_rel, max_it, cluster))

    def set_option(self, name: str, value: int):
        check(self.lib.sol_plan_set_option(self.handle, name.encode(), int(value)))

    def query(self, name: str) -> int:
        v = C.c_int()
        check(self.lib.sol_plan_query(self.handle, name.encode(), C.byref(v)))
        return v.value

    def direct_rows(self) -> int:
        """k of the capacitance matrix of the direct projection (0 when the iterative solvers are in use)."""
        return self.query("direct_rows")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.sol_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- allocation helpers -------------------------------------------------------------------
    def faces(self, B: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return (torch.empty(B, self.Y + 1, self.X, device=self.device), torch.empty(B, self.Y, self.X + 1, device=self.device))

    def cells(self, B: int, C_: int = 0) -> torch.Tensor:
        shape = (B, self.Y, self.X) + ((C_,) if C_ else ())
        return torch.empty(*shape, device=self.device)

    # ---- stages ---------------------------------------------------------------------------------
    def diffuse_bc(self, re, vy, vx, dt=1.0, res=None):
        B = vy.shape[0]
        oy, ox = self.faces(B)
        check(self.lib.sol_diffuse_bc(self.handle, _stream(), B, _ptr(re), dt, float(self.X if res is None else res),
                                      _ptr(vy), _ptr(vx), _ptr(oy), _ptr(ox)))
        return oy, ox

    def diffuse_bc_bwd(self, re, gy, gx, dt=1.0, res=None):
        B = gy.shape[0]
        oy, ox = self.faces(B)
        check(self.lib.sol_diffuse_bc_bwd(self.handle, _stream(), B, _ptr(re), dt, float(self.X if res is None else res),
                                          _ptr(gy), _ptr(gx), _ptr(oy), _ptr(ox)))
        return oy, ox

    def diffuse_advect(self, re, vy, vx, rho=None, dt=1.0, res=None):
        """diffuse_bc + advect in one launch (shared-memory staged halo): returns (vy1, vx1, vy2, vx2[, rho_out])."""
        B = vy.shape[0]
        y1, x1 = self.faces(B); y2, x2 = self.faces(B)
        orho = self.cells(B) if rho is not None else None
        check(self.lib.sol_diffuse_advect(self.handle, _stream(), B, _ptr(re), dt, float(self.X if res is None else res), _ptr(vy), _ptr(vx),
                                          _ptr(rho), _ptr(y1), _ptr(x1), _ptr(y2), _ptr(x2), _ptr(orho)))
        return (y1, x1, y2, x2, orho) if rho is not None else (y1, x1, y2, x2)

    def advect(self, vy, vx, rho=None, dt=1.0):
        B = vy.shape[0]
        oy, ox = self.faces(B)
        orho = self.cells(B) if rho is not None else None
        check(self.lib.sol_advect(self.handle, _stream(), B, dt, _ptr(vy), _ptr(vx), _ptr(rho), _ptr(oy), _ptr(ox), _ptr(orho)))
        return (oy, ox, orho) if rho is not None else (oy, ox)

    def advect_bwd(self, vy, vx, gy_out, gx_out, dt=1.0):
        B = vy.shape[0]
        gy, gx = self.faces(B)
        check(self.lib.sol_advect_bwd(self.handle, _stream(), B, dt, _ptr(vy), _ptr(vx), _ptr(gy_out), _ptr(gx_out), _ptr(gy), _ptr(gx)))
        return gy, gx

    def divergence(self, vy, vx):
        B = vy.shape[0]
        d = self.cells(B)
        check(self.lib.sol_divergence(self.handle, _stream(), B, _ptr(vy), _ptr(vx), _ptr(d)))
        return d

    def pressure_solve(self, div):
        """PoissonSolver.solve plug-in slot: returns (pressure, iterations[B])."""
        B = div.shape[0]
        p = self.cells(B)
        it = torch.zeros(B, dtype=torch.int32, device=self.device)
        check(self.lib.sol_pressure_solve(self.handle, _stream(), B, _ptr(div), _ptr(p), _ptr(it, torch.int32)))
        return p, it

    def project(self, vy, vx, return_pressure=False):
        """divergence_free(): returns (vy, vx[, p], iterations)."""
        B = vy.shape[0]
        oy, ox = self.faces(B)
        p = self.cells(B) if return_pressure else None
        it = torch.zeros(B, dtype=torch.int32, device=self.device)
        check(self.lib.sol_project(self.handle, _stream(), B, _ptr(vy), _ptr(vx), _ptr(oy), _ptr(ox), _ptr(p), _ptr(it, torch.int32)))
        return (oy, ox, p, it) if return_pressure else (oy, ox, it)

    def step_fwd(self, re, vy, vx, rho=None, dt=1.0, res=None, return_aux=False):
        """One KarmanFlow.step.  Returns dict(vy, vx[, rho], p, vy1, vx1, iters)."""
        B = vy.shape[0]
        oy, ox = self.faces(B); y1, x1 = self.faces(B); sy, sx = self.faces(B)
        orho = self.cells(B) if rho is not None else None
        p = self.cells(B)
        it = torch.zeros(B, dtype=torch.int32, device=self.device)
        check(self.lib.sol_step_fwd(self.handle, _stream(), B, _ptr(re), dt, float(self.X if res is None else res),
                                    _ptr(rho), _ptr(vy), _ptr(vx), _ptr(orho), _ptr(oy), _ptr(ox), _ptr(p),
                                    _ptr(y1), _ptr(x1), _ptr(sy), _ptr(sx), _ptr(it, torch.int32)))
        return dict(vy=oy, vx=ox, rho=orho, p=p, vy1=y1, vx1=x1, iters=it)

    def step_bwd(self, re, vy1, vx1, gy_out, gx_out, dt=1.0, res=None):
        B = vy1.shape[0]
        gy, gx = self.faces(B); sy, sx = self.faces(B)
        it = torch.zeros(B, dtype=torch.int32, device=self.device)
        check(self.lib.sol_step_bwd(self.handle, _stream(), B, _ptr(re), dt, float(self.X if res is None else res),
                                    _ptr(vy1), _ptr(vx1), _ptr(gy_out), _ptr(gx_out), _ptr(gy), _ptr(gx), _ptr(sy), _ptr(sx),
                                    _ptr(it, torch.int32)))
        return gy, gx, it

    def burgers_step(self, vy, vx, dt, viscosity=0.1, ky=None, kx=None, fy=None, fx=None):
        B = vy.shape[0]
        oy, ox = self.faces(B); sy, sx = self.faces(B)
        check(self.lib.sol_burgers_step(self.handle, _stream(), B, dt, viscosity, _ptr(ky), _ptr(kx), _ptr(vy), _ptr(vx),
                                        _ptr(fy), _ptr(fx), _ptr(oy), _ptr(ox), _ptr(sy), _ptr(sx)))
        return oy, ox

    def burgers_step_bwd(self, vy, vx, gy_out, gx_out, dt, viscosity=0.1, ky=None, kx=None):
        B = vy.shape[0]
        gy, gx = self.faces(B); sy, sx = self.faces(B)
        check(self.lib.sol_burgers_step_bwd(self.handle, _stream(), B, dt, viscosity, _ptr(ky), _ptr(kx), _ptr(vy), _ptr(vx),
                                            _ptr(gy_out), _ptr(gx_out), _ptr(gy), _ptr(gx), _ptr(sy), _ptr(sx)))
        return gy, gx

    def to_feature(self, vy, vx, re, sig):
        B = vy.shape[0]
        f = self.cells(B, 3)
        check(self.lib.sol_to_feature(self.handle, _stream(), B, _ptr(vy), _ptr(vx), _ptr(re), sig[0], sig[1], sig[2], _ptr(f)))
        return f


# ---- convolutions (no plan needed) ---------------------------------------------------------------
def conv5x5(x: torch.Tensor, w: torch.Tensor, bias=None, addend=None, ref=None, act=_lib.SOL_ACT_NONE, slope=LEAKY_ALPHA):
    """x [B,Y,X,Cin] NHWC, w [5,5,Cin,Cout] (Keras) -> [B,Y,X,Cout]."""
    lib = _lib.load()
    B, Y, X, Cin = x.shape
    Cout = w.shape[-1]
    out = torch.empty(B, Y, X, Cout, device=x.device)
    check(lib.sol_conv5x5(_stream(), B, Y, X, Cin, Cout, _ptr(x), _ptr(w), _ptr(bias), _ptr(addend), _ptr(ref), act, slope, _ptr(out)))
    return out


def conv5x5_split_weights(w: torch.Tensor) -> torch.Tensor:
    """tf32 hi/lo operand layout of a [5,5,32,32] weight tensor for conv5x5_c32_presplit (tensor-core path)."""
    lib = _lib.load()
    assert tuple(w.shape) == (5, 5, 32, 32)
    ws = torch.empty(lib.sol_conv5x5_split_floats(), device=w.device)
    check(lib.sol_conv5x5_split_weights(_stream(), _ptr(w), _ptr(ws)))
    return ws


def conv5x5_c32_presplit(x: torch.Tensor, wsplit: torch.Tensor, bias=None, addend=None, ref=None, act=_lib.SOL_ACT_NONE,
                         slope=LEAKY_ALPHA, out=None, weights_settled=False):
    """32->32 layer on the tensor cores with weights split once by conv5x5_split_weights.  weights_settled=True
    promises that wsplit was complete before the previous kernel on this stream was launched."""
    lib = _lib.load()
    B, Y, X, Cin = x.shape
    assert Cin == 32
    if out is None:
        out = torch.empty(B, Y, X, 32, device=x.device)
    check(lib.sol_conv5x5_c32_presplit(_stream(), B, Y, X, _ptr(x), _ptr(wsplit), _ptr(bias), _ptr(addend), _ptr(ref), act, slope, _ptr(out),
                                       1 if weights_settled else 0))
    return out


def conv5x5_flip_weights(w: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    Cin, Cout = w.shape[2], w.shape[3]
    wT = torch.empty(5, 5, Cout, Cin, device=w.device)
    check(lib.sol_conv5x5_flip_weights(_stream(), Cin, Cout, _ptr(w), _ptr(wT)))
    return wT


def conv5x5_wgrad(x: torch.Tensor, g: torch.Tensor, accumulate_into: Optional[Tuple[torch.Tensor, torch.Tensor]] = None):
    lib = _lib.load()
    B, Y, X, Cin = x.shape
    Cout = g.shape[-1]
    if accumulate_into is None:
        dW = torch.empty(5, 5, Cin, Cout, device=x.device); db = torch.empty(Cout, device=x.device); acc = 0
    else:
        dW, db = accumulate_into; acc = 1
    n = lib.sol_conv5x5_wgrad_workspace(Cin, Cout)
    part = torch.empty(n, device=x.device) if n else None
    check(lib.sol_conv5x5_wgrad(_stream(), B, Y, X, Cin, Cout, _ptr(x), _ptr(g), _ptr(dW), _ptr(db), acc, _ptr(part)))
    return dW, db


def model_param_count(model=_lib.SOL_MODEL_MARS_MOON, cin0=3) -> int:
    return int(_lib.load().sol_model_param_count(model, cin0))


def adam_tf1(theta, grad, m, v, t: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    check(_lib.load().sol_adam_tf1(_stream(), theta.numel(), _ptr(theta), _ptr(grad), _ptr(m), _ptr(v), int(t), lr, beta1, beta2, eps,
                                   grad_scale))


class Unroll:
    """The msteps-unrolled training iteration (karman_train.py:393-457, sess.run at :502)."""

    def __init__(self, plan: Plan, msteps: int, B: int, sig: Sequence[float], dt: float = 1.0, res: Optional[float] = None,
                 model: int = _lib.SOL_MODEL_MARS_MOON, cin0: int = 3, with_density: bool = False, use_graph: bool = False):
        self.plan, self.lib = plan, plan.lib
        self.msteps, self.B = int(msteps), int(B)
        self.cfg = UnrollCfg(model, cin0, self.msteps, self.B, dt, float(plan.X if res is None else res),
                             float(sig[0]), float(sig[1]), float(sig[2]), int(with_density), int(use_graph))
        nbytes = self.lib.sol_unroll_workspace_bytes(plan.handle, C.byref(self.cfg))
        if nbytes == 0:
            raise SolError("sol_unroll_workspace_bytes: %s" % self.lib.sol_last_error_string().decode())
        self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=plan.device)
        base = self.workspace.data_ptr()
        self._ws_ptr = (base + 255) // 256 * 256
        h = C.c_void_p()
        check(self.lib.sol_unroll_create(plan.handle, C.byref(self.cfg), self._ws_ptr, nbytes, C.byref(h)))
        self.handle = h
        self.nparams = model_param_count(model, cin0)
        self.loss_steps = torch.zeros(self.msteps, device=plan.device)

    def set_burgers(self, viscosity: float = 0.1, ky=None, kx=None, f_vy=None, f_vx=None, sig_f: Sequence[float] = (1.0, 1.0)):
        """Burgers scene of an unroll on a periodic plan (burgers/burgers_train.py:379-437): diffusion kernels as in
        Plan.burgers_step, per-step forces f_vy [m,B,Y+1,X] / f_vx [m,B,Y,X+1] (None with cin0 = 2: --noforce)."""
        self._burgers_keep = (ky, kx, f_vy, f_vx)       # the library keeps the raw pointers
        check(self.lib.sol_unroll_set_burgers(self.handle, float(viscosity), _ptr(ky), _ptr(kx), _ptr(f_vy), _ptr(f_vx),
                                              float(sig_f[0]), float(sig_f[1])))

    def close(self):
        if getattr(self, "handle", None):
            self.lib.sol_unroll_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def forward(self, weights, re, vy0, vx0, gt_vy=None, gt_vx=None, rho0=None, return_pred=False):
        pv = px = pr = None
        if return_pred:
            pv = torch.empty(self.msteps, self.B, self.plan.Y + 1, self.plan.X, device=self.plan.device)
            px = torch.empty(self.msteps, self.B, self.plan.Y, self.plan.X + 1, device=self.plan.device)
            if rho0 is not None and self.cfg.with_density:
                pr = torch.empty(self.msteps, self.B, self.plan.Y, self.plan.X, device=self.plan.device)
        check(self.lib.sol_unroll_forward(self.handle, _stream(), _ptr(weights), _ptr(re), _ptr(rho0), _ptr(vy0), _ptr(vx0),
                                          _ptr(gt_vy), _ptr(gt_vx), _ptr(self.loss_steps), _ptr(pv), _ptr(px), _ptr(pr)))
        return (self.loss_steps, pv, px, pr) if return_pred else self.loss_steps

    def rollout(self, weights, re, vy0, vx0, nsteps: int, rho0=None):
        """Forward-only rollout of nsteps corrected frames (karman_apply.py:138-151) in ONE library call: returns
        (vy [n,B,Y+1,X], vx [n,B,Y,X+1], rho [n,B,Y,X] or None)."""
        dev, p = self.plan.device, self.plan
        pv = torch.empty(nsteps, self.B, p.Y + 1, p.X, device=dev)
        px = torch.empty(nsteps, self.B, p.Y, p.X + 1, device=dev)
        pr = torch.empty(nsteps, self.B, p.Y, p.X, device=dev) if (rho0 is not None and self.cfg.with_density) else None
        check(self.lib.sol_unroll_rollout(self.handle, _stream(), _ptr(weights), _ptr(re), _ptr(rho0), _ptr(vy0), _ptr(vx0), int(nsteps),
                                          _ptr(pv), _ptr(px), _ptr(pr)))
        return pv, px, pr

    def backward(self, weights, grad_out=None, want_input_grad=False):
        g = torch.empty(self.nparams, device=self.plan.device) if grad_out is None else grad_out
        gy = gx = None
        if want_input_grad:
            gy, gx = self.plan.faces(self.B)
        check(self.lib.sol_unroll_backward(self.handle, _stream(), _ptr(weights), _ptr(g), _ptr(gy), _ptr(gx)))
        return (g, gy, gx) if want_input_grad else g

    def train_iter(self, weights, re, vy0, vx0, gt_vy, gt_vx, grad_out, rho0=None):
        check(self.lib.sol_unroll_train_iter(self.handle, _stream(), _ptr(weights), _ptr(re), _ptr(rho0), _ptr(vy0), _ptr(vx0),
                                             _ptr(gt_vy), _ptr(gt_vx), _ptr(self.loss_steps), _ptr(grad_out)))
        return self.loss_steps

    def cg_iters(self) -> torch.Tensor:
        p = C.c_void_p(); n = C.c_int()
        check(self.lib.sol_unroll_cg_iters(self.handle, C.byref(p), C.byref(n)))
        out = torch.empty(n.value, dtype=torch.int32, device=self.plan.device)
        # device-to-device copy of the iteration counters out of the workspace
        off = p.value - self.workspace.data_ptr()
        out.copy_(self.workspace[off:off + 4 * n.value].view(torch.int32))
        return out.view(2, self.msteps, self.B)
