"""CPU ORACLE — test infrastructure only, never the product path.

A CPU restatement (torch float64 by default, float32 switch) of the reference's
unrolled solver-in-the-loop step for karman-2d and burgers, written to be the
*checker* for the CUDA kernels in ``solver_in_the_loop_b200/csrc``.  Only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.

PARITY UNPINNED.  The arithmetic of the reference's hot path lives in the
third-party packages ``phiflow==1.5.1`` (commit 4f5e678) and ``tensorflow==1.15``
(reference README.md:21-24).  Neither is vendored under /root/reference nor
installable in the build container (no network), and the reference ships no
tests, golden vectors or fixtures.  This oracle therefore
restates the *published* PhiFlow-1.5.1 algorithm as described in SURVEY.md
Appendix A and anchors on the reference's own call sites:

  karman-2d/karman_train.py:77-90    to_feature / to_staggered
  karman-2d/karman_train.py:101-138  model_mars_moon
  karman-2d/karman_train.py:166-185  KarmanFlow.step (viscosity, BC, IncompressibleFlow.step)
  karman-2d/karman_train.py:363-372  domain, velocity BC mask
  karman-2d/karman_train.py:393-457  msteps unroll, loss, Adam
  karman-2d/karman_apply.py:138-151  eager form of one step + correction
  burgers/burgers_train.py:178-187   BurgersTest.step / step_with_f

Every choice that is a recollection of PhiFlow internals (not pinned by code in
/root/reference) sits behind a named switch in :class:`Switches` so it can be
flipped when real PhiFlow output becomes available.

All fields are struct-of-arrays, matching the CUDA side:
  vy  [B, Y+1, X]   y-faces   (reference: velocity.data[0], "v first")
  vx  [B, Y, X+1]   x-faces   (reference: velocity.data[1])
  rho [B, Y, X]     cell centres
  p,d [B, Y, X]
y is the long / flow axis (karman_train.py:363: resolution=[2*res, res]).
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch
import torch.nn.functional as F

LEAKY_ALPHA = 0.3  # keras.layers.LeakyReLU() default (karman_train.py:105)


@dataclasses.dataclass
class Switches:
    """[PHI-RECALL] choices, SURVEY.md §8c."""
    inflow_after_advect: bool = True        # effects applied after advection (phiflow 1.5.1 order)
    density_extrap: str = "zero"            # 'zero' (constant extrapolation) | 'replicate'
    diffuse_units: str = "index"            # CenteredGrid(raw) has unit cells -> no 1/dx^2
    cg_batch_global_stop: bool = False      # reference SparseCG stops on the batch-global max|r|
    burgers_diffusion: str = "fft"          # periodic diffuse(): spectral decay | 'fd' explicit 5-pt
    periodic_wrap_size: str = "array"       # staggered comps wrap modulo the array size (res+1)


DEFAULT_SWITCHES = Switches()


# --------------------------------------------------------------------------------------
# geometry  (karman_train.py:166-171, 363-372)
# --------------------------------------------------------------------------------------
class KarmanGeom:
    """Masks of the karman-2d scene in index space.

    box = [0, 2L] x [0, L] (y, x), L = 100 (``--len``), resolution [Y, X] = [2*res, res]
    (karman_train.py:363).  Obstacle Sphere([50, 50], 10), Inflow box[5:10, 25:75]
    in *physical* units (karman_train.py:170-171).
    """

    def __init__(self, Y: int, X: int, L: float = 100.0, obstacle=((50.0, 50.0), 10.0),
                 inflow=((5.0, 10.0), (25.0, 75.0))):
        self.Y, self.X, self.L = int(Y), int(X), float(L)
        self.dx = self.L / self.X
        cy = (np.arange(self.Y) + 0.5) * self.dx
        cx = (np.arange(self.X) + 0.5) * self.dx
        CY, CX = np.meshgrid(cy, cx, indexing="ij")
        if obstacle is not None:
            (oy, ox), rad = obstacle
            self.solid = (((CY - oy) ** 2 + (CX - ox) ** 2) <= rad ** 2)
        else:
            self.solid = np.zeros((self.Y, self.X), dtype=bool)
        if inflow is not None:
            (y0, y1), (x0, x1) = inflow
            self.inflow = ((CY >= y0) & (CY <= y1) & (CX >= x0) & (CX <= x1)).astype(np.float64)
        else:
            self.inflow = np.zeros((self.Y, self.X))
        # velocity BC on the y-component grid [Y+1, X]  (karman_train.py:366-372) [PINNED]
        vn = np.zeros((self.Y + 1, self.X))
        vn[0:2, 0:self.X - 1] = 1.0
        vn[:, 0:1] = 1.0
        vn[:, -1:] = 1.0
        self.bc_mask_y = vn
        self.bc_val_y = vn.copy()
        # accessibility: 1 - solid inside, 1 outside the domain (OPEN); activity: 0 outside
        acc = 1.0 - self.solid.astype(np.float64)
        acc_pad = np.pad(acc, 1, constant_values=1.0)
        self.face_my = np.minimum(acc_pad[0:self.Y + 1, 1:self.X + 1], acc_pad[1:self.Y + 2, 1:self.X + 1])  # [Y+1, X]
        self.face_mx = np.minimum(acc_pad[1:self.Y + 1, 0:self.X + 1], acc_pad[1:self.Y + 1, 1:self.X + 2])  # [Y, X+1]
        self.active = acc.copy()
        diag = (acc_pad[0:self.Y, 1:self.X + 1] + acc_pad[2:self.Y + 2, 1:self.X + 1]
                + acc_pad[1:self.Y + 1, 0:self.X] + acc_pad[1:self.Y + 1, 2:self.X + 2])
        self.diag = np.maximum(diag, 1.0)  # reference clips the (negative) diagonal to <= -1
        self._lu = None
        self._A = None

    # Laplace matrix over ALL cells (solid rows decouple: -diag * p = d), SURVEY a10.
    def laplace_matrix(self) -> sp.csr_matrix:
        if self._A is not None:
            return self._A
        Y, X = self.Y, self.X
        N = Y * X
        idx = np.arange(N).reshape(Y, X)
        act = self.active
        rows, cols, vals = [], [], []

        def couple(a_idx, b_idx, a_act, b_act):
            w = (a_act * b_act).ravel()
            m = w > 0
            rows.append(a_idx.ravel()[m]); cols.append(b_idx.ravel()[m]); vals.append(w[m])
            rows.append(b_idx.ravel()[m]); cols.append(a_idx.ravel()[m]); vals.append(w[m])

        couple(idx[:-1, :], idx[1:, :], act[:-1, :], act[1:, :])
        couple(idx[:, :-1], idx[:, 1:], act[:, :-1], act[:, 1:])
        rows.append(idx.ravel()); cols.append(idx.ravel()); vals.append(-self.diag.ravel())
        A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(N, N)).tocsr()
        self._A = A
        return A

    def lu(self):
        if self._lu is None:
            self._lu = spla.splu(self.laplace_matrix().tocsc())
        return self._lu


def _t(a, dtype, device="cpu"):
    return torch.as_tensor(np.asarray(a), dtype=dtype, device=device)


# --------------------------------------------------------------------------------------
# stencils
# --------------------------------------------------------------------------------------
def lap_replicate(c: torch.Tensor) -> torch.Tensor:
    """5-point Laplace with replicate ('boundary') padding, dx = 1."""
    cp = F.pad(c.unsqueeze(1), (1, 1, 1, 1), mode="replicate").squeeze(1)
    return cp[:, 2:, 1:-1] + cp[:, :-2, 1:-1] + cp[:, 1:-1, 2:] + cp[:, 1:-1, :-2] - 4.0 * c


def diffuse_bc(vy, vx, alpha, bc_mask_y, bc_val_y):
    """karman_train.py:175-181: explicit viscosity on each component grid, then the
    Dirichlet inflow/free-stream BC on the y component."""
    a = alpha.reshape(-1, 1, 1)
    cy = vy + a * lap_replicate(vy)
    cx = vx + a * lap_replicate(vx)
    cy = cy * (1.0 - bc_mask_y) + bc_val_y
    return cy, cx


def bilerp(field: torch.Tensor, py: torch.Tensor, px: torch.Tensor, mode: str = "replicate") -> torch.Tensor:
    """Bilinear sample of field[B,H,W] at index coordinates (py, px) [B,h,w].

    Weights from the unclamped fractional part; the two integer corners are clamped
    independently (mode='replicate'), wrapped ('periodic') or contribute zero when
    outside ('zero')  (general_grid_sample_nd semantics).
    The floor has zero gradient; the gradient w.r.t. coordinates flows through the weights.
    """
    B, H, W = field.shape
    fy = torch.floor(py.detach())
    fx = torch.floor(px.detach())
    wy = py - fy
    wx = px - fx
    j0 = fy.long(); i0 = fx.long()
    j1 = j0 + 1; i1 = i0 + 1
    bidx = torch.arange(B, device=field.device).reshape(B, 1, 1).expand_as(j0)

    def fetch(j, i):
        if mode == "replicate":
            return field[bidx, j.clamp(0, H - 1), i.clamp(0, W - 1)]
        if mode == "periodic":
            return field[bidx, torch.remainder(j, H), torch.remainder(i, W)]
        if mode == "zero":
            inside = ((j >= 0) & (j <= H - 1) & (i >= 0) & (i <= W - 1)).to(field.dtype)
            return field[bidx, j.clamp(0, H - 1), i.clamp(0, W - 1)] * inside
        raise ValueError(mode)

    v00 = fetch(j0, i0); v01 = fetch(j0, i1); v10 = fetch(j1, i0); v11 = fetch(j1, i1)
    return ((1 - wy) * ((1 - wx) * v00 + wx * v01) + wy * ((1 - wx) * v10 + wx * v11))


def _grid(B, H, W, dtype, device):
    j = torch.arange(H, dtype=dtype, device=device).reshape(1, H, 1).expand(B, H, W)
    i = torch.arange(W, dtype=dtype, device=device).reshape(1, 1, W).expand(B, H, W)
    return j, i


def advect_velocity(vy, vx, s: float, mode: str = "replicate"):
    """Semi-Lagrangian self-advection of the staggered velocity.

    s = dt/dx (cells per unit velocity).  Each component is advected at its own face
    points; the other component is bilinearly interpolated there.
    """
    B, Yp1, X = vy.shape
    Y = Yp1 - 1
    dt_, dev = vy.dtype, vy.device
    # y-faces (j, i): vx index coords (row=j-1/2, col=i+1/2)
    j, i = _grid(B, Y + 1, X, dt_, dev)
    ux = bilerp(vx, j - 0.5, i + 0.5, mode)
    vy_new = bilerp(vy, j - s * vy, i - s * ux, mode)
    # x-faces (j, i): vy index coords (row=j+1/2, col=i-1/2)
    j, i = _grid(B, Y, X + 1, dt_, dev)
    uy = bilerp(vy, j + 0.5, i - 0.5, mode)
    vx_new = bilerp(vx, j - s * uy, i - s * vx, mode)
    return vy_new, vx_new


def advect_density(rho, vy, vx, s: float, extrap: str = "zero"):
    """Semi-Lagrangian advection of the cell-centred density."""
    B, Y, X = rho.shape
    j, i = _grid(B, Y, X, rho.dtype, rho.device)
    uy = 0.5 * (vy[:, :-1, :] + vy[:, 1:, :])
    ux = 0.5 * (vx[:, :, :-1] + vx[:, :, 1:])
    return bilerp(rho, j - s * uy, i - s * ux, extrap)


def divergence(vy, vx):
    """d = vy[j+1,i]-vy[j,i] + vx[j,i+1]-vx[j,i], index units."""
    return vy[:, 1:, :] - vy[:, :-1, :] + vx[:, :, 1:] - vx[:, :, :-1]


def apply_A(p, active, diag):
    """y = A p with p = 0 outside the domain, solid cells decoupled."""
    pa = p * active
    pp = F.pad(pa, (1, 1, 1, 1))
    nb = pp[:, 2:, 1:-1] + pp[:, :-2, 1:-1] + pp[:, 1:-1, 2:] + pp[:, 1:-1, :-2]
    return active * nb - diag * p


def cg_reference(d, active, diag, tol=1e-5, max_it=2000, batch_global_stop=False):
    """PhiFlow-1.5.1 SparseCG recurrences and stop rule — not differentiable.

    x=0; r=p=d; q=Ap; loop{ t=p.q; a=(p.r)/t; x+=a p; r-=a q; b=(r.q)/t; p=r-b p; q=Ap }
    while max|r| >= tol and it < max_it; divide_no_nan.
    Returns (x, iterations[B]).
    """
    B = d.shape[0]
    x = torch.zeros_like(d)
    r = d.clone(); p = d.clone()
    q = apply_A(p, active, diag)
    its = torch.zeros(B, dtype=torch.long)
    alive = torch.ones(B, dtype=torch.bool)

    def dn(a, b):
        return torch.where(b != 0, a / torch.where(b != 0, b, torch.ones_like(b)), torch.zeros_like(a))

    for it in range(max_it):
        rmax = r.abs().amax(dim=(1, 2))
        if batch_global_stop:
            alive = (rmax.max() >= tol).expand(B)
        else:
            alive = alive & (rmax >= tol)
        if not bool(alive.any()):
            break
        m = alive.to(d.dtype).reshape(B, 1, 1)
        t = (p * q).sum(dim=(1, 2)); pr = (p * r).sum(dim=(1, 2))
        a = dn(pr, t).reshape(B, 1, 1) * m
        x = x + a * p
        r = r - a * q
        b = dn((r * q).sum(dim=(1, 2)), t).reshape(B, 1, 1)
        p_new = r - b * p
        p = torch.where(m > 0, p_new, p)
        q = apply_A(p, active, diag)
        its = its + alive.long()
    return x, its


class _DirectSolve(torch.autograd.Function):
    """p = A^{-1} d by sparse LU in float64; backward = the same solve (A symmetric, no
    gradient to A) — the implicit gradient the reference's SparseCG registers."""

    @staticmethod
    def forward(ctx, d, geom: KarmanGeom):
        ctx.geom = geom
        lu = geom.lu()
        dn = d.detach().cpu().double().numpy().reshape(d.shape[0], -1)
        out = np.stack([lu.solve(row) for row in dn]).reshape(d.shape)
        return torch.as_tensor(out, dtype=d.dtype)

    @staticmethod
    def backward(ctx, g):
        lu = ctx.geom.lu()
        gn = g.detach().cpu().double().numpy().reshape(g.shape[0], -1)
        out = np.stack([lu.solve(row) for row in gn]).reshape(g.shape)
        return torch.as_tensor(out, dtype=g.dtype), None


class _CGSolve(torch.autograd.Function):
    @staticmethod
    def forward(ctx, d, active, diag, tol, max_it, glob, stats):
        ctx.save_for_backward(active, diag)
        ctx.cfg = (tol, max_it, glob, stats)
        x, its = cg_reference(d, active, diag, tol, max_it, glob)
        if stats is not None:
            stats.setdefault("fwd_iters", []).append(its.clone())
        return x

    @staticmethod
    def backward(ctx, g):
        active, diag = ctx.saved_tensors
        tol, max_it, glob, stats = ctx.cfg
        x, its = cg_reference(g.contiguous(), active, diag, tol, max_it, glob)
        if stats is not None:
            stats.setdefault("bwd_iters", []).append(its.clone())
        return x, None, None, None, None, None, None


def pressure_solve(d, geom: KarmanGeom, solver="direct", tol=1e-5, max_it=2000, stats=None,
                   switches: Switches = DEFAULT_SWITCHES):
    if solver == "direct":
        return _DirectSolve.apply(d, geom)
    active = _t(geom.active, d.dtype); diag = _t(geom.diag, d.dtype)
    return _CGSolve.apply(d, active, diag, tol, max_it, switches.cg_batch_global_stop, stats)


def project(vy, vx, geom: KarmanGeom, solver="direct", tol=1e-5, max_it=2000, stats=None,
            switches: Switches = DEFAULT_SWITCHES):
    """divergence_free: hard BC mask, divergence, Poisson solve, gradient subtract."""
    my = _t(geom.face_my, vy.dtype); mx = _t(geom.face_mx, vy.dtype)
    vy = vy * my; vx = vx * mx
    d = divergence(vy, vx)
    p = pressure_solve(d, geom, solver, tol, max_it, stats, switches)
    pp = F.pad(p, (1, 1, 1, 1))
    gy = pp[:, 1:, 1:-1] - pp[:, :-1, 1:-1]     # [B, Y+1, X]: p[j]-p[j-1]
    gx = pp[:, 1:-1, 1:] - pp[:, 1:-1, :-1]     # [B, Y, X+1]
    return vy - my * gy, vx - mx * gx, p, d


def karman_step(rho, vy, vx, re, geom: KarmanGeom, dt=1.0, res=None, solver="direct", tol=1e-5,
                max_it=2000, stats=None, switches: Switches = DEFAULT_SWITCHES, return_aux=False):
    """One KarmanFlow.step (karman_train.py:173-185)."""
    res = geom.X if res is None else res
    dtp = vy.dtype
    alpha = dt * float(res) * float(res) / re.to(dtp)
    vy1, vx1 = diffuse_bc(vy, vx, alpha, _t(geom.bc_mask_y, dtp), _t(geom.bc_val_y, dtp))
    s = dt / geom.dx
    infl = _t(geom.inflow, dtp) * dt
    if rho is not None:
        if not switches.inflow_after_advect:
            rho = rho + infl
        rho2 = advect_density(rho, vy1, vx1, s, switches.density_extrap)
        if switches.inflow_after_advect:
            rho2 = rho2 + infl
    else:
        rho2 = None
    vy2, vx2 = advect_velocity(vy1, vx1, s)
    vy3, vx3, p, d = project(vy2, vx2, geom, solver, tol, max_it, stats, switches)
    if return_aux:
        return rho2, vy3, vx3, dict(vy1=vy1, vx1=vx1, vy2=vy2, vx2=vx2, p=p, d=d)
    return rho2, vy3, vx3


# --------------------------------------------------------------------------------------
# correction network (karman_train.py:92-138) — Keras weight layout [kh, kw, Cin, Cout]
# --------------------------------------------------------------------------------------
def model_layers(model: str = "mars_moon", cin0: int = 3) -> List[Tuple[int, int]]:
    if model == "mars_moon":
        return [(cin0, 32)] + [(32, 32)] * 10 + [(32, 2)]
    if model == "mercury":
        return [(cin0, 32), (32, 64), (64, 2)]
    raise ValueError(model)


def param_count(model="mars_moon", cin0=3) -> int:
    return sum(25 * ci * co + co for ci, co in model_layers(model, cin0))


def init_params(model="mars_moon", cin0=3, seed=0, dtype=torch.float64) -> List[torch.Tensor]:
    """Glorot-uniform kernels, zero biases (Keras Conv2D defaults), flat list [w0,b0,w1,b1,...]."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for ci, co in model_layers(model, cin0):
        lim = math.sqrt(6.0 / (25 * ci + 25 * co))
        w = (torch.rand(5, 5, ci, co, generator=g, dtype=torch.float64) * 2 - 1) * lim
        out += [w.to(dtype), torch.zeros(co, dtype=dtype)]
    return out


def flatten_params(params: Sequence[torch.Tensor]) -> torch.Tensor:
    return torch.cat([p.reshape(-1) for p in params])


def unflatten_params(flat: torch.Tensor, model="mars_moon", cin0=3) -> List[torch.Tensor]:
    out, o = [], 0
    for ci, co in model_layers(model, cin0):
        n = 25 * ci * co
        out.append(flat[o:o + n].reshape(5, 5, ci, co)); o += n
        out.append(flat[o:o + co]); o += co
    return out


def _conv(x_nhwc, w, b):
    """'same' zero-padded 5x5 cross-correlation, NHWC in/out, Keras kernel layout."""
    y = F.conv2d(x_nhwc.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), b, padding=2)
    return y.permute(0, 2, 3, 1)


def cnn_forward(params: Sequence[torch.Tensor], feat: torch.Tensor, model="mars_moon", return_acts=False):
    """feat [B,Y,X,Cin] -> [B,Y,X,2]."""
    lrelu = lambda t: F.leaky_relu(t, LEAKY_ALPHA)
    acts = []
    if model == "mars_moon":
        a = lrelu(_conv(feat, params[0], params[1])); acts.append(a)
        for k in range(1, 6):
            t = lrelu(_conv(a, params[2 * (2 * k - 1)], params[2 * (2 * k - 1) + 1])); acts.append(t)
            u = _conv(t, params[2 * (2 * k)], params[2 * (2 * k) + 1])
            a = lrelu(a + u); acts.append(a)
        out = _conv(a, params[22], params[23])
    elif model == "mercury":
        a = F.relu(_conv(feat, params[0], params[1])); acts.append(a)
        a = F.relu(_conv(a, params[2], params[3])); acts.append(a)
        out = _conv(a, params[4], params[5])
    else:
        raise ValueError(model)
    return (out, acts) if return_acts else out


def to_feature(vy, vx, re, sig):
    """karman_train.py:77-86 + normalisation :416-420.  sig = (s_vy, s_vx, s_re)."""
    B, Yp1, X = vy.shape
    Y = Yp1 - 1
    f0 = vy[:, :Y, :] / sig[0]
    f1 = vx[:, :, :X] / sig[1]
    f2 = (re.to(vy.dtype) / sig[2]).reshape(B, 1, 1).expand(B, Y, X)
    return torch.stack([f0, f1, f2], dim=-1)


def apply_correction(vy, vx, corr, sig):
    """to_staggered + add (karman_train.py:88-90, 421-426): row Y of vy and col X of vx get 0."""
    cy = F.pad(corr[..., 0] * sig[0], (0, 0, 0, 1))
    cx = F.pad(corr[..., 1] * sig[1], (0, 1, 0, 0))
    return vy + cy, vx + cx


def unrolled_loss(params, rho0, vy0, vx0, re, gt_vy, gt_vx, geom, sig, msteps, dt=1.0, model="mars_moon",
                  solver="direct", tol=1e-5, max_it=2000, stats=None, switches=DEFAULT_SWITCHES,
                  return_states=False):
    """karman_train.py:393-436: msteps x (step -> CNN correction -> add), l2 loss per step.

    gt_vy [m,B,Y+1,X], gt_vx [m,B,Y,X+1].  Returns (loss, per-step losses[, states]).
    """
    rho, vy, vx = rho0, vy0, vx0
    losses, states = [], []
    for i in range(msteps):
        rho, vy, vx = karman_step(rho, vy, vx, re, geom, dt=dt, solver=solver, tol=tol, max_it=max_it,
                                  stats=stats, switches=switches)
        feat = to_feature(vy, vx, re, sig)
        corr = cnn_forward(params, feat, model)
        vy, vx = apply_correction(vy, vx, corr, sig)
        li = 0.5 * (((gt_vy[i] - vy) / sig[0]) ** 2).sum() + 0.5 * (((gt_vx[i] - vx) / sig[1]) ** 2).sum()
        losses.append(li)
        if return_states:
            states.append((rho, vy, vx))
    loss = sum(losses) / msteps
    if return_states:
        return loss, losses, states
    return loss, losses


# --------------------------------------------------------------------------------------
# optimiser (karman_train.py:449-457): tf.compat.v1.train.AdamOptimizer
# --------------------------------------------------------------------------------------
def adam_tf1_step(theta, g, m, v, t: int, lr: float, b1=0.9, b2=0.999, eps=1e-8):
    """TF1 Adam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); theta -= lr_t*m/(sqrt(v)+eps) (eps outside
    the bias correction — differs from torch.optim.Adam).  Returns new (theta, m, v)."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    return theta - lr_t * m / (v.sqrt() + eps), m, v


# --------------------------------------------------------------------------------------
# synthetic inputs: reference warm start (karman.py:107-110) spun up by the oracle
# --------------------------------------------------------------------------------------
REYNOLDS_TRAIN = [10000.0 * 2 ** (i + 4) for i in range(6)]   # karman-2d/Makefile:22


def warm_start(geom: KarmanGeom, B: int, dtype=torch.float64):
    Y, X = geom.Y, geom.X
    vy = torch.ones(B, Y + 1, X, dtype=dtype)
    vx = torch.zeros(B, Y, X + 1, dtype=dtype)
    # karman.py:109 pokes the packed tensor [.., Y/2+10:Y/2+20, X/2-2:X/2+2, 1] (packed is [Y+1, X+1])
    P = (Y + 1) // 2; Q = (X + 1) // 2
    vx[:, P + 10:P + 20, Q - 2:Q + 2] = 1.0
    rho = torch.zeros(B, Y, X, dtype=dtype)
    return rho, vy, vx


def make_case(Y=128, X=64, B=3, msteps=4, spin=40, seed=0, dtype=torch.float64, noise=0.01, re_list=None):
    """Deterministic synthetic training case: (geom, rho0, vy0, vx0, re, gt_vy, gt_vx, sig)."""
    geom = KarmanGeom(Y, X)
    re_list = REYNOLDS_TRAIN if re_list is None else re_list
    re = torch.tensor([re_list[b % len(re_list)] for b in range(B)], dtype=torch.float64)
    rho, vy, vx = warm_start(geom, B, torch.float64)
    with torch.no_grad():
        for _ in range(spin):
            rho, vy, vx = karman_step(rho, vy, vx, re, geom)
        g = torch.Generator().manual_seed(seed)
        gts_y, gts_x = [], []
        r2, y2, x2 = rho, vy, vx
        for _ in range(msteps):
            r2, y2, x2 = karman_step(r2, y2, x2, re, geom)
            gts_y.append(y2 + noise * torch.randn(y2.shape, generator=g, dtype=torch.float64))
            gts_x.append(x2 + noise * torch.randn(x2.shape, generator=g, dtype=torch.float64))
    sig = (float(vy.abs().std()), float(vx.abs().std()) + 1e-3, float(np.std(np.abs(re_list))))
    cast = lambda t: t.to(dtype)
    return geom, cast(rho), cast(vy), cast(vx), cast(re), cast(torch.stack(gts_y)), cast(torch.stack(gts_x)), sig


# --------------------------------------------------------------------------------------
# Burgers (burgers_train.py:178-187) — periodic, staggered, no pressure solve
# --------------------------------------------------------------------------------------
def burgers_diffuse_fft(c: torch.Tensor, amount: float) -> torch.Tensor:
    """[PHI-RECALL, low confidence] periodic diffuse(): spectral decay exp(-(2 pi |k|)^2 * amount)
    on the component array (cycles per cell, array-size periodic)."""
    H, W = c.shape[-2:]
    ky = torch.fft.fftfreq(H, dtype=torch.float64).reshape(H, 1)
    kx = torch.fft.fftfreq(W, dtype=torch.float64).reshape(1, W)
    ker = torch.exp(-(2 * math.pi) ** 2 * (ky ** 2 + kx ** 2) * amount).to(c.dtype)
    return torch.fft.ifft2(torch.fft.fft2(c) * ker).real.to(c.dtype)


def lap_periodic(c):
    return (torch.roll(c, 1, -2) + torch.roll(c, -1, -2) + torch.roll(c, 1, -1) + torch.roll(c, -1, -1) - 4 * c)


def burgers_step(vy, vx, dt, dx, viscosity=0.1, fy=None, fx=None, switches: Switches = DEFAULT_SWITCHES):
    """Burgers.step: advect -> diffuse(nu*dt) -> (+ dt*f) on periodic staggered comps.
    Component arrays are [B, R+1, R] / [B, R, R+1] and wrap modulo the array size
    (PhiFlow-1 quirk)."""
    s = dt / dx
    vy2, vx2 = advect_velocity(vy, vx, s, mode="periodic")
    amount = viscosity * dt
    if switches.burgers_diffusion == "fft":
        vy3 = burgers_diffuse_fft(vy2, amount); vx3 = burgers_diffuse_fft(vx2, amount)
    else:
        vy3 = vy2 + amount * lap_periodic(vy2); vx3 = vx2 + amount * lap_periodic(vx2)
    if fy is not None:
        vy3 = vy3 + dt * fy; vx3 = vx3 + dt * fx
    return vy3, vx3


def burgers_to_feature(vy, vx, fy, fx, sig_v, sig_f):
    """burgers_train.py:75-90 + normalisation :398-415: [vy, vx (, fy, fx)][:, :-1, :-1] / std, channel-last.
    fy is None: to_feature_noforce."""
    Y, X = vx.shape[1], vy.shape[2]
    ch = [vy[:, :Y, :] / sig_v[0], vx[:, :, :X] / sig_v[1]]
    if fy is not None:
        ch += [fy[:, :Y, :] / sig_f[0], fx[:, :, :X] / sig_f[1]]
    return torch.stack(ch, dim=-1)


def burgers_unrolled_loss(params, vy0, vx0, f_vy, f_vx, gt_vy, gt_vx, dx, dt, sig_v, sig_f, msteps, viscosity=0.1,
                          model="mars_moon", switches: Switches = DEFAULT_SWITCHES, return_states=False):
    """burgers_train.py:379-437: msteps x (step_with_f -> CNN on [v, f_i]/std -> to_staggered*std_v -> add), per-step
    tf.nn.l2_loss of (gt - v)/std_v, total = sum / msteps.  f_vy/f_vx [m,B,..] or None (--noforce: plain step and
    velocity-only features)."""
    vy, vx = vy0, vx0
    losses, states = [], []
    for i in range(msteps):
        fy = None if f_vy is None else f_vy[i]
        fx = None if f_vx is None else f_vx[i]
        vy, vx = burgers_step(vy, vx, dt, dx, viscosity, fy, fx, switches)
        corr = cnn_forward(params, burgers_to_feature(vy, vx, fy, fx, sig_v, sig_f), model)
        vy, vx = apply_correction(vy, vx, corr, sig_v)
        losses.append(0.5 * (((gt_vy[i] - vy) / sig_v[0]) ** 2).sum() + 0.5 * (((gt_vx[i] - vx) / sig_v[1]) ** 2).sum())
        if return_states:
            states.append((vy, vx))
    loss = sum(losses) / msteps
    return (loss, losses, states) if return_states else (loss, losses)


def make_burgers_case(R=32, B=1, msteps=1, L=32.0, dt=0.1, seed=0, dtype=torch.float64, noise=0.01, force=True):
    """Deterministic synthetic Burgers case at the reference's SOL settings (burgers/Makefile:75-77: -l 32, --dt 0.1,
    32x32 after 4x down-sampling): smooth random periodic velocity and forcing (sums of a few sine modes, the shape
    burgers.py:89-121 generates), ground truth = uncorrected trajectory + noise.
    Returns (dx, vy0, vx0, f_vy, f_vx, gt_vy, gt_vx, sig_v, sig_f)."""
    g = torch.Generator().manual_seed(seed)
    dx = L / R

    def smooth(shape, amp):
        H, W = shape[-2:]
        yy = torch.arange(H, dtype=torch.float64).reshape(H, 1) / H
        xx = torch.arange(W, dtype=torch.float64).reshape(1, W) / W
        out = torch.zeros(shape, dtype=torch.float64)
        for _ in range(4):
            ky, kx = [int(k) for k in torch.randint(-2, 3, (2,), generator=g)]
            ph = torch.rand(shape[:-2] + (1, 1), generator=g, dtype=torch.float64) * 2 * math.pi
            a = (torch.rand(shape[:-2] + (1, 1), generator=g, dtype=torch.float64) * 2 - 1) * amp
            out = out + a * torch.sin(2 * math.pi * (ky * yy + kx * xx) + ph)
        return out

    vy = smooth((B, R + 1, R), 1.0)
    vx = smooth((B, R, R + 1), 1.0)
    f_vy = smooth((msteps, B, R + 1, R), 0.5) if force else None
    f_vx = smooth((msteps, B, R, R + 1), 0.5) if force else None
    gy, gx = [], []
    y2, x2 = vy, vx
    with torch.no_grad():
        for i in range(msteps):
            y2, x2 = burgers_step(y2, x2, dt, dx, 0.1, None if f_vy is None else f_vy[i], None if f_vx is None else f_vx[i])
            gy.append(y2 + noise * torch.randn(y2.shape, generator=g, dtype=torch.float64))
            gx.append(x2 + noise * torch.randn(x2.shape, generator=g, dtype=torch.float64))
    sig_v = (float(vy.std()) + 1e-3, float(vx.std()) + 1e-3)
    sig_f = (float(f_vy.std()) + 1e-3, float(f_vx.std()) + 1e-3) if force else (1.0, 1.0)
    c = lambda t: None if t is None else t.to(dtype)
    return dx, c(vy), c(vx), c(f_vy), c(f_vx), c(torch.stack(gy)), c(torch.stack(gx)), sig_v, sig_f
