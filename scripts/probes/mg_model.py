"""CPU model (numpy, fp64) of the multigrid-preconditioned CG of sol_cg_mg.cu — same hierarchy (cell-centred 2x
coarsening, piecewise-constant P, R = P^T, rediscretised coarse operators on coarsened masks, damped Jacobi, exact
coarsest solve) — to try smoother / cycle variants offline and count PCG iterations at the reference's stop rule."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import sol_oracle as so  # noqa: E402


def build(act0):
    levels = [act0.astype(bool)]
    while levels[-1].shape[1] > 4:
        a = levels[-1]
        s = a[0::2, 0::2].astype(int) + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2]
        levels.append(s >= 2)
    H = []
    for a in levels:
        Yl, Xl = a.shape
        pad = np.ones((Yl + 2, Xl + 2))
        pad[1:-1, 1:-1] = a
        d = pad[:-2, 1:-1] + pad[2:, 1:-1] + pad[1:-1, :-2] + pad[1:-1, 2:]
        d = np.maximum(d, 1.0)
        H.append(dict(act=a, diag=d))
    # dense coarsest
    a = levels[-1]; Yc, Xc = a.shape; N = Yc * Xc
    A = np.zeros((N, N))
    for j in range(Yc):
        for i in range(Xc):
            c = j * Xc + i
            if not a[j, i]:
                A[c, c] = 1.0; continue
            A[c, c] = -H[-1]["diag"][j, i]
            for jj, ii in ((j - 1, i), (j + 1, i), (j, i - 1), (j, i + 1)):
                if 0 <= jj < Yc and 0 <= ii < Xc and a[jj, ii]:
                    A[c, jj * Xc + ii] = 1.0
    Inv = np.linalg.inv(A)
    m = a.reshape(-1)
    Inv = Inv * m[:, None] * m[None, :]
    H[-1]["inv"] = Inv
    return H


def applyA(L, u):
    a = L["act"]
    p = np.zeros((u.shape[0] + 2, u.shape[1] + 2))
    p[1:-1, 1:-1] = u * a
    nb = p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:]
    return (nb - L["diag"] * u) * a


def smooth(L, u, b, om):
    return u + (-om / L["diag"]) * (b - applyA(L, u)) * L["act"]


def vcycle(H, l, b, pre, post, scale=1.0):
    L = H[l]
    if l == len(H) - 1:
        return (L["inv"] @ b.reshape(-1)).reshape(b.shape)
    u = np.zeros_like(b)
    for om in pre:
        u = smooth(L, u, b, om)
    r = (b - applyA(L, u)) * L["act"]
    rc = (r[0::2, 0::2] + r[0::2, 1::2] + r[1::2, 0::2] + r[1::2, 1::2]) * H[l + 1]["act"]
    ec = vcycle(H, l + 1, rc, pre, post, scale)
    u = u + scale * np.kron(ec, np.ones((2, 2))) * L["act"]
    for om in post:
        u = smooth(L, u, b, om)
    return u


def pcg(H, d, pre, post, tol=1e-5, scale=1.0, maxit=200):
    L = H[0]
    x = np.zeros_like(d); r = d * L["act"]; p = np.zeros_like(d); rz = 0.0
    it = 0
    while np.abs(r).max() >= tol and it < maxit:
        z = vcycle(H, 0, r, pre, post, scale)
        rzn = (r * z).sum()
        beta = 0.0 if it == 0 else rzn / rz
        rz = rzn
        p = z + beta * p
        q = applyA(L, p)
        alpha = rz / (p * q).sum()
        x += alpha * p; r -= alpha * q
        it += 1
    return x, it


if __name__ == "__main__":
    Y, X = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 64)
    geom, rho, vy, vx, re, gty, gtx, sig = so.make_case(Y=Y, X=X, B=3, msteps=1, spin=30)
    s = 1.0 / geom.dx
    vy1, vx1 = so.diffuse_bc(vy, vx, (1.0 * X * X / re).reshape(-1, 1, 1), torch.tensor(geom.bc_mask_y), torch.tensor(geom.bc_val_y)) if hasattr(geom, "bc_val_y") else (vy, vx)
    ay, ax = so.advect_velocity(vy1, vx1, s)
    _, _, _, rd = so.project(ay, ax, geom)
    g = torch.Generator().manual_seed(3)
    rhs = [rd[b].numpy() for b in range(3)]
    # an adjoint-like right-hand side: divergence of a random face field
    ry = torch.randn(vy.shape, generator=g, dtype=torch.float64) * 0.1; rx = torch.randn(vx.shape, generator=g, dtype=torch.float64) * 0.1
    _, _, _, rd2 = so.project(ry, rx, geom)
    rhs += [rd2[0].numpy()]
    H = build(np.asarray(geom.active))
    c1, c2 = 1.0 / (1.5 + 0.5 * np.cos(np.pi / 4)), 1.0 / (1.5 - 0.5 * np.cos(np.pi / 4))
    variants = {
        "V(2,2) w=0.8 (current)": ((0.8, 0.8), (0.8, 0.8), 1.0),
        "V(2,2) cheb(%.3f,%.3f)" % (c1, c2): ((c1, c2), (c2, c1), 1.0),
        "V(2,2) cheb reversed": ((c2, c1), (c1, c2), 1.0),
        "V(1,1) w=0.8": ((0.8,), (0.8,), 1.0),
        "V(2,2) w=0.8 scale 1.3": ((0.8, 0.8), (0.8, 0.8), 1.3),
        "V(2,2) cheb scale 1.3": ((c1, c2), (c2, c1), 1.3),
        "V(2,2) w=(0.6,1.0)": ((0.6, 1.0), (1.0, 0.6), 1.0),
        "V(2,2) w=(0.7,0.95)": ((0.7, 0.95), (0.95, 0.7), 1.0),
        "V(3,3) w=0.8": ((0.8,) * 3, (0.8,) * 3, 1.0),
        "V(2,1) w=0.8": ((0.8, 0.8), (0.8,), 1.0),
    }
    for name, (pre, post, sc) in variants.items():
        its = [pcg(H, d, pre, post, scale=sc)[1] for d in rhs]
        print("%-32s iterations %s" % (name, its), flush=True)


# ---- variant: cell-centred bilinear prolongation (R = P^T), masked by the fluid cells ----
def prolong_bilinear(ec, actc, actf):
    Yc, Xc = ec.shape
    e = np.where(actc, ec, 0.0)
    p = np.pad(e, 1, mode="edge")            # replicate outside (open boundary p=0 would be 'constant'; try both)
    out = np.zeros((2 * Yc, 2 * Xc))
    for dj in (0, 1):
        for di in (0, 1):
            sj = -1 if dj == 0 else 1
            si = -1 if di == 0 else 1
            c = p[1:-1, 1:-1]
            ny = p[1 + sj:Yc + 1 + sj, 1:-1]
            nx = p[1:-1, 1 + si:Xc + 1 + si]
            nd = p[1 + sj:Yc + 1 + sj, 1 + si:Xc + 1 + si]
            out[dj::2, di::2] = (9 * c + 3 * ny + 3 * nx + nd) / 16.0
    return out * actf


def restrict_bilinear_T(r, actc, actf):
    # adjoint of prolong_bilinear with zero (not replicate) outside handling approximated: use explicit transpose via autograd-free loops
    Yf, Xf = r.shape; Yc, Xc = Yf // 2, Xf // 2
    acc = np.zeros((Yc + 2, Xc + 2))
    rr = r * actf
    for dj in (0, 1):
        for di in (0, 1):
            sj = -1 if dj == 0 else 1
            si = -1 if di == 0 else 1
            blk = rr[dj::2, di::2]
            acc[1:-1, 1:-1] += 9 / 16.0 * blk
            acc[1 + sj:Yc + 1 + sj, 1:-1] += 3 / 16.0 * blk
            acc[1:-1, 1 + si:Xc + 1 + si] += 3 / 16.0 * blk
            acc[1 + sj:Yc + 1 + sj, 1 + si:Xc + 1 + si] += 1 / 16.0 * blk
    # fold the replicate padding back (transpose of edge padding)
    acc[1, :] += acc[0, :]; acc[-2, :] += acc[-1, :]
    acc[:, 1] += acc[:, 0]; acc[:, -2] += acc[:, -1]
    return acc[1:-1, 1:-1] * actc


def vcycle_b(H, l, b, pre, post, scale=1.0, mixed=False):
    L = H[l]
    if l == len(H) - 1:
        return (L["inv"] @ b.reshape(-1)).reshape(b.shape)
    u = np.zeros_like(b)
    for om in pre:
        u = smooth(L, u, b, om)
    r = (b - applyA(L, u)) * L["act"]
    if mixed:
        rc = (r[0::2, 0::2] + r[0::2, 1::2] + r[1::2, 0::2] + r[1::2, 1::2]) * H[l + 1]["act"]
    else:
        rc = restrict_bilinear_T(r, H[l + 1]["act"], L["act"])
    ec = vcycle_b(H, l + 1, rc, pre, post, scale, mixed)
    u = u + scale * prolong_bilinear(ec, H[l + 1]["act"], L["act"])
    for om in post:
        u = smooth(L, u, b, om)
    return u


def pcg_b(H, d, pre, post, tol=1e-5, scale=1.0, maxit=100, mixed=False):
    L = H[0]
    x = np.zeros_like(d); r = d * L["act"]; p = np.zeros_like(d); rz = 0.0
    it = 0
    while np.abs(r).max() >= tol and it < maxit:
        z = vcycle_b(H, 0, r, pre, post, scale, mixed)
        rzn = (r * z).sum()
        beta = 0.0 if it == 0 else rzn / rz
        rz = rzn
        p = z + beta * p
        q = applyA(L, p)
        alpha = rz / (p * q).sum()
        x += alpha * p; r -= alpha * q
        it += 1
    return x, it


if __name__ == "__main__":
    for name, (pre, post, sc) in {"bilinear V(2,2) w=0.8": ((0.8, 0.8), (0.8, 0.8), 1.0), "bilinear V(1,1) w=0.8": ((0.8,), (0.8,), 1.0),
                                   "bilinear V(2,2) scale 0.5": ((0.8, 0.8), (0.8, 0.8), 0.5), "bilinear V(1,1) scale 0.5": ((0.8,), (0.8,), 0.5),
                                   "bilinear V(2,2) scale 0.7": ((0.8, 0.8), (0.8, 0.8), 0.7), "bilinear V(1,1) scale 0.7": ((0.8,), (0.8,), 0.7)}.items():
        its = [pcg_b(H, d, pre, post, scale=sc)[1] for d in rhs]
        print("%-32s iterations %s" % (name, its), flush=True)


# ---- variant: red-black Gauss-Seidel smoothing (symmetric: R,B before / B,R after) ----
def rb_halfsweep(L, u, b, color, om=1.0):
    Yl, Xl = u.shape
    jj, ii = np.meshgrid(np.arange(Yl), np.arange(Xl), indexing="ij")
    m = ((jj + ii) % 2 == color) & L["act"]
    a = L["act"]
    p = np.zeros((Yl + 2, Xl + 2)); p[1:-1, 1:-1] = u * a
    nb = p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:]
    unew = (nb - b) / L["diag"]
    out = u.copy()
    out[m] = (1 - om) * u[m] + om * unew[m]
    return out


def vcycle_rb(H, l, b, nu, om=1.0):
    L = H[l]
    if l == len(H) - 1:
        return (L["inv"] @ b.reshape(-1)).reshape(b.shape)
    u = np.zeros_like(b)
    for _ in range(nu):
        u = rb_halfsweep(L, u, b, 0, om); u = rb_halfsweep(L, u, b, 1, om)
    r = (b - applyA(L, u)) * L["act"]
    rc = (r[0::2, 0::2] + r[0::2, 1::2] + r[1::2, 0::2] + r[1::2, 1::2]) * H[l + 1]["act"]
    ec = vcycle_rb(H, l + 1, rc, nu, om)
    u = u + np.kron(ec, np.ones((2, 2))) * L["act"]
    for _ in range(nu):
        u = rb_halfsweep(L, u, b, 1, om); u = rb_halfsweep(L, u, b, 0, om)
    return u


def pcg_rb(H, d, nu, tol=1e-5, om=1.0, maxit=100):
    L = H[0]
    x = np.zeros_like(d); r = d * L["act"]; p = np.zeros_like(d); rz = 0.0
    it = 0
    while np.abs(r).max() >= tol and it < maxit:
        z = vcycle_rb(H, 0, r, nu, om)
        rzn = (r * z).sum()
        beta = 0.0 if it == 0 else rzn / rz
        rz = rzn
        p = z + beta * p
        q = applyA(L, p)
        alpha = rz / (p * q).sum()
        x += alpha * p; r -= alpha * q
        it += 1
    return x, it


if __name__ == "__main__":
    for nu in (1, 2):
        for om in (1.0, 1.15):
            its = [pcg_rb(H, d, nu, om=om)[1] for d in rhs]
            print("RB-GS V(%d,%d) omega=%.2f          iterations %s" % (nu, nu, om, its), flush=True)
