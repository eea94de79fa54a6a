"""CPU model of the direct pressure solver: fast Poisson solve on the obstacle-free rectangle (DST-I in both directions, the
"open" boundary is p = 0 one cell outside) + capacitance-matrix (Woodbury) correction for the rows the obstacle changes.
Checks the algebra against the sparse LU of the oracle's Laplace matrix, in float64 and with float32 arithmetic."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import sol_oracle as so  # noqa: E402


def dst_matrix(n, dtype=np.float64):
    k = np.arange(1, n + 1)
    return (np.sqrt(2.0 / (n + 1)) * np.sin(np.pi * np.outer(k, k) / (n + 1))).astype(dtype)


def build(geom, dtype=np.float64):
    Y, X = geom.Y, geom.X
    Sy, Sx = dst_matrix(Y), dst_matrix(X)
    lam = (-4.0 + 2.0 * np.cos(np.pi * np.arange(1, Y + 1) / (Y + 1))[:, None] + 2.0 * np.cos(np.pi * np.arange(1, X + 1) / (X + 1))[None, :])
    A = geom.laplace_matrix().toarray() if Y * X <= 8192 else None
    # rows the obstacle changes: solid cells and their fluid neighbours
    solid = geom.solid
    pad = np.pad(solid, 1)
    near = pad[:-2, 1:-1] | pad[2:, 1:-1] | pad[1:-1, :-2] | pad[1:-1, 2:]
    rows = np.flatnonzero((solid | near).ravel())
    k = len(rows)
    N = Y * X
    # A0: -4 on the diagonal, +1 to the in-domain neighbours
    import scipy.sparse as sp
    idx = np.arange(N).reshape(Y, X)
    r, c = [], []
    for a, b in ((idx[:-1, :], idx[1:, :]), (idx[:, :-1], idx[:, 1:])):
        r += [a.ravel(), b.ravel()]; c += [b.ravel(), a.ravel()]
    A0 = sp.coo_matrix((np.ones(sum(len(x) for x in r)), (np.concatenate(r), np.concatenate(c))), shape=(N, N)).tocsr() - 4.0 * sp.identity(N)
    Asp = geom.laplace_matrix()
    Rt = (Asp - A0).tocsr()[rows, :]                 # k x N sparse: the changed rows
    solve0 = lambda D: Sy @ ((Sy @ D @ Sx) / lam) @ Sx          # DST-I is symmetric and orthogonal
    W = np.zeros((N, k))
    for q, cidx in enumerate(rows):
        e = np.zeros((Y, X)); e.ravel()[cidx] = 1.0
        W[:, q] = solve0(e).ravel()
    M = np.linalg.inv(np.eye(k) + Rt @ W)
    return dict(Sy=Sy.astype(dtype), Sx=Sx.astype(dtype), ilam=(1.0 / lam).astype(dtype), rows=rows, Rt=Rt.astype(dtype), W=W.astype(dtype), M=M.astype(dtype), k=k)


def solve(S, d):
    dt = S["Sy"].dtype
    D = d.astype(dt)
    p0 = S["Sy"] @ ((S["Sy"] @ D @ S["Sx"]) * S["ilam"]) @ S["Sx"]
    s = S["Rt"] @ p0.ravel()
    t = S["M"] @ s
    return (p0.ravel() - S["W"] @ t).reshape(d.shape)


if __name__ == "__main__":
    Y, X = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (128, 64)
    geom, rho, vy, vx, re, gty, gtx, sig = so.make_case(Y=Y, X=X, B=2, msteps=1, spin=20)
    _, _, rp, rd = so.project(vy * 1.01, vx, geom)
    d = rd[0].numpy(); pref = rp[0].numpy()
    for dt in (np.float64, np.float32):
        S = build(geom, dt)
        p = solve(S, d)
        act = geom.active > 0
        err = np.linalg.norm((p - pref)[act]) / np.linalg.norm(pref[act])
        print("%s: changed rows k = %d, rel L2 error of p on fluid cells %.2e, max |p| on solid cells %.2e (ref %.2e)" %
              (dt.__name__, S["k"], err, np.abs(p[~act]).max(), np.abs(pref[~act]).max()))
