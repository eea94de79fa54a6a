"""Per-cell stage arithmetic of the CUDA kernels (sol_cells.cuh compiled for the host) vs the CPU
oracle — forward values and adjoints (against torch autograd of the oracle).  No GPU needed."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import sol_oracle as so


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(t):
    return np.ascontiguousarray(t.detach().numpy().astype(np.float32))


@pytest.fixture(scope="module")
def case():
    torch.manual_seed(0)
    geom, rho, vy, vx, re, gty, gtx, sig = so.make_case(Y=32, X=32, B=2, msteps=1, spin=12)
    rho = rho + 0.3 * torch.rand_like(rho)
    return geom, rho, vy, vx, re


def test_diffuse_bc_forward_and_adjoint(emu, case):
    geom, rho, vy, vx, re = case
    B, Y, X = vy.shape[0], geom.Y, geom.X
    bcm = geom.bc_mask_y.astype(np.float32); bcv = geom.bc_val_y.astype(np.float32)
    re32 = _f32(re); vy32, vx32 = _f32(vy), _f32(vx)
    oy, ox = np.empty_like(vy32), np.empty_like(vx32)
    # use a large alpha (low Re) so that the Laplace term is well above fp32 noise
    re_small = np.array([200.0, 400.0], dtype=np.float32)
    emu.emu_diffuse_bc(B, Y, X, _p(re_small), C.c_float(1.0 * X * X), _p(vy32), _p(vx32), _p(bcm), _p(bcv), _p(oy), _p(ox))
    vyt = vy.clone().requires_grad_(); vxt = vx.clone().requires_grad_()
    alpha = 1.0 * X * X / torch.tensor(re_small, dtype=torch.float64)
    ry, rx = so.diffuse_bc(vyt, vxt, alpha, torch.tensor(geom.bc_mask_y), torch.tensor(geom.bc_val_y))
    assert np.abs(oy - ry.detach().numpy()).max() < 2e-5
    assert np.abs(ox - rx.detach().numpy()).max() < 2e-5
    gy = torch.randn_like(ry); gx = torch.randn_like(rx)
    (ry * gy).sum().backward(retain_graph=True); (rx * gx).sum().backward()
    gy32, gx32 = _f32(gy), _f32(gx)
    iy, ix = np.empty_like(vy32), np.empty_like(vx32)
    emu.emu_diffuse_bc_bwd(B, Y, X, _p(re_small), C.c_float(1.0 * X * X), _p(gy32), _p(gx32), _p(bcm), _p(iy), _p(ix))
    assert np.abs(iy - vyt.grad.numpy()).max() < 5e-5
    assert np.abs(ix - vxt.grad.numpy()).max() < 5e-5


@pytest.mark.parametrize("periodic", [0, 1])
def test_advect_forward_and_adjoint(emu, case, periodic):
    geom, rho, vy, vx, re = case
    B, Y, X = vy.shape[0], geom.Y, geom.X
    s = 1.0 / geom.dx * 2.5          # exaggerate the CFL number so that samples cross cell boundaries
    vy = vy * 1.7; vx = vx + 0.4 * torch.randn_like(vx)
    vy32, vx32, rho32 = _f32(vy), _f32(vx), _f32(rho)
    oy, ox, orho = np.empty_like(vy32), np.empty_like(vx32), np.empty_like(rho32)
    infl = geom.inflow.astype(np.float32)
    emu.emu_advect(B, Y, X, C.c_float(s), C.c_float(1.0), periodic, _p(vy32), _p(vx32), _p(rho32) if not periodic else None,
                   _p(infl) if not periodic else None, _p(oy), _p(ox), _p(orho) if not periodic else None)
    vyt = vy.clone().requires_grad_(); vxt = vx.clone().requires_grad_()
    ry, rx = so.advect_velocity(vyt, vxt, s, mode="periodic" if periodic else "replicate")
    assert np.abs(oy - ry.detach().numpy()).max() < 2e-5
    assert np.abs(ox - rx.detach().numpy()).max() < 2e-5
    if not periodic:
        rr = so.advect_density(rho, vy, vx, s, "zero") + torch.tensor(geom.inflow)
        assert np.abs(orho - rr.numpy()).max() < 2e-5
    gy = torch.randn_like(ry); gx = torch.randn_like(rx)
    ((ry * gy).sum() + (rx * gx).sum()).backward()
    iy, ix = np.empty_like(vy32), np.empty_like(vx32)
    emu.emu_advect_bwd(B, Y, X, C.c_float(s), periodic, _p(vy32), _p(vx32), _p(_f32(gy)), _p(_f32(gx)), _p(iy), _p(ix))
    # fp32 weights near cell boundaries: compare in relative L2
    ey = np.linalg.norm(iy - vyt.grad.numpy()) / np.linalg.norm(vyt.grad.numpy())
    ex = np.linalg.norm(ix - vxt.grad.numpy()) / np.linalg.norm(vxt.grad.numpy())
    assert ey < 1e-5 and ex < 1e-5, (ey, ex)


def test_projection_pieces(emu, case):
    geom, rho, vy, vx, re = case
    B, Y, X = vy.shape[0], geom.Y, geom.X
    my = geom.face_my.astype(np.float32); mx = geom.face_mx.astype(np.float32)
    vy32, vx32 = _f32(vy), _f32(vx)
    d = np.empty((B, Y, X), dtype=np.float32)
    emu.emu_divergence(B, Y, X, _p(vy32), _p(vx32), _p(my), _p(mx), _p(d))
    ref = so.divergence(vy * torch.tensor(geom.face_my), vx * torch.tensor(geom.face_mx)).numpy()
    assert np.abs(d - ref).max() < 1e-5
    # A p of the emulated stencil == oracle matrix
    p = np.random.default_rng(0).standard_normal((B, Y, X)).astype(np.float32)
    out = np.empty_like(p)
    act = geom.active.astype(np.uint8); diag = geom.diag.astype(np.float32)
    emu.emu_laplace(B, Y, X, _p(p), _p(act), _p(diag), _p(out))
    A = geom.laplace_matrix()
    refA = np.stack([A @ p[b].reshape(-1).astype(np.float64) for b in range(B)]).reshape(B, Y, X)
    assert np.abs(out - refA).max() < 1e-4
    assert abs(A - A.T).max() == 0          # symmetric
    # gradient subtract with the exact pressure reproduces the oracle projection
    vy3, vx3, pr, dd = so.project(vy, vx, geom)
    oy, ox = np.empty_like(vy32), np.empty_like(vx32)
    emu.emu_gradsub(B, Y, X, _p(vy32), _p(vx32), _p(_f32(pr)), _p(my), _p(mx), _p(oy), _p(ox))
    assert np.abs(oy - vy3.numpy()).max() < 2e-5
    assert np.abs(ox - vx3.numpy()).max() < 2e-5


@pytest.mark.parametrize("Y,X", [(64, 32), (128, 64), (256, 128)], ids=["64x32", "128x64", "256x128"])
def test_direct_projection_host_logic(emu, Y, X):
    """The direct pressure solver's host precomputation (sol_direct_host.h: sine-transform matrices, changed rows of the operator,
    capacitance-corrected basis) + the fp32 arithmetic of its two kernels, against the oracle's float64 sparse LU."""
    geom = so.KarmanGeom(Y, X)
    g = torch.Generator().manual_seed(11)
    vy = torch.randn(2, Y + 1, X, generator=g, dtype=torch.float64); vx = torch.randn(2, Y, X + 1, generator=g, dtype=torch.float64)
    _, _, pref, d = so.project(vy, vx, geom)
    act = np.ascontiguousarray(geom.active.astype(np.uint8)); diag = _f32(torch.tensor(geom.diag))
    d32 = _f32(d); out = np.zeros_like(d32)
    emu.emu_direct_solve.restype = C.c_int
    k = emu.emu_direct_solve(Y, X, 2, act.ctypes.data_as(C.c_void_p), _p(diag), _p(d32), _p(out))
    a = geom.active > 0
    surf = np.zeros_like(a)
    surf[1:, :] |= a[1:, :] != a[:-1, :]; surf[:-1, :] |= a[1:, :] != a[:-1, :]
    surf[:, 1:] |= a[:, 1:] != a[:, :-1]; surf[:, :-1] |= a[:, 1:] != a[:, :-1]
    assert k == int(surf.sum())                             # only the cells on either side of the obstacle surface
    fluid = geom.active > 0
    err = np.linalg.norm((out - pref.numpy())[:, fluid]) / np.linalg.norm(pref.numpy()[:, fluid])
    print("direct projection (host emulation) vs sparse LU:", err, "changed rows", k)
    assert err < (5e-6 if Y <= 128 else 1e-5)               # fp32 transforms: the round-off grows with the grid
