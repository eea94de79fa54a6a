"""Known-answer / self-consistency tests that pin the CPU oracle.  The reference
ships no tests or golden vectors and PhiFlow/TF cannot be installed here ("parity unpinned"), so
these pin every convention the oracle restates."""
import math

import numpy as np
import pytest
import torch

from oracle import sol_oracle as so


@pytest.fixture(scope="module")
def geom():
    return so.KarmanGeom(64, 32)


def test_geometry_constants(geom):
    # karman_train.py:170-171: Inflow(box[5:10, 25:75]), Obstacle(Sphere([50, 50], 10)) in physical units
    dx = geom.dx
    assert dx == 100.0 / 32
    j, i = np.nonzero(geom.solid)
    cy, cx = (j + 0.5) * dx, (i + 0.5) * dx
    assert np.all((cy - 50) ** 2 + (cx - 50) ** 2 <= 100.0)
    assert geom.solid.sum() == 32                       # disc of radius 10 on a 3.125-cell grid
    jj, ii = np.nonzero(geom.inflow)
    assert ((jj + 0.5) * dx >= 5).all() and ((jj + 0.5) * dx <= 10).all()
    assert ((ii + 0.5) * dx >= 25).all() and ((ii + 0.5) * dx <= 75).all()
    # BC mask exactly as the numpy code at karman_train.py:366-372
    vn = np.zeros((65, 32)); vn[0:2, 0:31] = 1; vn[:, 0:1] = 1; vn[:, -1:] = 1
    assert np.array_equal(geom.bc_mask_y, vn)


def test_diffuse_constant_is_fixed_point_and_heat_kernel():
    c = torch.full((1, 9, 7), 3.0, dtype=torch.float64)
    assert torch.equal(so.lap_replicate(c), torch.zeros_like(c))
    d = torch.zeros(1, 9, 7, dtype=torch.float64); d[0, 4, 3] = 1.0
    out = d + 0.1 * so.lap_replicate(d)
    assert out[0, 4, 3] == pytest.approx(0.6) and out[0, 3, 3] == pytest.approx(0.1) and out.sum() == pytest.approx(1.0)
    # replicate padding conserves the sum (Neumann) even at the edge
    e = torch.zeros(1, 9, 7, dtype=torch.float64); e[0, 0, 0] = 1.0
    assert (e + 0.2 * so.lap_replicate(e)).sum() == pytest.approx(1.0)


def test_advection_linear_field_exact_and_shift():
    # constant velocity (uy, ux): a linear field is advected exactly; shift = u*dt/dx cells
    Y, X = 16, 16
    s = 0.32
    vy = torch.full((1, Y + 1, X), 1.0, dtype=torch.float64)
    vx = torch.zeros(1, Y, X + 1, dtype=torch.float64)
    j, i = so._grid(1, Y, X, torch.float64, "cpu")
    rho = 2.0 * j + 0.5 * i
    out = so.advect_density(rho, vy, vx, s, "replicate")
    assert torch.allclose(out[:, 1:, :], (rho - 2.0 * s)[:, 1:, :], atol=1e-12)
    # uniform flow is a fixed point of the velocity self-advection
    ay, ax = so.advect_velocity(vy, vx, s)
    assert torch.allclose(ay, vy) and torch.allclose(ax, vx)


def test_density_zero_outside():
    rho = torch.ones(1, 8, 8, dtype=torch.float64)
    vy = torch.full((1, 9, 8), 1.0, dtype=torch.float64); vx = torch.zeros(1, 8, 9, dtype=torch.float64)
    out = so.advect_density(rho, vy, vx, 0.25, "zero")
    assert torch.allclose(out[0, 0], torch.full((8,), 0.75, dtype=torch.float64))    # blends with the zero outside
    assert torch.allclose(out[0, 1:], torch.ones(7, 8, dtype=torch.float64))


def test_staggered_pack_roundtrip_and_feature_layout():
    B, Y, X = 2, 6, 5
    g = torch.Generator().manual_seed(0)
    vy = torch.randn(B, Y + 1, X, generator=g, dtype=torch.float64); vx = torch.randn(B, Y, X + 1, generator=g, dtype=torch.float64)
    packed = torch.zeros(B, Y + 1, X + 1, 2, dtype=torch.float64)
    packed[:, :, :-1, 0] = vy; packed[:, :-1, :, 1] = vx                      # staggered_tensor(): "v first, u second"
    assert torch.equal(packed[:, :, :-1, 0], vy) and torch.equal(packed[:, :-1, :, 1], vx)
    re = torch.tensor([1.0e5, 2.0e5], dtype=torch.float64)
    f = so.to_feature(vy, vx, re, (1.0, 1.0, 1.0))
    assert torch.equal(f[..., 0:2], packed[:, :-1, :-1, 0:2]) and torch.equal(f[0, :, :, 2], torch.full((Y, X), 1.0e5, dtype=torch.float64))
    corr = torch.randn(B, Y, X, 2, generator=g, dtype=torch.float64)
    ny, nx = so.apply_correction(vy, vx, corr, (2.0, 3.0))
    assert torch.equal(ny[:, Y], vy[:, Y]) and torch.equal(nx[:, :, X], vx[:, :, X])
    assert torch.allclose(ny[:, :Y] - vy[:, :Y], 2.0 * corr[..., 0]) and torch.allclose(nx[:, :, :X] - vx[:, :, :X], 3.0 * corr[..., 1])


def test_projection_properties(geom):
    g = torch.Generator().manual_seed(1)
    vy = torch.randn(2, 65, 32, generator=g, dtype=torch.float64); vx = torch.randn(2, 64, 33, generator=g, dtype=torch.float64)
    py, px, p, d = so.project(vy, vx, geom)
    act = torch.tensor(geom.active)
    assert float((so.divergence(py, px) * act).abs().max()) < 1e-12
    assert float(py[:, torch.tensor(geom.face_my) == 0].abs().max()) == 0.0
    qy, qx, _, _ = so.project(py, px, geom)
    assert torch.allclose(qy, py, atol=1e-12) and torch.allclose(qx, px, atol=1e-12)        # idempotent
    A = geom.laplace_matrix()
    assert abs(A - A.T).max() == 0                                                            # symmetric
    # reference-style CG converges to the direct solve; float32 run reports its iteration count
    x, its = so.cg_reference(d, act, torch.tensor(geom.diag), tol=1e-10, max_it=5000)
    assert torch.allclose(x, p, atol=1e-7) and int(its.max()) < 1000
    # self-adjoint projection: <P a, b> = <a, P b>
    ay = torch.randn(2, 65, 32, generator=g, dtype=torch.float64); ax = torch.randn(2, 64, 33, generator=g, dtype=torch.float64)
    Pa = so.project(ay, ax, geom)
    lhs = (Pa[0] * vy).sum() + (Pa[1] * vx).sum(); rhs = (ay * py).sum() + (ax * px).sum()
    assert float(abs(lhs - rhs)) < 1e-9 * float(abs(lhs))


def test_uniform_flow_is_fixed_point_without_obstacle():
    geom = so.KarmanGeom(32, 16, obstacle=None)
    vy = torch.ones(1, 33, 16, dtype=torch.float64); vx = torch.zeros(1, 32, 17, dtype=torch.float64)
    rho = torch.zeros(1, 32, 16, dtype=torch.float64)
    re = torch.tensor([1.0e5], dtype=torch.float64)
    _, ny, nx = so.karman_step(rho, vy, vx, re, geom)
    assert torch.allclose(ny, vy, atol=1e-10) and torch.allclose(nx, vx, atol=1e-10)


def test_step_adjoint_dot_product_and_finite_difference():
    geom, rho, vy, vx, re, gty, gtx, sig = so.make_case(Y=32, X=32, B=1, msteps=2, spin=8)
    params = [p.requires_grad_() for p in so.init_params(seed=0)]
    loss, _ = so.unrolled_loss(params, rho, vy, vx, re, gty, gtx, geom, sig, 2)
    loss.backward()
    # float64 central finite difference of the loss w.r.t. a few CNN weights through m=2 steps
    rng = np.random.default_rng(0)
    for (k, idx) in [(0, (2, 2, 1, 5)), (6, (0, 4, 3, 7)), (22, (2, 2, 9, 1)), (23, (0,))]:
        eps = 1e-7
        base = params[k].detach().clone()
        vals = []
        for sgn in (+1, -1):
            pp = [q.detach().clone() for q in params]
            pp[k][idx] = base[idx] + sgn * eps
            l, _ = so.unrolled_loss(pp, rho, vy, vx, re, gty, gtx, geom, sig, 2)
            vals.append(float(l))
        fd = (vals[0] - vals[1]) / (2 * eps)
        an = float(params[k].grad[idx])
        assert abs(fd - an) < 1e-4 * max(1.0, abs(an)), (k, idx, fd, an)


def test_cnn_matches_keras_conventions():
    # zero 'same' padding, cross-correlation, Keras layout [kh,kw,Cin,Cout], LeakyReLU(0.3)
    x = torch.zeros(1, 7, 7, 1, dtype=torch.float64); x[0, 3, 3, 0] = 1.0
    w = torch.arange(25, dtype=torch.float64).reshape(5, 5, 1, 1)
    y = so._conv(x, w, None)
    # correlation: y[j,i] = sum_{a,b} x[j+a-2, i+b-2] w[a,b]  ->  impulse response is the flipped kernel
    assert torch.equal(y[0, 1:6, 1:6, 0], torch.flip(w[:, :, 0, 0], dims=(0, 1)))
    assert so.param_count() == 260354 and so.param_count(cin0=4) == 261154 and so.param_count("mercury") == 2432 + 51264 + 3202
    assert torch.nn.functional.leaky_relu(torch.tensor(-1.0), so.LEAKY_ALPHA).item() == pytest.approx(-0.3)


def test_adam_tf1_differs_from_torch_adam_in_epsilon_placement():
    th = torch.tensor([1.0], dtype=torch.float64); g = torch.tensor([1e-6], dtype=torch.float64)
    t1, m, v = so.adam_tf1_step(th, g, torch.zeros(1, dtype=torch.float64), torch.zeros(1, dtype=torch.float64), 1, 0.1)
    lr_t = 0.1 * math.sqrt(1 - 0.999) / (1 - 0.9)
    expect = 1.0 - lr_t * (0.1 * 1e-6) / (math.sqrt(0.001 * 1e-12) + 1e-8)
    assert float(t1) == pytest.approx(expect, rel=1e-12)


def test_burgers_step_properties():
    R = 16
    g = torch.Generator().manual_seed(3)
    vy = 0.3 * torch.randn(1, R + 1, R, generator=g, dtype=torch.float64); vx = 0.3 * torch.randn(1, R, R + 1, generator=g, dtype=torch.float64)
    # constant field is a fixed point; diffusion conserves the mean (k=0 mode untouched)
    cy = torch.full_like(vy, 0.5); cx = torch.full_like(vx, -0.25)
    oy, ox = so.burgers_step(cy, cx, dt=0.1, dx=1.0)
    assert torch.allclose(oy, cy, atol=1e-12) and torch.allclose(ox, cx, atol=1e-12)
    d = so.burgers_diffuse_fft(vy, 0.01)
    assert float(abs(d.mean() - vy.mean())) < 1e-14 and float(d.std()) < float(vy.std())
    fy = torch.ones_like(vy); fx = torch.zeros_like(vx)
    a = so.burgers_step(vy, vx, 0.1, 1.0); b = so.burgers_step(vy, vx, 0.1, 1.0, fy=fy, fx=fx)
    assert torch.allclose(b[0] - a[0], 0.1 * fy) and torch.allclose(b[1], a[1])


def test_projection_against_an_independent_solver():
    """The oracle's sparse-LU pressure solve vs an independent route to the same answer: sine-transform fast Poisson solve of the
    obstacle-free rectangle + capacitance-matrix correction of the rows the obstacle changes (float64, scripts/probes/direct_model.py)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("direct_model", os.path.join(root, "scripts", "probes", "direct_model.py"))
    dm = importlib.util.module_from_spec(spec); spec.loader.exec_module(dm)
    geom = so.KarmanGeom(64, 32)
    g = torch.Generator().manual_seed(21)
    vy = torch.randn(1, 65, 32, generator=g, dtype=torch.float64); vx = torch.randn(1, 64, 33, generator=g, dtype=torch.float64)
    _, _, pref, d = so.project(vy, vx, geom)
    S = dm.build(geom, np.float64)
    p = dm.solve(S, d[0].numpy())
    fluid = geom.active > 0
    err = np.linalg.norm((p - pref[0].numpy())[fluid]) / np.linalg.norm(pref[0].numpy()[fluid])
    assert err < 1e-10, err
    assert S["k"] > int((geom.active == 0).sum())
