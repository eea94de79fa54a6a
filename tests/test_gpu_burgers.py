"""BASELINE config 1 (burgers 2D 32x32, msteps=1, batch=1): forward + adjoint of the periodic
Burgers step on the GPU vs the CPU oracle, and the reference-shaped step_with_f surface."""
import pytest
import torch

from oracle import sol_oracle as so

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.mark.parametrize("fd", [False, True], ids=["fft-kernel", "explicit-5pt"])
def test_burgers_step_forward_and_adjoint(cuda_device, fd):
    from solver_in_the_loop_b200 import engine, phi_compat
    R, B, dt, nu, dx = 32, 2, 0.1, 0.1, 1.0
    g = torch.Generator().manual_seed(0)
    vy = 0.5 * torch.randn(B, R + 1, R, generator=g, dtype=torch.float64)
    vx = 0.5 * torch.randn(B, R, R + 1, generator=g, dtype=torch.float64)
    fy = torch.randn(B, R + 1, R, generator=g, dtype=torch.float64); fx = torch.randn(B, R, R + 1, generator=g, dtype=torch.float64)
    sw = so.Switches(burgers_diffusion="fd" if fd else "fft")
    vyt = vy.clone().requires_grad_(); vxt = vx.clone().requires_grad_()
    ry, rx = so.burgers_step(vyt, vxt, dt, dx, nu, fy, fx, switches=sw)
    plan = engine.Plan.periodic(R, R, B, dx)
    d = lambda t: t.to(cuda_device, torch.float32).contiguous()
    ky = kx = None
    if not fd:
        ky = phi_compat.periodic_diffusion_kernel(R + 1, R, nu * dt, cuda_device)
        kx = phi_compat.periodic_diffusion_kernel(R, R + 1, nu * dt, cuda_device)
    oy, ox = plan.burgers_step(d(vy), d(vx), dt, nu, ky, kx, d(fy), d(fx))
    print("burgers fwd rel", rel(oy, ry), rel(ox, rx))
    assert rel(oy, ry) < 5e-6 and rel(ox, rx) < 5e-6
    gy = torch.randn(ry.shape, generator=g, dtype=torch.float64); gx = torch.randn(rx.shape, generator=g, dtype=torch.float64)
    ((ry * gy).sum() + (rx * gx).sum()).backward()
    iy, ix = plan.burgers_step_bwd(d(vy), d(vx), d(gy), d(gx), dt, nu, ky, kx)
    print("burgers bwd rel", rel(iy, vyt.grad), rel(ix, vxt.grad))
    assert rel(iy, vyt.grad) < 2e-5 and rel(ix, vxt.grad) < 2e-5


def test_burgers_compat_surface(cuda_device):
    from solver_in_the_loop_b200.phi_compat import BurgersTest, BurgersVelocitySMAC, Domain, PERIODIC, StaggeredGrid, box
    R, B = 32, 1
    dm = Domain(resolution=[R, R], box=box([32, 32]), boundaries=PERIODIC)
    g = torch.Generator().manual_seed(1)
    vy = 0.3 * torch.randn(B, R + 1, R, generator=g); vx = 0.3 * torch.randn(B, R, R + 1, generator=g)
    st = BurgersVelocitySMAC(dm, batch_size=B).copied_with(velocity=StaggeredGrid([vy.to(cuda_device), vx.to(cuda_device)], dm.box))
    fr = BurgersVelocitySMAC(dm, batch_size=B)
    sim = BurgersTest()
    out = sim.step_with_f(st, fr, dt=0.1)
    ry, rx = so.burgers_step(vy.double(), vx.double(), 0.1, 1.0, 0.1)
    assert rel(out.velocity._vy, ry) < 5e-6 and rel(out.velocity._vx, rx) < 5e-6
    assert out.velocity.staggered_tensor().shape == (B, R + 1, R + 1, 2)


def _burgers_params(cin0):
    params = so.init_params(cin0=cin0, seed=0)
    params = [0.3 * p for p in params]      # a mild correction, as a trained net gives
    for k in range(1, len(params), 2):
        params[k] = 0.01 * torch.randn(params[k].shape, generator=torch.Generator().manual_seed(k), dtype=torch.float64)
    return params


@pytest.mark.parametrize("R,B,m,force", [(32, 1, 1, True), (32, 5, 4, True), (32, 2, 2, False)],
                         ids=["C1-32x32-b1-m1", "SOL04-32x32-b5-m4", "noforce-b2-m2"])
def test_burgers_unrolled_training_parity(cuda_device, R, B, m, force):
    """BASELINE config 1 (burgers 32x32, msteps=1, batch=1) and burgers/Makefile:75-77 (SOL-04: -m 4 -b 5): corrected
    states, per-step losses, weight gradients and input gradient of the unrolled Burgers iteration
    (burgers_train.py:379-437) vs the float64 oracle with autograd."""
    from solver_in_the_loop_b200 import engine, phi_compat
    dt, nu = 0.1, 0.1
    dx, vy, vx, fy, fx, gty, gtx, sv, sf = so.make_burgers_case(R=R, B=B, msteps=m, dt=dt, force=force)
    cin0 = 4 if force else 2
    params = _burgers_params(cin0)
    pr = [p.clone().requires_grad_() for p in params]
    vy0 = vy.clone().requires_grad_(); vx0 = vx.clone().requires_grad_()
    loss, losses, states = so.burgers_unrolled_loss(pr, vy0, vx0, fy, fx, gty, gtx, dx, dt, sv, sf, m, nu, return_states=True)
    loss.backward()
    gref = so.flatten_params([p.grad for p in pr])

    d = lambda t: None if t is None else t.to(cuda_device, torch.float32).contiguous()
    plan = engine.Plan.periodic(R, R, B, dx)
    un = engine.Unroll(plan, m, B, (sv[0], sv[1], 1.0), dt=dt, cin0=cin0)
    assert un.nparams == so.param_count(cin0=cin0)
    ky = phi_compat.periodic_diffusion_kernel(R + 1, R, nu * dt, cuda_device)
    kx = phi_compat.periodic_diffusion_kernel(R, R + 1, nu * dt, cuda_device)
    un.set_burgers(nu, ky, kx, d(fy), d(fx), sf)
    w = d(so.flatten_params(params))
    ls, pv, px, _ = un.forward(w, None, d(vy), d(vx), d(gty), d(gtx), return_pred=True)
    for i in range(m):
        print("step", i, "state rel", rel(pv[i], states[i][0]), rel(px[i], states[i][1]), "loss", float(ls[i]), float(losses[i]))
        assert rel(pv[i], states[i][0]) < 1e-5 and rel(px[i], states[i][1]) < 1e-5
        assert abs(float(ls[i]) - float(losses[i])) < 1e-5 * abs(float(losses[i]))
    gw, gy0, gx0 = un.backward(w, want_input_grad=True)
    print("grad rel", rel(gw, gref), "input grad rel", rel(gy0, vy0.grad), rel(gx0, vx0.grad))
    assert rel(gw, gref) < 2e-5
    # (input gradient: discontinuous coordinate derivative of the semi-Lagrangian sample, see test_gpu_quoted_configs.py)
    assert rel(gy0, vy0.grad) < 1e-3 and rel(gx0, vx0.grad) < 1e-3


def test_burgers_trainer_descends_and_graph_matches_eager(cuda_device):
    """BurgersTrainer (the sess.run of burgers_train.py:487): CUDA-graph replay == eager launches, and Adam steps on a
    fixed batch reduce the loss."""
    from solver_in_the_loop_b200 import engine
    from solver_in_the_loop_b200.trainer import BurgersTrainer
    R, B, m, dt = 32, 5, 4, 0.1
    dx, vy, vx, fy, fx, gty, gtx, sv, sf = so.make_burgers_case(R=R, B=B, msteps=m, dt=dt)
    d = lambda t: t.to(cuda_device, torch.float32).contiguous()
    w0 = so.flatten_params(_burgers_params(4)).float()
    out = {}
    for graph in (False, True):
        plan = engine.Plan.periodic(R, R, B, dx)
        tr = BurgersTrainer(plan, m, B, sv, sf, viscosity=0.1, dt=dt, lr=1e-4, weights=w0, use_graph=graph)
        args = (d(vy), d(vx), d(fy), d(fx), d(gty), d(gtx))
        out[graph] = [float(tr.train_step(*args)) for _ in range(6)]
    print("eager", out[False]); print("graph", out[True])
    assert out[False][-1] < out[False][0]
    for a, b in zip(out[False], out[True]):
        assert abs(a - b) < 1e-4 * abs(a)
