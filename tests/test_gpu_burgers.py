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
