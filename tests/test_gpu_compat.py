"""The reference's own apply loop (karman-2d/karman_apply.py:138-151) written against
solver_in_the_loop_b200.phi_compat — same statements, same argument names — vs the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import sol_oracle as so

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_reference_apply_loop(cuda_device):
    from solver_in_the_loop_b200.phi_compat import (OPEN, CorrectionModel, Domain, Fluid, KarmanFlow, StaggeredGrid, box, to_feature,
                                                    to_staggered, unstack_staggered_tensor)
    res, L, Re = 32, 100, 3.2e5
    st = Fluid(Domain(resolution=[res * 2, res], box=box[0:L * 2, 0:L], boundaries=OPEN), buoyancy_factor=0)
    # warm start exactly as karman_apply.py:96-101
    vn = st.velocity.staggered_tensor()
    vn[..., 0] = 1.0
    vn[..., vn.shape[1] // 2 + 10:vn.shape[1] // 2 + 20, vn.shape[2] // 2 - 2:vn.shape[2] // 2 + 2, 1] = 1.0
    v0 = StaggeredGrid(unstack_staggered_tensor(vn), st.velocity.box)
    st = st.copied_with(velocity=v0)
    vnp = np.zeros(tuple(st.velocity.data[0].data.shape))
    vnp[..., 0:2, 0:vnp.shape[2] - 1, 0] = 1.0
    vnp[..., 0:vnp.shape[1], 0:1, 0] = 1.0
    vnp[..., 0:vnp.shape[1], -1:, 0] = 1.0
    velBCy = vnp; velBCyMask = np.copy(vnp)
    data_stats = {"std": (1.0, (0.2, 0.08)), "ext.std": [1.7e6]}
    model = CorrectionModel(seed=0)
    model.set_weights([0.2 * w for w in model.get_weights()])
    simulator = KarmanFlow()
    simulator.cg.update(tol_abs=1e-7, tol_rel=1e-6, max_it=4000)

    geom = so.KarmanGeom(2 * res, res, L)
    rho, vy, vx = so.warm_start(geom, 1)
    params = [torch.as_tensor(w, dtype=torch.float64) for w in model.get_weights()]
    sig = (0.2, 0.08, 1.7e6)
    re_t = torch.tensor([Re], dtype=torch.float64)
    for i in range(1, 4):
        st = simulator.step(st, re=Re, res=res, velBCy=velBCy, velBCyMask=velBCyMask)
        inputf = to_feature(st, Re) / torch.tensor([*data_stats["std"][1], data_stats["ext.std"][0]], device=cuda_device)
        cv_pred = model.predict(inputf) * torch.tensor(data_stats["std"][1], device=cuda_device)
        cv = to_staggered(cv_pred, st.velocity.box)
        st = st.copied_with(velocity=st.velocity + cv)
        # oracle
        rho, vy, vx = so.karman_step(rho, vy, vx, re_t, geom)
        corr = so.cnn_forward(params, so.to_feature(vy, vx, re_t, sig))
        vy, vx = so.apply_correction(vy, vx, corr, sig)
        e = (rel(st.velocity.data[0].data[..., 0], vy), rel(st.velocity.data[1].data[..., 0], vx), rel(st.density.data[..., 0], rho))
        print("apply step", i, e)
        assert e[0] < 1e-5 and e[1] < 1e-5 and e[2] < 1e-5
    assert st.velocity.staggered_tensor().shape == (1, 2 * res + 1, res + 1, 2)


def test_pressure_solver_plugin_slot(cuda_device):
    """karman_train.py:51,167-168: `KarmanFlow(pressure_solver=...)`.  None / DirectProjection() = the exact direct projection (no
    iterations), SparseCG(accuracy) = the CG kernels with the reference's stop rule, CUDASolver() (--cuda) = the preconditioned CG;
    anything else is refused.  All three give the same step within their stop rule."""
    from solver_in_the_loop_b200.engine import SolError
    from solver_in_the_loop_b200.phi_compat import (OPEN, CUDASolver, DirectProjection, Domain, Fluid, KarmanFlow, SparseCG, StaggeredGrid, box,
                                                    unstack_staggered_tensor)
    res, L, Re = 32, 100, 3.2e5
    st0 = Fluid(Domain(resolution=[res * 2, res], box=box[0:L * 2, 0:L], boundaries=OPEN), buoyancy_factor=0)
    vn = st0.velocity.staggered_tensor()
    vn[..., 0] = 1.0
    vn[..., vn.shape[1] // 2 + 10:vn.shape[1] // 2 + 20, vn.shape[2] // 2 - 2:vn.shape[2] // 2 + 2, 1] = 1.0
    st0 = st0.copied_with(velocity=StaggeredGrid(unstack_staggered_tensor(vn), st0.velocity.box))
    vnp = np.zeros(tuple(st0.velocity.data[0].data.shape))
    vnp[..., 0:2, 0:vnp.shape[2] - 1, 0] = 1.0
    vnp[..., 0:vnp.shape[1], 0:1, 0] = 1.0
    vnp[..., 0:vnp.shape[1], -1:, 0] = 1.0
    outs = {}
    for name, solver in (("default", None), ("direct", DirectProjection()), ("cg", SparseCG(accuracy=1e-5)), ("cuda", CUDASolver())):
        sim = KarmanFlow(pressure_solver=solver)
        st = st0
        for _ in range(3):
            st = sim.step(st, re=Re, res=res, velBCy=vnp, velBCyMask=np.copy(vnp))
        outs[name] = (st.velocity.data[0].data.clone(), st.velocity.data[1].data.clone(), int(sim.last_iterations.max()))
    assert outs["default"][2] == 0 and outs["direct"][2] == 0
    assert outs["cg"][2] > outs["cuda"][2] > 0                  # plain CG needs more iterations than the preconditioned one
    for name in ("direct", "cg", "cuda"):
        tol = 0.0 if name == "direct" else 5e-3                 # the reference's stop rule truncates the pressure at ~7e-3
        assert rel(outs[name][0], outs["default"][0]) <= tol and rel(outs[name][1], outs["default"][1]) <= tol
    with pytest.raises(SolError):
        KarmanFlow(pressure_solver=object())


def test_phi2_flavoured_surface(cuda_device):
    """karman-2d-phi2/karman_train.py:149-196 signature: step(density_in, velocity_in, re, res, ...) -> [density, velocity];
    physical-units viscosity and inflow-before-advection, checked against the oracle with the matching switches."""
    from solver_in_the_loop_b200.phi_compat import OPEN, CenteredGrid, Domain, KarmanFlowPhi2, StaggeredGrid, box
    res, L, B, dt = 32, 100, 2, 1.0
    dm = Domain(resolution=[2 * res, res], box=box[0:2 * L, 0:L], boundaries=OPEN)
    geom = so.KarmanGeom(2 * res, res, L)
    rho, vy, vx = so.warm_start(geom, B)
    re = torch.tensor([1.6e5, 6.4e5], dtype=torch.float64)
    for _ in range(5):
        rho, vy, vx = so.karman_step(rho, vy, vx, re, geom)
    sim = KarmanFlowPhi2(dm)
    sim._plan(B, cuda_device).set_cg(tol_abs=1e-7, tol_rel=1e-6, max_it=4000)
    f = lambda t: t.to(cuda_device, torch.float32)
    den, vel = sim.step(CenteredGrid(f(rho), dm.box), StaggeredGrid([f(vy), f(vx)], dm.box), re=re, res=res, dt=dt)
    # oracle: alpha = dt^2 res^2/(re dx^2) == karman_step with res_eff = res*sqrt(dt)/dx; inflow added before the advection
    dx = L / res
    r_rho, r_vy, r_vx = so.karman_step(rho, vy, vx, re, geom, dt=dt, res=res * dt ** 0.5 / dx, switches=so.Switches(inflow_after_advect=False))
    assert rel(vel._vy, r_vy) < 1e-5 and rel(vel._vx, r_vx) < 1e-5 and rel(den._t, r_rho) < 1e-5
    assert vel.staggered_tensor().shape == (B, 2 * res + 1, res + 1, 2) and sim.solve_info["pressure"].data.shape == (B, 2 * res, res, 1)


def test_model_mercury_predict(cuda_device):
    from solver_in_the_loop_b200.phi_compat import model_mercury
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 24, 16, 3, generator=g)
    m = model_mercury(x, seed=1)
    ws = [torch.as_tensor(w).double() for w in m.get_weights()]
    ws[1] = 0.1 * torch.randn(32, generator=g, dtype=torch.float64); ws[3] = 0.1 * torch.randn(64, generator=g, dtype=torch.float64)
    m.set_weights([w.numpy() for w in ws])
    ref = so.cnn_forward(ws, x.double(), model="mercury")
    assert rel(m.predict(x.to(cuda_device)), ref) < 5e-6
