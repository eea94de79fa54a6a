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
        assert e[0] < 1e-5 and e[1] < 1e-4 and e[2] < 1e-4
    assert st.velocity.staggered_tensor().shape == (1, 2 * res + 1, res + 1, 2)
