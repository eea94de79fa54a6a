"""--adplr and --clip-grad of the train scripts (karman_train.py:146-163, 449-457, 492)."""
import numpy as np
import pytest
import torch

from oracle import sol_oracle as so


def test_adplr_schedule_matches_reference_table():
    """lr_schedule(epoch, lr) applied per epoch as the epoch loop does (karman_train.py:492): x0.1 at epochs 11, 16, 21 and
    x0.5 at 23 (zero-based epoch index), constant elsewhere."""
    from solver_in_the_loop_b200.trainer import lr_schedule
    lr, seen = 1e-3, []
    for j in range(30):
        lr = lr_schedule(j, lr)
        seen.append(lr)
    expect, cur = [], 1e-3
    for j in range(30):
        cur *= {11: 1e-1, 16: 1e-1, 21: 1e-1, 23: 0.5}.get(j, 1.0)
        expect.append(cur)
    assert np.allclose(seen, expect, rtol=1e-12)
    assert seen[10] == 1e-3 and abs(seen[29] - 0.5e-6) < 1e-18


@pytest.mark.gpu
@pytest.mark.parametrize("clip", [False, True], ids=["plain", "clip-grad"])
def test_trainer_step_matches_oracle_adam(cuda_device, clip):
    """One SolTrainer optimiser step (graph path) == oracle gradient -> optional per-variable tf.clip_by_norm(1e-3) -> TF1 Adam."""
    from solver_in_the_loop_b200 import engine
    from solver_in_the_loop_b200.trainer import SolTrainer
    Y, X, B, m = 64, 32, 2, 2
    geom, rho, vy, vx, re, gty, gtx, sig = so.make_case(Y=Y, X=X, B=B, msteps=m, spin=10)
    params = [p * 0.3 for p in so.init_params(seed=0)]
    w0 = so.flatten_params(params)
    pr = [p.clone().requires_grad_() for p in params]
    loss, _ = so.unrolled_loss(pr, rho, vy, vx, re, gty, gtx, geom, sig, m)
    loss.backward()
    g = so.flatten_params([p.grad for p in pr])
    if clip:
        o = 0
        for ci, co in so.model_layers():
            for n in (25 * ci * co, co):
                nrm = float(g[o:o + n].norm())
                g[o:o + n] *= min(1.0, 1e-3 / nrm)
                o += n
        assert float(g.norm()) < float(so.flatten_params([p.grad for p in pr]).norm())      # the clip is active in this case
    theta, _, _ = so.adam_tf1_step(w0, g, torch.zeros_like(w0), torch.zeros_like(w0), 1, 1e-4)
    plan = engine.Plan.karman(Y, X, B)
    tr = SolTrainer(plan, m, B, sig, lr=1e-4, weights=w0.float(), clip_grad=clip, use_graph=False)
    d = lambda t: t.to(device=cuda_device, dtype=torch.float32).contiguous()
    l = tr.train_step(d(re), d(vy), d(vx), d(gty), d(gtx))
    torch.cuda.synchronize()
    rel = lambda a, b: float((a.double().cpu() - b).norm() / b.norm())
    # the update is lr * m/(sqrt(v)+eps) ~ lr * sign(g): compare the UPDATE, not the weights it barely moves
    upd, upd_ref = tr.weights.double().cpu() - w0, theta - w0
    print("loss rel", abs(float(l) - float(loss)) / float(loss), "grad rel", rel(tr.grad, g), "update rel", float((upd - upd_ref).norm() / upd_ref.norm()))
    assert abs(float(l) - float(loss)) < 1e-5 * float(loss)
    assert rel(tr.grad, g) < 2e-5
    assert float((upd - upd_ref).norm() / upd_ref.norm()) < 1e-3


def test_pressure_solver_plugin_classes_cpu():
    """karman_train.py:51,167-168: the pressure_solver slot of IncompressibleFlow / KarmanFlow (host logic only; the GPU behaviour is
    covered by tests/test_gpu_compat.py)."""
    import pytest
    from solver_in_the_loop_b200._lib import SolError
    from solver_in_the_loop_b200.phi_compat import CUDASolver, DirectProjection, KarmanFlow, SparseCG
    assert KarmanFlow().direct_solve == 1 and KarmanFlow(DirectProjection()).direct_solve == 1
    cg = KarmanFlow(pressure_solver=SparseCG(accuracy=3e-6, max_iterations=777))
    assert cg.direct_solve == 0 and cg.cg_precond == 0 and cg.cg == dict(tol_abs=3e-6, tol_rel=0.0, max_it=777, cluster=0)
    cu = KarmanFlow(pressure_solver=CUDASolver())
    assert cu.direct_solve == 0 and cu.cg_precond == 1 and cu.cg["tol_abs"] == 1e-5 and cu.cg["max_it"] == 2000
    with pytest.raises(SolError):
        KarmanFlow(pressure_solver="SparseCG")
    with pytest.raises(SolError):
        KarmanFlow(make_input_divfree=True)
