"""Error behaviour and edge cases of the C ABI on the GPU: every misuse returns a status code + message (SolError in Python), never a
crash or a silent fallback; the smallest legal shapes (one simulation, one unrolled step) work."""
import pytest
import torch

from oracle import sol_oracle as so

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_unroll_misuse_is_reported(cuda_device):
    from solver_in_the_loop_b200 import _lib, engine
    plan = engine.Plan.karman(64, 32, 2)
    sig = (0.4, 0.1, 1.7e6)
    with pytest.raises(engine.SolError):
        engine.Unroll(plan, 2, 3, sig)                      # more simulations than the plan holds
    with pytest.raises(engine.SolError):
        engine.Unroll(plan, 0, 2, sig)                      # msteps < 1
    with pytest.raises(engine.SolError):
        engine.Unroll(plan, 2, 2, (0.4, 0.0, 1.0))          # non-positive sigma
    with pytest.raises(engine.SolError):
        engine.Unroll(plan, 2, 2, sig, cin0=4)              # karman features have 3 channels
    un = engine.Unroll(plan, 2, 2, sig)
    w = torch.zeros(un.nparams, device=cuda_device)
    with pytest.raises(engine.SolError):
        un.backward(w)                                      # no forward yet
    vy, vx = plan.faces(2)
    re = torch.full((2,), 1e5, device=cuda_device)
    un.forward(w, re, vy, vx)                               # forward without ground truth: fine ...
    with pytest.raises(engine.SolError):
        un.backward(w)                                      # ... but then there is no loss to differentiate
    with pytest.raises(engine.SolError):
        un.forward(w, None, vy, vx)                         # the karman scene needs Re
    with pytest.raises(engine.SolError):
        un.forward(w.cpu(), re, vy, vx)                     # host tensor: no CPU fallback
    with pytest.raises(engine.SolError):
        un.forward(w.double(), re, vy, vx)                  # wrong dtype
    per = engine.Plan.periodic(32, 32, 1, 1.0)
    with pytest.raises(engine.SolError):
        engine.Unroll(per, 1, 1, (1.0, 1.0, 1.0), cin0=3)   # burgers features have 4 (or 2) channels
    ub = engine.Unroll(per, 1, 1, (1.0, 1.0, 1.0), dt=0.1, cin0=4)
    by, bx = per.faces(1)
    with pytest.raises(engine.SolError):
        ub.forward(torch.zeros(ub.nparams, device=cuda_device), None, by, bx)     # periodic plan without sol_unroll_set_burgers
    with pytest.raises(engine.SolError):
        ub.set_burgers(0.1, None, None, None, None)         # 4 feature channels need the forces
    with pytest.raises(engine.SolError):
        per.project(by, bx)                                 # no pressure solve on a periodic plan
    with pytest.raises(engine.SolError):
        engine.set_option("no_such_option", 1)


def test_unsupported_settings_are_reported_not_emulated(cuda_device):
    from solver_in_the_loop_b200 import engine
    plan = engine.Plan.karman(96, 48, 1)                    # X = 48: only the generic CG kernel covers this width ...
    vy, vx = plan.faces(1)
    plan.project(vy, vx)
    plan.set_cg(tol_abs=1e-5, tol_rel=0.0, max_it=2000, cluster=2)      # ... and it has no cluster variant: refused, not emulated
    with pytest.raises(engine.SolError) as e:
        plan.project(vy, vx)
    assert "error 3" in str(e.value) or "X in" in str(e.value)
    x = torch.zeros(1, 8, 8, 32, device=cuda_device)
    with pytest.raises(engine.SolError):
        engine.conv5x5(x, torch.zeros(5, 5, 32, 32, device=cuda_device), act=2)    # DLRELU without its reference tensor


def test_smallest_shapes(cuda_device):
    """One simulation, one unrolled step (BASELINE config 1 has batch 1, msteps 1) on the karman scene too."""
    from solver_in_the_loop_b200 import engine
    Y, X, B, m = 64, 32, 1, 1
    geom, rho, vy, vx, re, gty, gtx, sig = so.make_case(Y=Y, X=X, B=B, msteps=m, spin=12)
    params = so.init_params(seed=0)
    pr = [p.clone().requires_grad_() for p in params]
    loss, losses = so.unrolled_loss(pr, rho, vy, vx, re, gty, gtx, geom, sig, m)
    loss.backward()
    gref = so.flatten_params([p.grad for p in pr])
    plan = engine.Plan.karman(Y, X, B)
    plan.set_cg(tol_abs=1e-7, tol_rel=1e-6, max_it=4000, cluster=0)
    d = lambda t: t.to(cuda_device, torch.float32).contiguous()
    for graph in (False, True):
        un = engine.Unroll(plan, m, B, sig, use_graph=graph)
        g = torch.zeros(un.nparams, device=cuda_device)
        for _ in range(3):
            ls = un.train_iter(d(so.flatten_params(params)), d(re), d(vy), d(vx), d(gty), d(gtx), g)
        torch.cuda.synchronize()
        assert abs(float(ls[0]) - float(losses[0])) < 1e-4 * abs(float(losses[0]))
        assert rel(g, gref) < 1e-4
