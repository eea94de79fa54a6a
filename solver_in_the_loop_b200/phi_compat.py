"""Look-alikes of the PhiFlow-1.5.1 objects the reference scripts touch on the hot path, backed by
the CUDA engine — so that the reference's own loop (karman-2d/karman_apply.py:138-151)

    st = simulator.step(st, re=..., res=..., velBCy=velBCy, velBCyMask=velBCyMask)
    inputf = to_feature(st, re) / [*std_v, std_re]
    cv = to_staggered(model.predict(inputf) * std_v, st.velocity.box)
    st = st.copied_with(velocity=st.velocity + cv)

runs unchanged with ``from solver_in_the_loop_b200.phi_compat import *`` instead of
``from phi.flow import *``.  Only the surface listed in SURVEY.md §8b is provided; tensors are
torch CUDA tensors (fp32).  Layout conventions [PINNED by the reference]: packed staggered tensor
[B, Y+1, X+1, 2] with channel 0 = y ("v first", karman_train.py:367), component grids
velocity.data[0].data [B, Y+1, X, 1] and velocity.data[1].data [B, Y, X+1, 1].
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib, engine
from .engine import Plan, SolError

OPEN = "open"
PERIODIC = "periodic"


# ---- geometry ------------------------------------------------------------------------------------
class AABox:
    def __init__(self, lower, upper):
        self.lower = np.asarray(lower, dtype=np.float64)
        self.upper = np.asarray(upper, dtype=np.float64)

    @property
    def size(self):
        return self.upper - self.lower

    def __eq__(self, o):
        return isinstance(o, AABox) and np.array_equal(self.lower, o.lower) and np.array_equal(self.upper, o.upper)

    def __repr__(self):
        return "AABox(%s, %s)" % (self.lower.tolist(), self.upper.tolist())


class _BoxFactory:
    """``box[0:200, 0:100]`` (karman_train.py:363) or ``box([len, len])`` (burgers_train.py:344)."""

    def __getitem__(self, item):
        item = item if isinstance(item, tuple) else (item,)
        return AABox([s.start or 0 for s in item], [s.stop for s in item])

    def __call__(self, size):
        return AABox([0] * len(size), list(size))


box = _BoxFactory()


class Sphere:
    def __init__(self, center, radius):
        self.center, self.radius = np.asarray(center, dtype=np.float64), float(radius)


class Obstacle:
    def __init__(self, geometry):
        self.geometry = geometry


class Inflow:
    def __init__(self, geometry, rate=1.0):
        self.geometry, self.rate = geometry, float(rate)


class Gravity:
    def __init__(self, gravity=-9.81):
        self.gravity = gravity


class Domain:
    def __init__(self, resolution, box=None, boundaries=OPEN):
        self.resolution = [int(r) for r in resolution]
        self.box = AABox([0, 0], self.resolution) if box is None else box
        self.boundaries = boundaries

    @property
    def dx(self):
        return self.box.size / np.asarray(self.resolution)


# ---- fields --------------------------------------------------------------------------------------
class CenteredGrid:
    """A scalar grid; ``.data`` is [B, H, W, 1] (a view of the contiguous [B, H, W] storage)."""

    def __init__(self, data, box=None, name=None):
        t = data
        if isinstance(t, np.ndarray):
            t = torch.as_tensor(t, dtype=torch.float32)
        if t.dim() == 4:
            t = t[..., 0]
        self._t = t.contiguous()
        self.box = box

    @property
    def data(self):
        return self._t.unsqueeze(-1)

    @property
    def resolution(self):
        return list(self._t.shape[1:3])


class StaggeredGrid:
    def __init__(self, data, box=None, name=None):
        if isinstance(data, StaggeredGrid):
            data, box = [data._vy, data._vx], (box or data.box)
        if isinstance(data, (list, tuple)):
            comps = [c._t if isinstance(c, CenteredGrid) else (c[..., 0] if c.dim() == 4 else c) for c in data]
            self._vy, self._vx = comps[0].contiguous(), comps[1].contiguous()
        else:       # packed [B, Y+1, X+1, 2]
            t = torch.as_tensor(data) if isinstance(data, np.ndarray) else data
            self._vy = t[:, :, :-1, 0].contiguous()
            self._vx = t[:, :-1, :, 1].contiguous()
        self.box = box

    @property
    def data(self) -> List[CenteredGrid]:
        return [CenteredGrid(self._vy), CenteredGrid(self._vx)]

    def staggered_tensor(self) -> torch.Tensor:
        B, Yp1, X = self._vy.shape
        t = torch.zeros(B, Yp1, X + 1, 2, dtype=self._vy.dtype, device=self._vy.device)
        t[:, :, :-1, 0] = self._vy
        t[:, :-1, :, 1] = self._vx
        return t

    def _bin(self, o, op):
        if isinstance(o, StaggeredGrid):
            return StaggeredGrid([op(self._vy, o._vy), op(self._vx, o._vx)], self.box)
        return StaggeredGrid([op(self._vy, o), op(self._vx, o)], self.box)

    def __add__(self, o):
        return self._bin(o, torch.add)

    def __sub__(self, o):
        return self._bin(o, torch.sub)

    def __mul__(self, o):
        return self._bin(o, torch.mul)

    __rmul__ = __mul__


def unstack_staggered_tensor(t):
    return [t[:, :, :-1, 0:1], t[:, :-1, :, 1:2]]


class Fluid:
    """Fluid(Domain(...), buoyancy_factor=0, batch_size=B) (karman_train.py:363)."""

    def __init__(self, domain: Domain, density=None, velocity=None, buoyancy_factor=0.0, batch_size=None, device=None):
        self.domain = domain
        self.buoyancy_factor = buoyancy_factor
        B = 1 if batch_size is None else int(batch_size)
        Y, X = domain.resolution
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.density = density if density is not None else CenteredGrid(torch.zeros(B, Y, X, device=dev), domain.box)
        self.velocity = velocity if velocity is not None else StaggeredGrid(
            [torch.zeros(B, Y + 1, X, device=dev), torch.zeros(B, Y, X + 1, device=dev)], domain.box)
        self.age = 0.0

    @property
    def _batch_size(self):
        return self.velocity._vy.shape[0]

    def copied_with(self, density=None, velocity=None, age=None):
        dev = self.velocity._vy.device

        def as_dev(t):
            if isinstance(t, np.ndarray):
                t = torch.as_tensor(t, dtype=torch.float32)
            return t.to(device=dev, dtype=torch.float32)

        f = Fluid.__new__(Fluid)
        f.domain, f.buoyancy_factor = self.domain, self.buoyancy_factor
        f.age = self.age if age is None else age
        if density is None:
            f.density = self.density
        else:
            f.density = density if isinstance(density, CenteredGrid) else CenteredGrid(as_dev(density), self.domain.box)
        if velocity is None:
            f.velocity = self.velocity
        else:
            f.velocity = velocity if isinstance(velocity, StaggeredGrid) else StaggeredGrid(as_dev(velocity), self.domain.box)
        return f

    def staggered_grid(self, name="field", value=0):
        v = self.velocity
        return StaggeredGrid([torch.full_like(v._vy, float(value)), torch.full_like(v._vx, float(value))], v.box)


# ---- pressure-solver plug-in slot (karman_train.py:51,167-168: pressure_solver=None | SparseCG(...) | CUDASolver()) ---------------
class SparseCG:
    """PhiFlow's ``SparseCG(accuracy=1e-5, max_iterations=2000)``: selects the on-chip CG kernels with the reference's recurrences and
    stop rule (max|r| < accuracy) instead of the default direct projection."""

    def __init__(self, accuracy=1e-5, gradient_accuracy="same", max_iterations=2000, max_gradient_iterations="same", autodiff=False):
        self.accuracy = float(accuracy)
        self.max_iterations = int(max_iterations)
        self.preconditioned = False


class CUDASolver(SparseCG):
    """``phi.tf.tf_cuda_pressuresolver.CUDASolver`` (the reference's --cuda flag): same stop rule; here the multigrid-preconditioned
    CG kernel."""

    def __init__(self, accuracy=1e-5, max_iterations=2000):
        SparseCG.__init__(self, accuracy=accuracy, max_iterations=max_iterations)
        self.preconditioned = True


class DirectProjection:
    """The engine's default: precomputed fast Poisson solve + capacitance correction (exact, no stop rule)."""


class IncompressibleFlow:
    def __init__(self, pressure_solver=None, make_input_divfree=False, make_output_divfree=True):
        if make_input_divfree or not make_output_divfree:
            raise SolError("only make_input_divfree=False, make_output_divfree=True (the reference's setting) is implemented")
        self._plans = {}
        self.cg = dict(tol_abs=1e-5, tol_rel=0.0, max_it=2000, cluster=0)     # SparseCG defaults
        self.direct_solve, self.cg_precond = 1, 1
        if isinstance(pressure_solver, SparseCG):
            self.cg = dict(tol_abs=pressure_solver.accuracy, tol_rel=0.0, max_it=pressure_solver.max_iterations, cluster=0)
            self.direct_solve, self.cg_precond = 0, int(pressure_solver.preconditioned)
        elif pressure_solver is not None and not isinstance(pressure_solver, DirectProjection):
            raise SolError("pressure_solver must be None, DirectProjection(), SparseCG(...) or CUDASolver(...) from phi_compat "
                           "(PhiFlow solver objects cannot run inside the sm_100a engine)")
        self.last_iterations = None

    def _configure(self, plan: Plan) -> Plan:
        plan.set_cg(**self.cg)
        plan.set_option("direct_solve", self.direct_solve)
        plan.set_option("cg_precond", self.cg_precond)
        return plan


class KarmanFlow(IncompressibleFlow):
    """karman-2d/karman_train.py:166-185.  Geometry in physical units of the state's box."""

    def __init__(self, pressure_solver=None, make_input_divfree=False, make_output_divfree=True):
        IncompressibleFlow.__init__(self, pressure_solver, make_input_divfree, make_output_divfree)
        self.infl = Inflow(box[5:10, 25:75])
        self.obst = Obstacle(Sphere([50, 50], 10))

    def _plan(self, smoke: Fluid, velBCy, velBCyMask) -> Plan:
        Y, X = smoke.domain.resolution
        B = smoke._batch_size
        dev = smoke.velocity._vy.device
        def digest(a_):      # content, not identity: callers often rebuild (np.copy) the BC arrays every step
            a_ = a_.detach().cpu().numpy() if isinstance(a_, torch.Tensor) else np.asarray(a_)
            return hash(np.ascontiguousarray(a_, dtype=np.float32).tobytes())

        key = (Y, X, B, str(dev), digest(velBCy), digest(velBCyMask))
        if key in self._plans:
            return self._plans[key]
        dxy = smoke.domain.dx
        if abs(dxy[0] - dxy[1]) > 1e-9 * dxy[0]:
            raise SolError("square cells required (box/resolution = %s)" % dxy)
        dx = float(dxy[1])
        lo = smoke.domain.box.lower
        cy = lo[0] + (np.arange(Y) + 0.5) * dx
        cx = lo[1] + (np.arange(X) + 0.5) * dx
        CY, CX = np.meshgrid(cy, cx, indexing="ij")
        c, r = self.obst.geometry.center, self.obst.geometry.radius
        solid = ((CY - c[0]) ** 2 + (CX - c[1]) ** 2) <= r ** 2
        b = self.infl.geometry
        inflow = (((CY >= b.lower[0]) & (CY <= b.upper[0]) & (CX >= b.lower[1]) & (CX <= b.upper[1])) * self.infl.rate).astype(np.float32)

        def bc(a):
            a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
            return np.ascontiguousarray(a.reshape(-1, Y + 1, X)[0], dtype=np.float32)

        plan = self._configure(Plan(Y, X, B, dx, _lib.SOL_BOUNDARY_OPEN, solid, inflow, bc(velBCyMask), bc(velBCy), device=dev))
        self._plans[key] = plan
        return plan

    def step(self, smoke: Fluid, re, res, velBCy, velBCyMask, dt=1.0, gravity=None) -> Fluid:
        plan = self._plan(smoke, velBCy, velBCyMask)
        B = smoke._batch_size
        dev = smoke.velocity._vy.device
        if isinstance(re, torch.Tensor):
            re_t = re.to(device=dev, dtype=torch.float32).reshape(-1)
        else:
            re_t = torch.as_tensor(np.asarray(re, dtype=np.float32).reshape(-1), device=dev)
        if re_t.numel() == 1 and B > 1:
            re_t = re_t.expand(B)
        re_t = re_t.contiguous()
        out = plan.step_fwd(re_t, smoke.velocity._vy.float().contiguous(), smoke.velocity._vx.float().contiguous(),
                            rho=smoke.density._t.float().contiguous(), dt=float(dt), res=float(res))
        self.last_iterations = out["iters"]
        new = smoke.copied_with(density=CenteredGrid(out["rho"], smoke.density.box),
                                velocity=StaggeredGrid([out["vy"], out["vx"]], smoke.velocity.box))
        new.age = smoke.age + dt
        return new


# ---- correction network surface --------------------------------------------------------------------
def to_feature(smokestate, ext_const_channel):
    """karman_train.py:77-86 / karman_apply.py:44-52: [vy, vx, Re] channel-last, [B, Y, X, 3]."""
    states = smokestate if isinstance(smokestate, (list, tuple)) else [smokestate]
    feats = []
    for s in states:
        feats.append(s.velocity.staggered_tensor()[:, :-1, :-1, 0:2])
    d = states[0].density._t
    if isinstance(ext_const_channel, torch.Tensor):
        ext = ext_const_channel.to(device=d.device, dtype=d.dtype).reshape(-1, 1, 1, 1)
    else:
        ext = torch.as_tensor(np.asarray(ext_const_channel, dtype=np.float32), device=d.device).reshape(-1, 1, 1, 1)
    feats.append(torch.ones_like(d).unsqueeze(-1) * ext)
    return torch.cat(feats, dim=-1)


def to_staggered(tensor_cen, box):
    """karman_train.py:88-90: zero-pad [B,Y,X,2] to [B,Y+1,X+1,2] and wrap as a StaggeredGrid."""
    t = torch.nn.functional.pad(tensor_cen, (0, 0, 0, 1, 0, 1))
    return StaggeredGrid(t, box=box)


class CorrectionModel:
    """model_mars_moon (karman_train.py:101-138) on the CUDA conv kernels; weights in Keras order."""

    LAYERS = [(3, 32)] + [(32, 32)] * 10 + [(32, 2)]

    def __init__(self, weights: Optional[Sequence] = None, cin0: int = 3, seed: int = 0, device=None, model: str = "mars_moon"):
        from .trainer import glorot_uniform_params, model_layers
        self.model = model
        self.layers = model_layers(model, cin0)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.flat = glorot_uniform_params(model, cin0=cin0, seed=seed).to(dev)
        if weights is not None:
            self.set_weights(weights)

    def _views(self):
        out, o = [], 0
        for ci, co in self.layers:
            n = 25 * ci * co
            out.append(self.flat[o:o + n].view(5, 5, ci, co)); o += n
            out.append(self.flat[o:o + co]); o += co
        return out

    def get_weights(self):
        return [v.detach().cpu().numpy() for v in self._views()]

    def set_weights(self, ws):
        for v, w in zip(self._views(), ws):
            v.copy_(torch.as_tensor(np.asarray(w), dtype=torch.float32).to(v.device))

    def __call__(self, x):
        x = x.to(dtype=torch.float32).contiguous()
        v = self._views()
        L = _lib.SOL_ACT_LRELU
        if self.model == "mercury":      # karman_train.py:92-99: relu = leaky with slope 0
            a = engine.conv5x5(x, v[0], v[1], act=L, slope=0.0)
            a = engine.conv5x5(a, v[2], v[3], act=L, slope=0.0)
            return engine.conv5x5(a, v[4], v[5])
        a = engine.conv5x5(x, v[0], v[1], act=L)
        for k in range(1, 6):
            t = engine.conv5x5(a, v[2 * (2 * k - 1)], v[2 * (2 * k - 1) + 1], act=L)
            a = engine.conv5x5(t, v[2 * (2 * k)], v[2 * (2 * k) + 1], addend=a, act=L)
        return engine.conv5x5(a, v[22], v[23])

    predict = __call__

    def save(self, path):
        np.savez(path, *self.get_weights())

    @classmethod
    def load(cls, path, **kw):
        if str(path).endswith(".h5"):
            # the reference Makefiles name Keras checkpoints (model.h5); this package's train scripts write model.npz beside them
            import os
            alt = str(path)[:-3] + ".npz"
            if not os.path.exists(alt):
                raise SolError("%s is a Keras HDF5 checkpoint: convert it with `python -m solver_in_the_loop_b200.scripts.keras_h5_to_npz "
                               "model.h5 model.npz` (needs h5py) or pass the model.npz written by this package's train scripts" % path)
            path = alt
        z = np.load(path)
        ws = [z["arr_%d" % i] for i in range(len(z.files))]
        kw.setdefault("cin0", int(ws[0].shape[2]))                        # 3 karman, 4 / 2 burgers
        kw.setdefault("model", "mercury" if len(ws) == 6 else "mars_moon")
        return cls(weights=ws, **kw)


def model_mars_moon(tensor_in=None, **kw):
    cin0 = 3 if tensor_in is None else int(tensor_in.shape[-1])
    return CorrectionModel(cin0=cin0, **kw)


def model_mercury(tensor_in=None, **kw):
    """karman_train.py:92-99: Conv2D(32, relu) -> Conv2D(64, relu) -> Conv2D(2)."""
    cin0 = 3 if tensor_in is None else int(tensor_in.shape[-1])
    return CorrectionModel(cin0=cin0, model="mercury", **kw)


# ---- Burgers (burgers/burgers_train.py:172-187) ------------------------------------------------------
def burgers_to_feature(smokestates, forcestates):
    """burgers_train.py:75-82 ``to_feature(smokestates, forcestates)``: [vy, vx, fy, fx][:, :-1, :-1] channel-last (the karman
    script's ``to_feature(state, Re)`` has the same name but another signature, hence the prefix here)."""
    return torch.cat([s.velocity.staggered_tensor()[:, :-1, :-1, 0:2] for s in smokestates] +
                     [f.velocity.staggered_tensor()[:, :-1, :-1, 0:2] for f in forcestates], dim=-1)


def to_feature_noforce(smokestates):
    """burgers_train.py:84-90."""
    return torch.cat([s.velocity.staggered_tensor()[:, :-1, :-1, 0:2] for s in smokestates], dim=-1)



def periodic_diffusion_kernel(H: int, W: int, amount: float, device) -> torch.Tensor:
    """Real-space circular kernel of PhiFlow's periodic diffuse(): ifft2(exp(-(2 pi |k|)^2 amount))
    on an [H, W] component array (a scene constant, like the masks of a karman plan)."""
    ky = torch.fft.fftfreq(H, dtype=torch.float64).reshape(H, 1)
    kx = torch.fft.fftfreq(W, dtype=torch.float64).reshape(1, W)
    ker = torch.fft.ifft2(torch.exp(-(2 * np.pi) ** 2 * (ky ** 2 + kx ** 2) * amount)).real
    return ker.to(device=device, dtype=torch.float32).contiguous()


class BurgersVelocitySMAC:
    """BurgersVelocitySMAC(domain, batch_size=B): a periodic staggered velocity state."""

    def __init__(self, domain: Domain, velocity=None, batch_size=None, device=None):
        self.domain = domain
        B = 1 if batch_size is None else int(batch_size)
        Y, X = domain.resolution
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.velocity = velocity if velocity is not None else StaggeredGrid(
            [torch.zeros(B, Y + 1, X, device=dev), torch.zeros(B, Y, X + 1, device=dev)], domain.box)

    @property
    def _batch_size(self):
        return self.velocity._vy.shape[0]

    def copied_with(self, velocity=None):
        if velocity is None:
            return self
        if not isinstance(velocity, StaggeredGrid):
            t = torch.as_tensor(velocity, dtype=torch.float32) if isinstance(velocity, np.ndarray) else velocity
            velocity = StaggeredGrid(t.to(self.velocity._vy.device), self.domain.box)
        return BurgersVelocitySMAC(self.domain, velocity)


class BurgersTest:
    """Burgers.step: advect -> diffuse(viscosity*dt) [-> + dt*f] on a periodic domain."""

    def __init__(self, default_viscosity=0.1, viscosity=None, diffusion_substeps=1):
        self.viscosity = default_viscosity if viscosity is None else viscosity
        if diffusion_substeps != 1:
            raise SolError("only diffusion_substeps=1 (the reference's setting) is implemented")
        self._plans, self._kernels = {}, {}

    def _setup(self, v: BurgersVelocitySMAC, dt):
        Y, X = v.domain.resolution
        B = v._batch_size
        dev = v.velocity._vy.device
        key = (Y, X, B, str(dev))
        if key not in self._plans:
            self._plans[key] = Plan.periodic(Y, X, B, float(v.domain.dx[1]), device=dev)
        kk = (Y, X, float(dt), str(dev))
        if kk not in self._kernels:
            amount = self.viscosity * dt
            self._kernels[kk] = (periodic_diffusion_kernel(Y + 1, X, amount, dev), periodic_diffusion_kernel(Y, X + 1, amount, dev))
        return self._plans[key], self._kernels[kk]

    def step(self, v, dt=1.0, effects=()):
        plan, (ky, kx) = self._setup(v, dt)
        oy, ox = plan.burgers_step(v.velocity._vy.contiguous(), v.velocity._vx.contiguous(), float(dt), self.viscosity, ky, kx)
        return v.copied_with(velocity=StaggeredGrid([oy, ox], v.velocity.box))

    def step_with_f(self, v, f, dt=1.0):
        plan, (ky, kx) = self._setup(v, dt)
        oy, ox = plan.burgers_step(v.velocity._vy.contiguous(), v.velocity._vx.contiguous(), float(dt), self.viscosity, ky, kx,
                                   f.velocity._vy.contiguous(), f.velocity._vx.contiguous())
        return v.copied_with(velocity=StaggeredGrid([oy, ox], v.velocity.box))


# ---- PhiFlow-2 flavoured surface (karman-2d-phi2/karman_train.py:149-196) ---------------------------
class KarmanFlowPhi2:
    """``KarmanFlow(domain).step(density_in, velocity_in, re, res, ...) -> [density, velocity]`` — the signature of the
    PhiFlow-2 rewrite of the scene, on the same CUDA kernels.  The two in-repo differences of that script to the
    PhiFlow-1 one are honoured: viscosity is ``diffuse.explicit(velocity, dt*res^2/re, dt)`` in PHYSICAL units
    (:169; mapped onto the kernel's index-space alpha = dt^2*res^2/(re*dx^2)) and the inflow is added BEFORE the
    advection (:182).  Everything PhiFlow-2 does inside ``make_incompressible`` / ``semi_lagrangian`` is not pinned
    by code in the reference repository: this class is a signature-compatible surface, not a parity
    target.  ``domain`` is a phi_compat.Domain; density / velocity are CenteredGrid / StaggeredGrid look-alikes."""

    def __init__(self, domain: Domain):
        self.domain = domain
        Y, X = domain.resolution
        vn = np.zeros((Y + 1, X), dtype=np.float32)      # karman-2d-phi2/karman_train.py:153-159
        vn[0:2, 0:X - 1] = 1.0
        vn[:, 0:1] = 1.0
        vn[:, -1:] = 1.0
        self.vel_yBc, self.vel_yBcMask = vn, vn.copy()
        self.solve_info = {}
        self._plans = {}

    def _plan(self, B, dev) -> Plan:
        key = (B, str(dev))
        if key not in self._plans:
            Y, X = self.domain.resolution
            dx = float(self.domain.dx[1])
            lo = self.domain.box.lower
            CY, CX = np.meshgrid(lo[0] + (np.arange(Y) + 0.5) * dx, lo[1] + (np.arange(X) + 0.5) * dx, indexing="ij")
            solid = ((CY - 50.0) ** 2 + (CX - 50.0) ** 2) <= 10.0 ** 2                      # Sphere([50,50],10), :162
            self._inflow = torch.as_tensor(((CY >= 5) & (CY <= 10) & (CX >= 25) & (CX <= 75)).astype(np.float32), device=dev)   # Box[5:10,25:75], :161
            plan = Plan(Y, X, B, dx, _lib.SOL_BOUNDARY_OPEN, solid, None, self.vel_yBcMask, self.vel_yBc, device=dev)
            plan.set_cg(tol_abs=1e-5, tol_rel=0.0, max_it=2000, cluster=0)
            self._plans[key] = plan
        return self._plans[key]

    def step(self, density_in, velocity_in, re, res, buoyancy_factor=0, dt=1.0, make_input_divfree=False, make_output_divfree=True):
        if buoyancy_factor != 0 or make_input_divfree or not make_output_divfree:
            raise SolError("only buoyancy_factor=0, make_input_divfree=False, make_output_divfree=True (the script's settings) are implemented")
        vy, vx = velocity_in._vy.float().contiguous(), velocity_in._vx.float().contiguous()
        B, dev = vy.shape[0], vy.device
        plan = self._plan(B, dev)
        re_t = torch.as_tensor(np.asarray(re.detach().cpu() if isinstance(re, torch.Tensor) else re, dtype=np.float32).reshape(-1), device=dev)
        if re_t.numel() == 1 and B > 1:
            re_t = re_t.expand(B)
        dx = float(self.domain.dx[1])
        # alpha = dt * (dt*res^2/re) / dx^2  ==  dt * res_eff^2 / re   with   res_eff = res * sqrt(dt) / dx
        res_eff = float(res) * float(np.sqrt(dt)) / dx
        rho = (density_in._t.float() + self._inflow).contiguous()          # inflow before advection
        out = plan.step_fwd(re_t.contiguous(), vy, vx, rho=rho, dt=float(dt), res=res_eff)
        self.solve_info = {"pressure": CenteredGrid(out["p"], self.domain.box), "iterations": out["iters"]}
        return [CenteredGrid(out["rho"], self.domain.box), StaggeredGrid([out["vy"], out["vx"]], self.domain.box)]
