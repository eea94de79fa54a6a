"""Solver-in-the-loop trainer: the public API a user of the reference's karman_train.py calls once
per training iteration (reference: the graph built at karman-2d/karman_train.py:393-457 and run by
``sess.run([summary, train_step, total_loss])`` at :502).

One process per GPU.  The batch of simulations is sharded over ranks (data parallel over
independent simulations); the only collective is ONE all-reduce(SUM) per optimiser step
on a single flat buffer holding the 260,354 correction-net gradients plus the msteps per-step
losses, over NCCL (gloo in the CPU tests).  PyTorch owns parameter storage, streams and the
process group; every FLOP of the step runs in libsol_b200.so.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib, engine
from .dist import DataParallelStep


def model_layers(model: str = "mars_moon", cin0: int = 3):
    """(Cin, Cout) of the Conv2D layers of model_mars_moon / model_mercury (karman_train.py:92-138)."""
    if model == "mars_moon":
        return [(cin0, 32)] + [(32, 32)] * 10 + [(32, 2)]
    if model == "mercury":
        return [(cin0, 32), (32, 64), (64, 2)]
    raise ValueError(model)


def glorot_uniform_params(model: str = "mars_moon", cin0: int = 3, seed: int = 0) -> torch.Tensor:
    """Keras Conv2D default init (glorot_uniform kernels, zero biases), flat Keras-ordered fp32
    buffer (karman_train.py:101-138)."""
    import math
    layers = model_layers(model, cin0)
    g = torch.Generator().manual_seed(seed)
    parts = []
    for ci, co in layers:
        lim = math.sqrt(6.0 / (25 * ci + 25 * co))
        parts.append(((torch.rand(5, 5, ci, co, generator=g, dtype=torch.float64) * 2 - 1) * lim).reshape(-1))
        parts.append(torch.zeros(co, dtype=torch.float64))
    return torch.cat(parts).float()


def lr_schedule(epoch: int, current_lr: float) -> float:
    """--adplr schedule (karman_train.py:146-163)."""
    if epoch == 23:
        return current_lr * 0.5
    if epoch in (21, 16, 11):
        return current_lr * 1e-1
    return current_lr


class SolTrainer(DataParallelStep):
    def __init__(self, plan: engine.Plan, msteps: int, batch: int, sig: Sequence[float], lr: float = 1e-4,
                 weights: Optional[torch.Tensor] = None, seed: int = 0, dt: float = 1.0, use_graph: bool = True,
                 clip_grad: bool = False, process_group=None, with_density: bool = False, cin0: int = 3, model: str = "mars_moon"):
        self.plan, self.msteps, self.batch = plan, int(msteps), int(batch)
        self.lr, self.clip_grad = float(lr), bool(clip_grad)
        self.model = model
        model_id = {"mars_moon": _lib.SOL_MODEL_MARS_MOON, "mercury": _lib.SOL_MODEL_MERCURY}[model]      # --model (karman_train.py:33)
        self.unroll = engine.Unroll(plan, msteps, batch, sig, dt=dt, model=model_id, cin0=cin0, with_density=with_density,
                                    use_graph=use_graph)
        n = self.unroll.nparams
        dev = plan.device
        w0 = glorot_uniform_params(model, cin0=cin0, seed=seed) if weights is None else weights
        self.weights = w0.to(device=dev, dtype=torch.float32).contiguous().clone()
        # flat all-reduce bucket: [gradients | per-step losses] (dist.DataParallelStep)
        self._init_bucket(n, self.msteps, dev, process_group)
        self.layer_shapes = model_layers(model, cin0)
        self.unroll.loss_steps = self.loss_steps
        self.adam_m = torch.zeros(n, device=dev)
        self.adam_v = torch.zeros(n, device=dev)
        # pinned staging for the host-facing call
        self._pin = None
        self._dev_in = None

    # ---- device-resident batch -----------------------------------------------------------------
    def train_step(self, re, vy0, vx0, gt_vy, gt_vx, rho0=None, lr: Optional[float] = None) -> torch.Tensor:
        """One optimiser step on device tensors.  Returns the (global) total loss as a 0-d device
        tensor: sum_i loss_i / msteps summed over all simulations of all ranks (karman_train.py:436)."""
        self.unroll.train_iter(self.weights, re, vy0, vx0, gt_vy, gt_vx, self.grad, rho0=rho0)
        return self.reduce_and_update(self.lr if lr is None else lr, self.clip_grad)

    def _adam(self, lr: float):
        engine.adam_tf1(self.weights, self.grad, self.adam_m, self.adam_v, self.t, lr)

    # ---- host-facing call (what karman_train.py's feed_dict does) --------------------------------
    def train_step_host(self, re, vy0, vx0, gt_vy, gt_vx, lr: Optional[float] = None) -> float:
        """Same step fed from HOST tensors (numpy-backed / pinned): copies the batch to the device,
        runs the step and reads the loss back — the per-iteration host<->device traffic of the
        reference's sess.run (feed state_0, Re, gt_1..m; fetch total_loss)."""
        srcs = (re, vy0, vx0, gt_vy, gt_vx)
        if self._pin is None:
            self._pin = [torch.empty(s.shape, dtype=torch.float32).pin_memory() for s in srcs]
            self._dev_in = [torch.empty(s.shape, dtype=torch.float32, device=self.plan.device) for s in srcs]
            self._loss_pin = torch.empty((), dtype=torch.float32).pin_memory()
        for p, s, d in zip(self._pin, srcs, self._dev_in):
            if s.is_pinned():
                d.copy_(s, non_blocking=True)
            else:
                p.copy_(s)
                d.copy_(p, non_blocking=True)
        loss = self.train_step(*self._dev_in, lr=lr)
        self._loss_pin.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self._loss_pin)

    def h2d_bytes_per_step(self) -> int:
        B, m, p = self.batch, self.msteps, self.plan
        return 4 * (B + (m + 1) * B * (p.NY + p.NX))

    def state_dict(self):
        return dict(weights=self.weights.cpu(), adam_m=self.adam_m.cpu(), adam_v=self.adam_v.cpu(), t=self.t)

    def load_state_dict(self, sd):
        self.weights.copy_(sd["weights"]); self.adam_m.copy_(sd["adam_m"]); self.adam_v.copy_(sd["adam_v"]); self.t = int(sd["t"])


class BurgersTrainer(SolTrainer):
    """The unrolled training iteration of burgers/burgers_train.py:379-437 (sess.run at :487): msteps x
    (BurgersTest.step_with_f -> correction net on [v, f_i]/std -> add), l2 loss on v/std_v, TF1 Adam.

    ``sig_v`` / ``sig_f``: dataStats['std'][0] / [1] (velocity / force std per component, y first).
    ``noforce=True`` is the --noforce variant (features = velocity only, plain ``step``).
    """

    def __init__(self, plan: engine.Plan, msteps: int, batch: int, sig_v: Sequence[float], sig_f: Sequence[float] = (1.0, 1.0),
                 viscosity: float = 0.1, dt: float = 0.1, noforce: bool = False, spectral_diffusion: bool = True, **kw):
        if plan.boundary != _lib.SOL_BOUNDARY_PERIODIC:
            raise engine.SolError("BurgersTrainer needs a periodic plan (Plan.periodic)")
        super().__init__(plan, msteps, batch, (sig_v[0], sig_v[1], 1.0), dt=dt, cin0=2 if noforce else 4, **kw)
        dev = plan.device
        self.noforce = bool(noforce)
        self.ky = self.kx = None
        if spectral_diffusion:
            from .phi_compat import periodic_diffusion_kernel
            self.ky = periodic_diffusion_kernel(plan.Y + 1, plan.X, viscosity * dt, dev)
            self.kx = periodic_diffusion_kernel(plan.Y, plan.X + 1, viscosity * dt, dev)
        # persistent force buffers: the library keeps their addresses (they are part of the CUDA-graph key)
        self.f_vy = None if noforce else torch.zeros(self.msteps, self.batch, plan.Y + 1, plan.X, device=dev)
        self.f_vx = None if noforce else torch.zeros(self.msteps, self.batch, plan.Y, plan.X + 1, device=dev)
        self.unroll.set_burgers(viscosity, self.ky, self.kx, self.f_vy, self.f_vx, sig_f)

    def train_step(self, vy0, vx0, f_vy, f_vx, gt_vy, gt_vx, lr: Optional[float] = None) -> torch.Tensor:
        if not self.noforce:
            self.f_vy.copy_(f_vy, non_blocking=True)
            self.f_vx.copy_(f_vx, non_blocking=True)
        return SolTrainer.train_step(self, None, vy0, vx0, gt_vy, gt_vx, lr=lr)

    def train_step_host(self, vy0, vx0, f_vy, f_vx, gt_vy, gt_vx, lr: Optional[float] = None) -> float:
        """Host-fed variant: the feed_dict of burgers_train.py:482-486 (state, forces, ground truth) -> loss."""
        srcs = [vy0, vx0, gt_vy, gt_vx]
        if self._dev_in is None:
            self._dev_in = [torch.empty(s.shape, dtype=torch.float32, device=self.plan.device) for s in srcs]
            self._loss_pin = torch.empty((), dtype=torch.float32).pin_memory()
        for s, d in zip(srcs, self._dev_in):
            d.copy_(s, non_blocking=True)
        d = self._dev_in
        loss = self.train_step(d[0], d[1], f_vy, f_vx, d[2], d[3], lr=lr)
        self._loss_pin.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(self._loss_pin)
