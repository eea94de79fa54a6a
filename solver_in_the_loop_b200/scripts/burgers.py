"""Data generation for forced Burgers — the GPU counterpart of burgers/burgers.py (same flags, same output layout):
``sim_%06d/{velo,forc}_%06d.npz`` + ``params.pickle``.  The rollout is BurgersTest.step_with_f on the CUDA engine.

Two pieces of that script live inside PhiFlow and are restated here from recollection [PHI-RECALL] (they only shape the random
training data, not the solver): ``SinPotential`` (a travelling sine wave ``amplitude * sin(k.x + phase)`` per force, 20 forces summed,
phases advanced by ``dt*omega`` per step, burgers.py:89-114) and ``math.randfreq`` (a smooth random initial field with a
``(1/(|k|+1))^8`` spectrum, burgers.py:121).  The random draws follow the reference's order on numpy's global generator."""
import argparse
import glob
import logging
import os

import numpy as np
import torch

from . import device_index
from .. import formats
from ..phi_compat import PERIODIC, BurgersTest, BurgersVelocitySMAC, Domain, StaggeredGrid, box

log = logging.getLogger("burgers")


def parse(argv=None):
    ap = argparse.ArgumentParser(description="Parameter Parser", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--gpu", default="0"); ap.add_argument("--cuda", action="store_true")
    ap.add_argument("-o", "--output", default=None); ap.add_argument("--thumb", action="store_true", help="ignored (no image writer offline)")
    ap.add_argument("--noforce", action="store_true")
    ap.add_argument("-s", "--skipsteps", default=0, type=int); ap.add_argument("-t", "--simsteps", default=200, type=int)
    ap.add_argument("-r", "--res", default=32, type=int); ap.add_argument("-l", "--len", default=32, type=int)
    ap.add_argument("--dt", default=0.1, type=float)
    ap.add_argument("--initvH", default=None); ap.add_argument("--loadfH", default=None)
    ap.add_argument("-d", "--scale", default=4, type=int); ap.add_argument("--seed", default=0, type=int)
    ap.add_argument("--sim-index", default=None, type=int, help="index of the sim_%%06d folder (default: next free)")
    return ap.parse_args(argv)


def randfreq(shape, power=8):
    """Smooth random field [PHI-RECALL math.randfreq]: complex white noise shaped by (1/(|k|+1))^power in index frequencies."""
    H, W = shape[1:3]
    out = np.zeros(shape, dtype=np.float32)
    ky = np.abs(np.fft.fftfreq(H) * H).reshape(H, 1)
    kx = np.abs(np.fft.fftfreq(W) * W).reshape(1, W)
    k = np.sqrt(ky ** 2 + kx ** 2)
    fac = (1.0 / (k + 1.0)) ** power * power * np.sqrt(0.5 * (H + W))
    for c in range(shape[-1]):
        f0 = np.random.randn(H, W) + 1j * np.random.randn(H, W)
        out[0, :, :, c] = np.real(np.fft.ifft2(f0 * fac)) * np.sqrt(H * W)
    return out


class SinForces:
    """num_forces travelling sine waves (burgers.py:89-114)."""

    def __init__(self, num_forces=20):
        self.k, self.amp, self.phase, self.omega = [], [], [], []
        for _ in range(num_forces):
            angle = np.random.random() * np.pi
            unit = np.array([np.sin(angle), np.cos(angle)])                 # (y, x)
            self.k.append((np.random.random() + 1.0) * 0.8 * unit)
            self.amp.append((np.random.random(2) - 0.5) * 0.3)              # (f_y, f_x)
            self.phase.append(np.random.random() * 2 * np.pi)
            self.omega.append(np.random.random() * 0.8 - 0.4)

    def advance(self, dt):
        self.phase = [p + dt * o for p, o in zip(self.phase, self.omega)]

    def sample(self, res, dx):
        """Staggered sample: component y at ((j)dx, (i+.5)dx) on [res+1, res], component x at ((j+.5)dx, (i)dx) on [res, res+1]."""
        jy, iy = np.meshgrid(np.arange(res + 1) * dx, (np.arange(res) + 0.5) * dx, indexing="ij")
        jx, ix = np.meshgrid((np.arange(res) + 0.5) * dx, np.arange(res + 1) * dx, indexing="ij")
        fy = np.zeros_like(jy); fx = np.zeros_like(jx)
        for k, a, p in zip(self.k, self.amp, self.phase):
            fy += a[0] * np.sin(k[0] * jy + k[1] * iy + p)
            fx += a[1] * np.sin(k[0] * jx + k[1] * ix + p)
        return fy.astype(np.float32)[None], fx.astype(np.float32)[None]


def main(argv=None):
    p = vars(parse(argv))
    logging.basicConfig(level=logging.INFO)
    torch.cuda.set_device(device_index(p["gpu"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    np.random.seed(p["seed"])
    res, dx = p["res"], p["len"] / p["res"]
    forces = SinForces(20)
    dm = Domain([res, res], box=box([p["len"]] * 2), boundaries=PERIODIC)
    down = lambda path: torch.from_numpy(formats.downsample(formats.read_zipped_array(path), p["scale"], True).astype(np.float32)).to(dev)
    grid = lambda fy, fx: StaggeredGrid([torch.from_numpy(fy).to(dev), torch.from_numpy(fx).to(dev)], dm.box)
    st = BurgersVelocitySMAC(dm, batch_size=1, device=dev).copied_with(velocity=torch.from_numpy(randfreq((1, res + 1, res + 1, 2)) * 2).to(dev))
    fc = BurgersVelocitySMAC(dm, batch_size=1, device=dev).copied_with(velocity=grid(*forces.sample(res, dx)))
    fc_files = sorted(glob.glob(p["loadfH"])) if p["loadfH"] else None
    if p["initvH"]:
        st = st.copied_with(velocity=down(p["initvH"]))
    if fc_files:
        fc = fc.copied_with(velocity=down(fc_files[0]))
    sim_path = None
    if p["output"]:
        idx = p["sim_index"]
        if idx is None:
            idx = 0
            while os.path.isdir(formats.sim_dir(p["output"], idx)):
                idx += 1
        sim_path = formats.sim_dir(p["output"], idx)
        formats.write_params(sim_path, p)

    def write(frame):
        if sim_path is None:
            return
        for name, s in (("velo", st), ("forc", fc)):
            formats.write_zipped_array(os.path.join(sim_path, "%s_%06d.npz" % (name, frame)), s.velocity.staggered_tensor().cpu().numpy())

    if p["skipsteps"] == 0:
        write(0)
    simulator = BurgersTest()
    for i in range(1, max(p["simsteps"] + p["skipsteps"], 1)):
        st = simulator.step(v=st, dt=p["dt"]) if p["noforce"] else simulator.step_with_f(v=st, f=fc, dt=p["dt"])
        if fc_files is None:
            forces.advance(p["dt"])
            fc = fc.copied_with(velocity=grid(*forces.sample(res, dx)))
        else:
            fc = fc.copied_with(velocity=down(fc_files[i]))
        if p["skipsteps"] <= i:
            write(max(i - p["skipsteps"], 0))
    log.info("done: %s", sim_path)
    return st


if __name__ == "__main__":
    main()
