"""Data generation — the GPU counterpart of karman-2d/karman.py (same flags, same output layout):
``sim_%06d/{dens,velo}_%06d.npz`` + ``params.pickle``.  The rollout is KarmanFlow.step on the CUDA
engine (phi_compat), one launch sequence per frame."""
import argparse
import logging
import os

import numpy as np
import torch

from . import device_index
from .. import formats
from ..phi_compat import OPEN, Domain, Fluid, KarmanFlow, StaggeredGrid, box, unstack_staggered_tensor

log = logging.getLogger("karman")


def parse(argv=None):
    ap = argparse.ArgumentParser(description="Parameter Parser", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--gpu", default="0")
    ap.add_argument("-o", "--output", default=None, help="path to an output directory (a sim_%%06d folder is created inside)")
    ap.add_argument("--thumb", action="store_true", help="ignored (no image writer offline)")
    ap.add_argument("--cuda", action="store_true", help="accepted for compatibility; the CUDA engine is always used")
    ap.add_argument("-r", "--res", default=32, type=int, help="resolution of the reference axis")
    ap.add_argument("-l", "--len", default=100, type=int, help="length of the reference axis")
    ap.add_argument("--re", default=1e6, type=float, help="Reynolds number")
    ap.add_argument("--initdH", default=None); ap.add_argument("--initvH", default=None)
    ap.add_argument("-t", "--simsteps", default=1500, type=int)
    ap.add_argument("-s", "--skipsteps", default=999, type=int, help="skip first steps (vortices may not form)")
    ap.add_argument("-d", "--scale", default=4, type=int, help="down-sampling scale of hires (only with --initdH / --initvH)")
    ap.add_argument("--seed", default=0, type=int)
    ap.add_argument("--sim-index", default=None, type=int, help="index of the sim_%%06d folder (default: next free)")
    return ap.parse_args(argv)


def main(argv=None):
    p = vars(parse(argv))
    logging.basicConfig(level=logging.INFO)
    torch.cuda.set_device(device_index(p["gpu"]))
    np.random.seed(p["seed"])
    res, L = p["res"], p["len"]
    st = Fluid(Domain(resolution=[res * 2, res], box=box[0:L * 2, 0:L], boundaries=OPEN), buoyancy_factor=0)
    if p["initvH"]:      # karman.py:100-104: down-sampled hi-res frame (read_zipped_array reverts the channel order)
        vn = torch.from_numpy(formats.downsample(formats.read_zipped_array(p["initvH"]), p["scale"], True).astype(np.float32)).cuda()
    else:
        vn = st.velocity.staggered_tensor()
        vn[..., 0] = 1.0
        vn[..., vn.shape[1] // 2 + 10:vn.shape[1] // 2 + 20, vn.shape[2] // 2 - 2:vn.shape[2] // 2 + 2, 1] = 1.0
    d0 = None
    if p["initdH"]:
        d0 = torch.from_numpy(formats.downsample(formats.read_zipped_array(p["initdH"]), p["scale"], False).astype(np.float32)).cuda()
    st = st.copied_with(velocity=StaggeredGrid(unstack_staggered_tensor(vn), st.velocity.box), **({"density": d0} if d0 is not None else {}))
    bc = np.zeros(tuple(st.velocity.data[0].data.shape))
    bc[..., 0:2, 0:bc.shape[2] - 1, 0] = 1.0
    bc[..., 0:bc.shape[1], 0:1, 0] = 1.0
    bc[..., 0:bc.shape[1], -1:, 0] = 1.0
    velBCy, velBCyMask = bc, np.copy(bc)
    sim_path = None
    if p["output"]:
        idx = p["sim_index"]
        if idx is None:
            idx = 0
            while os.path.isdir(formats.sim_dir(p["output"], idx)):
                idx += 1
        sim_path = formats.sim_dir(p["output"], idx)
        formats.write_params(sim_path, p)
    simulator = KarmanFlow()

    def write(step):
        if sim_path is None:
            return
        dpath, vpath = formats.frame_paths(sim_path, step)
        formats.write_zipped_array(dpath, st.density.data.cpu().numpy())
        formats.write_zipped_array(vpath, st.velocity.staggered_tensor().cpu().numpy())

    if p["skipsteps"] == 0:
        write(0)
    for i in range(1, p["simsteps"]):
        st = simulator.step(st, re=p["re"], res=res, velBCy=velBCy, velBCyMask=velBCyMask)
        if p["skipsteps"] < i:
            write(i)
    log.info("done: %s", sim_path)
    return st


if __name__ == "__main__":
    main()
