"""Inference rollout — the drop-in for burgers/burgers_apply.py (same flags): BurgersTest.step_with_f + model.predict +
add correction for `--simsteps` frames, writing velTf/corTf npz frames (burgers_apply.py:129-151)."""
import argparse
import glob
import logging
import os
import pickle

import numpy as np
import torch

from . import device_index
from .. import formats
from ..phi_compat import (PERIODIC, BurgersTest, BurgersVelocitySMAC, CorrectionModel, Domain, StaggeredGrid, box, burgers_to_feature,
                          to_feature_noforce, to_staggered)

log = logging.getLogger("burgers_apply")


def parse(argv=None):
    ap = argparse.ArgumentParser(description="Parameter Parser", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--gpu", default="0")
    ap.add_argument("-t", "--simsteps", default=200, type=int); ap.add_argument("-r", "--res", default=32, type=int)
    ap.add_argument("-l", "--len", default=96, type=int); ap.add_argument("--dt", default=1.0, type=float)
    ap.add_argument("--noforce", action="store_true")
    ap.add_argument("--initvH", default=None); ap.add_argument("--loadfH", default=None)
    ap.add_argument("-s", "--scale", default=4, type=int); ap.add_argument("-o", "--output", default=None)
    ap.add_argument("--stats", default="/tmp/phiflow/data/dataStats.pickle")
    ap.add_argument("--model", default="/tmp/phiflow/tf/model.npz", help="model.npz (Keras-ordered weights)")
    return ap.parse_args(argv)


def main(argv=None):
    p = vars(parse(argv))
    logging.basicConfig(level=logging.INFO)
    torch.cuda.set_device(device_index(p["gpu"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    res = p["res"]
    dm = Domain(resolution=[res, res], box=box([p["len"]] * 2), boundaries=PERIODIC)
    down = lambda path: torch.from_numpy(formats.downsample(formats.read_zipped_array(path), p["scale"], True).astype(np.float32)).to(dev)
    st = BurgersVelocitySMAC(dm, batch_size=1, device=dev)
    fc, fc_files = None, None
    if not p["noforce"]:
        fc_files = sorted(glob.glob(p["loadfH"])) if p["loadfH"] else None
        if not fc_files:
            raise SystemExit("--loadfH: no force files (use --noforce for unforced rollouts)")
        fc = st.copied_with(velocity=down(fc_files[0]))
    if p["initvH"]:
        st = st.copied_with(velocity=down(p["initvH"]))
    with open(p["stats"], "rb") as f:
        data_stats = pickle.load(f)
    model = CorrectionModel.load(p["model"], device=dev)
    std_v = [float(s) for s in data_stats["std"][0]]
    std_in = torch.tensor(std_v + ([] if p["noforce"] else [float(s) for s in data_stats["std"][1]]), dtype=torch.float32, device=dev)
    std_out = torch.tensor(std_v, dtype=torch.float32, device=dev)
    sim_path = formats.sim_dir(p["output"], 0) if p["output"] else None
    cv = StaggeredGrid([torch.zeros_like(st.velocity._vy), torch.zeros_like(st.velocity._vx)], dm.box)

    def write(i):
        if sim_path is None:
            return
        for name, grid in (("velTf", st.velocity), ("corTf", cv)):
            formats.write_zipped_array(os.path.join(sim_path, "%s_%06d.npz" % (name, i)), grid.staggered_tensor().cpu().numpy())

    write(0)
    simulator = BurgersTest()
    for i in range(1, p["simsteps"]):
        if not p["noforce"]:
            st = simulator.step_with_f(v=st, f=fc, dt=p["dt"])
            fc = fc.copied_with(velocity=down(fc_files[i]))          # the features see the NEXT force frame (burgers_apply.py:134-136)
            inputf = burgers_to_feature([st], [fc]) / std_in
        else:
            st = simulator.step(v=st, dt=p["dt"])
            inputf = to_feature_noforce([st]) / std_in
        cv = to_staggered(model.predict(inputf) * std_out, st.velocity.box)
        st = st.copied_with(velocity=st.velocity + cv)
        log.info("step {:06d}".format(i))
        write(i)
    return st


if __name__ == "__main__":
    main()
