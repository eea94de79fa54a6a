"""Inference rollout — the drop-in for karman-2d/karman_apply.py (same flags): simulator.step +
model.predict + add correction for `--simsteps` frames, writing denTf/velTf/corTf npz frames."""
import argparse
import logging
import os
import pickle

import numpy as np
import torch

from . import device_index
from .. import formats
from ..phi_compat import OPEN, CorrectionModel, Domain, Fluid, KarmanFlow, StaggeredGrid, box, to_feature, to_staggered, unstack_staggered_tensor

log = logging.getLogger("karman_apply")


def parse(argv=None):
    ap = argparse.ArgumentParser(description="Parameter Parser", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--gpu", default="0"); ap.add_argument("--cuda", action="store_true")
    ap.add_argument("-o", "--output", default="/tmp/phiflow/run"); ap.add_argument("-r", "--res", default=32, type=int)
    ap.add_argument("-l", "--len", default=100, type=int); ap.add_argument("--re", default=1e6, type=float)
    ap.add_argument("--initdH", default=None); ap.add_argument("--initvH", default=None)
    ap.add_argument("-t", "--simsteps", default=500, type=int); ap.add_argument("-s", "--scale", default=4, type=int)
    ap.add_argument("--stats", default="/tmp/phiflow/data/dataStats.pickle", help="dataStats.pickle of the training run")
    ap.add_argument("--model", default="/tmp/phiflow/tf/model.h5", help="Keras-ordered weights: model.npz, or model.h5 with the .npz beside it")
    ap.add_argument("--fused", action="store_true", help="run all frames in ONE library call (sol_unroll_rollout) instead of the eager loop")
    return ap.parse_args(argv)


def main(argv=None):
    p = vars(parse(argv))
    logging.basicConfig(level=logging.INFO)
    torch.cuda.set_device(device_index(p["gpu"]))
    res, L = p["res"], p["len"]
    st = Fluid(Domain(resolution=[res * 2, res], box=box[0:L * 2, 0:L], boundaries=OPEN), buoyancy_factor=0)
    if p["initvH"]:
        vn = torch.from_numpy(formats.downsample(formats.read_zipped_array(p["initvH"]), p["scale"], True).astype(np.float32)).cuda()
    else:
        vn = st.velocity.staggered_tensor()
        vn[..., 0] = 1.0
        vn[..., vn.shape[1] // 2 + 10:vn.shape[1] // 2 + 20, vn.shape[2] // 2 - 2:vn.shape[2] // 2 + 2, 1] = 1.0
    v0 = StaggeredGrid(unstack_staggered_tensor(vn), st.velocity.box)
    d0 = None
    if p["initdH"]:
        d0 = torch.from_numpy(formats.downsample(formats.read_zipped_array(p["initdH"]), p["scale"], False).astype(np.float32)).cuda()
    st = st.copied_with(density=d0, velocity=v0)
    bc = np.zeros(tuple(st.velocity.data[0].data.shape))
    bc[..., 0:2, 0:bc.shape[2] - 1, 0] = 1.0
    bc[..., 0:bc.shape[1], 0:1, 0] = 1.0
    bc[..., 0:bc.shape[1], -1:, 0] = 1.0
    velBCy, velBCyMask = bc, np.copy(bc)
    with open(p["stats"], "rb") as f:
        data_stats = pickle.load(f)
    model = CorrectionModel.load(p["model"])
    std_in = torch.tensor([*data_stats["std"][1], data_stats["ext.std"][0]], dtype=torch.float32, device="cuda")
    std_out = torch.tensor(data_stats["std"][1], dtype=torch.float32, device="cuda")
    sim_path = formats.sim_dir(p["output"], 0) if p["output"] else None
    cv = st.staggered_grid(name="corr", value=0)

    def write(i):
        if sim_path is None:
            return
        frames = [("denTf", st.density.data), ("velTf", st.velocity.staggered_tensor())]
        if not p["fused"] or i == 0:       # the fused rollout applies the correction in-kernel and never materialises it
            frames.append(("corTf", cv.staggered_tensor()))
        for name, arr in frames:
            formats.write_zipped_array(os.path.join(sim_path, "%s_%06d.npz" % (name, i)), arr.cpu().numpy())

    write(0)
    simulator = KarmanFlow()
    if p["fused"]:
        from .. import _lib, engine
        n = p["simsteps"] - 1
        plan = simulator._plan(st, velBCy, velBCyMask)
        model_id = _lib.SOL_MODEL_MERCURY if model.model == "mercury" else _lib.SOL_MODEL_MARS_MOON
        un = engine.Unroll(plan, 1, 1, (float(std_out[0]), float(std_out[1]), float(std_in[2])), model=model_id, with_density=True)
        re_t = torch.full((1,), float(p["re"]), device="cuda")
        pv, px, pr = un.rollout(model.flat, re_t, st.velocity._vy.contiguous(), st.velocity._vx.contiguous(), n, rho0=st.density._t.contiguous())
        for i in range(1, p["simsteps"]):
            vel = StaggeredGrid([pv[i - 1], px[i - 1]], st.velocity.box)
            st = st.copied_with(density=pr[i - 1], velocity=vel)
            write(i)          # corTf is not materialised by the fused rollout (the correction is applied in-kernel)
        return st
    for i in range(1, p["simsteps"]):
        st = simulator.step(st, re=p["re"], res=res, velBCy=velBCy, velBCyMask=velBCyMask)
        inputf = to_feature([st], p["re"]) / std_in
        cv_pred = model.predict(inputf) * std_out
        cv = to_staggered(cv_pred, st.velocity.box)
        st = st.copied_with(velocity=st.velocity + cv)
        write(i)
    return st


if __name__ == "__main__":
    main()
