"""SOL training — the drop-in for karman-2d/karman_train.py (same flags; `--tf` is the output dir).
The TensorFlow graph + sess.run of the reference is one SolTrainer.train_step_host per iteration.
Multi-GPU: launch with torch.distributed.run; the simulations of a batch are sharded over ranks."""
import argparse
import logging
import os
import pickle
import random

import numpy as np
import torch

from . import device_index, load_init_weights, rank0_preprocess_then_barrier, reject_unimplemented
from .. import dist as sdist
from .. import engine
from ..dataset import PhifDataset
from ..trainer import SolTrainer, lr_schedule

log = logging.getLogger("karman_train")


def parse(argv=None):
    ap = argparse.ArgumentParser(description="Parameter Parser", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--gpu", default="0"); ap.add_argument("--cuda", action="store_true")
    ap.add_argument("--train", default=None); ap.add_argument("--skip-ds", action="store_true"); ap.add_argument("--only-ds", action="store_true")
    ap.add_argument("--log", default=None)
    ap.add_argument("-s", "--scale", default=4, type=int); ap.add_argument("-n", "--nsims", default=1, type=int)
    ap.add_argument("-b", "--sbatch", default=1, type=int); ap.add_argument("-t", "--simsteps", default=1500, type=int)
    ap.add_argument("-m", "--msteps", default=2, type=int); ap.add_argument("-e", "--epochs", default=10, type=int)
    ap.add_argument("--seed", default=None, type=int); ap.add_argument("-l", "--len", default=100, type=int)
    ap.add_argument("--model", default="mars_moon"); ap.add_argument("--reg-loss", action="store_true")
    ap.add_argument("--lr", default=1e-3, type=float); ap.add_argument("--adplr", action="store_true")
    ap.add_argument("--clip-grad", action="store_true"); ap.add_argument("--resume", default=-1, type=int)
    ap.add_argument("--inittf", default=None); ap.add_argument("--pretf", default=None)
    ap.add_argument("--tf", default="/tmp/phiflow/tf")
    return ap.parse_args(argv)


def main(argv=None):
    p = vars(parse(argv))
    logging.basicConfig(level=logging.INFO, **({"handlers": [logging.StreamHandler(), logging.FileHandler(p["log"])]} if p["log"] else {}))
    reject_unimplemented(p, ("pretf", "reg_loss"))          # karman_train.py:352-356,441-445: supervised-baseline weights / L2 regulariser
    rank, local, world = sdist.init_from_env("nccl")
    if world == 1:
        torch.cuda.set_device(device_index(p["gpu"]))
    if p["nsims"] % p["sbatch"]:
        p["nsims"] = (p["nsims"] // p["sbatch"]) * p["sbatch"]
    seed = 0 if p["seed"] is None else p["seed"]
    random.seed(seed); np.random.seed(seed)
    ds = rank0_preprocess_then_barrier(
        lambda skip: PhifDataset(p["train"], p["simsteps"], num_sims=p["nsims"], batch_size=p["sbatch"], print_fn=log.info,
                                 skip_preprocessing=skip, scale=p["scale"]), p["skip_ds"], rank, world)
    if p["only_ds"]:
        return
    if p["resume"] > 0:
        with open(p["tf"] + "/dataStats.pickle", "rb") as f:
            ds.dataStats = pickle.load(f)
    Y, X = ds.resolution
    lo, hi = sdist.shard_range(p["sbatch"], rank, world)          # this rank's simulations of every batch
    if hi - lo < 1:
        raise SystemExit("more ranks than simulations per batch")
    plan = engine.Plan.karman(int(Y), int(X), hi - lo, L=float(p["len"]))
    sig = (float(ds.dataStats["std"][1][0]), float(ds.dataStats["std"][1][1]), float(ds.dataStats["ext.std"][0]))
    trainer = SolTrainer(plan, p["msteps"], hi - lo, sig, lr=p["lr"], seed=seed, clip_grad=p["clip_grad"], model=p["model"])
    os.makedirs(p["tf"], exist_ok=True)
    if p["inittf"]:
        load_init_weights(trainer, p["inittf"], 3, p["model"])
    if p["resume"] < 1:
        if rank == 0:
            with open(p["tf"] + "/dataStats.pickle", "wb") as f:
                pickle.dump(ds.dataStats, f)
    else:
        trainer.load_state_dict(torch.load(p["tf"] + "/model_epoch{:04d}.pt".format(p["resume"])))
    current_lr = p["lr"]
    for j in range(p["epochs"]):
        ds.newEpoch(exclude_tail=p["msteps"])
        if j < p["resume"]:
            log.info("resume: skipping {} epoch".format(j + 1))       # keeps the shuffle RNG aligned (karman_train.py:485-490)
            continue
        current_lr = lr_schedule(j, current_lr) if p["adplr"] else p["lr"]
        for ib in range(ds.numOfBatchs):
            for i in range(ds.numOfSteps):
                re, vy0, vx0, gy, gx = PhifDataset.to_soa(ds.getData(consecutive_frames=p["msteps"], with_skip=1))
                t = lambda a, batch_axis=0: torch.from_numpy(np.ascontiguousarray(a[lo:hi] if batch_axis == 0 else a[:, lo:hi]))
                l2 = trainer.train_step_host(t(re), t(vy0), t(vx0), t(gy, 1), t(gx, 1), lr=current_lr)
                if rank == 0:
                    log.info("epoch {:03d}/{:03d}, batch {:03d}/{:03d}, step {:04d}/{:04d}: loss={}".format(
                        j + 1, p["epochs"], ib + 1, ds.numOfBatchs, i + 1, ds.numOfSteps, l2))
                ds.nextStep()
            ds.nextBatch()
        if j % 10 == 9 and rank == 0:
            torch.save(trainer.state_dict(), p["tf"] + "/model_epoch{:04d}.pt".format(j + 1))
    if rank == 0:
        torch.save(trainer.state_dict(), p["tf"] + "/model.pt")
        from ..phi_compat import CorrectionModel
        m = CorrectionModel(device=plan.device, model=p["model"]); m.flat.copy_(trainer.weights); m.save(p["tf"] + "/model.npz")   # Keras-ordered weights


if __name__ == "__main__":
    main()
