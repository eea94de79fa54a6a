"""PhifDataset (karman-2d/karman_train.py:187-337) re-stated for the struct-of-arrays engine:
pre-loads (optionally down-sampled) frames, computes the dataStats the training graph normalises
with, and reproduces newEpoch / nextBatch / nextStep / getData including the shuffle order."""
from __future__ import annotations

import glob
import os
import random
from typing import List

import numpy as np

from . import formats


class PhifDataset:
    def __init__(self, dirpath, num_frames, num_sims=None, batch_size=1, print_fn=print, skip_preprocessing=False, scale=4):
        self.dataSims = sorted(glob.glob(dirpath + "/sim_0*"))[0:num_sims]
        self.pathsDen = [sorted(glob.glob(s + "/dens_0*.npz")) for s in self.dataSims]
        self.pathsVel = [sorted(glob.glob(s + "/velo_0*.npz")) for s in self.dataSims]
        self.dataFrms = [np.arange(num_frames) for _ in self.dataSims]
        self.batchSize = batch_size
        self.epoch, self.epochIdx, self.batchIdx, self.stepIdx = None, 0, 0, 0
        self.printFn = print_fn
        self.numOfSims = len(self.dataSims) if num_sims is None else num_sims
        self.numOfBatchs = self.numOfSims // self.batchSize
        self.numOfFrames = self.numOfSteps = num_frames
        if not skip_preprocessing:
            for j, _ in enumerate(self.dataSims):
                for i in range(num_frames):
                    for paths, stag in ((self.pathsDen, False), (self.pathsVel, True)):
                        out = self.filenameToDownscaled(paths[j][i])
                        if not os.path.isfile(out):
                            formats.write_zipped_array(out, formats.downsample(formats.read_zipped_array(paths[j][i]), scale, stag))
        self.dataPreloaded = {
            s: [(formats.read_zipped_array(self.filenameToDownscaled(self.pathsDen[j][i])),
                 formats.read_zipped_array(self.filenameToDownscaled(self.pathsVel[j][i]))) for i in range(num_frames)]
            for j, s in enumerate(self.dataSims)}
        self.resolution = self.dataPreloaded[self.dataSims[0]][0][0].shape[1:3]
        cat = lambda k, c=None: np.concatenate([
            np.absolute((self.dataPreloaded[s][i][k] if c is None else self.dataPreloaded[s][i][k][..., c]).reshape(-1))
            for s in self.dataSims for i in range(num_frames)])
        self.dataStats = {"std": (np.std(cat(0)), (np.std(cat(1, 0)), np.std(cat(1, 1))))}
        self.extConstChannelPerSim = {s: [formats.read_params(s)["re"]] for s in self.dataSims}
        self.dataStats["ext.std"] = [np.std([np.absolute(self.extConstChannelPerSim[s][0]) for s in self.dataSims])]

    @staticmethod
    def filenameToDownscaled(fname):
        return os.path.dirname(fname) + "/ds_" + os.path.basename(fname)

    def newEpoch(self, exclude_tail=0, shuffle_data=True):
        self.numOfSteps = self.numOfFrames - exclude_tail
        pairs: List = []
        for i, _ in enumerate(self.dataSims):
            pairs += [(i, s) for s in self.dataFrms[i][0:len(self.dataFrms[i]) - exclude_tail]]
        if shuffle_data:
            random.shuffle(pairs)
        self.epoch = [list(pairs[i * self.numOfSteps:(i + 1) * self.numOfSteps]) for i in range(self.batchSize * self.numOfBatchs)]
        self.epochIdx += 1
        self.batchIdx = 0
        self.stepIdx = 0

    def nextBatch(self):
        self.batchIdx += self.batchSize
        self.stepIdx = 0

    def nextStep(self):
        self.stepIdx += 1

    def getData(self, consecutive_frames, with_skip=1):
        """Returns [d_frames, v_frames, ext] like the reference, plus nothing else; use ``to_soa`` to feed
        SolTrainer.train_step_host."""
        pick = lambda i, j, k: self.dataPreloaded[self.dataSims[self.epoch[self.batchIdx + i][self.stepIdx][0]]][
            self.epoch[self.batchIdx + i][self.stepIdx][1] + j * with_skip][k]
        d = [np.concatenate([pick(i, j, 0) for i in range(self.batchSize)], axis=0) for j in range(consecutive_frames + 1)]
        v = [np.concatenate([pick(i, j, 1) for i in range(self.batchSize)], axis=0) for j in range(consecutive_frames + 1)]
        ext = [self.extConstChannelPerSim[self.dataSims[self.epoch[self.batchIdx + i][self.stepIdx][0]]][0] for i in range(self.batchSize)]
        return [d, v, ext]

    @staticmethod
    def to_soa(adata):
        """(re[B], vy0, vx0, gt_vy[m,...], gt_vx[m,...]) float32 numpy arrays for the trainer."""
        d, v, ext = adata
        vy0, vx0 = formats.unpack_staggered(v[0])
        gts = [formats.unpack_staggered(f) for f in v[1:]]
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        return (f32(ext), f32(vy0), f32(vx0), f32(np.stack([g[0] for g in gts])), f32(np.stack([g[1] for g in gts])))


class BurgersPhifDataset(PhifDataset):
    """PhifDataset of burgers/burgers_train.py:189-337: frames are (velocity, force) pairs (``velo_*.npz`` /
    ``forc_*.npz``, both staggered and down-sampled with downsample{scale}xSMAC), ``resolution`` is the centred grid
    size, dataStats['std'] = ((std|vy|, std|vx|), (std|fy|, std|fx|)).  Shuffle / batch / step logic is inherited."""

    def __init__(self, dirpath, num_frames, num_sims=None, batch_size=1, print_fn=print, skip_preprocessing=False, scale=4):
        self.dataSims = sorted(glob.glob(dirpath + "/sim_0*"))[0:num_sims]
        self.pathsVel = [sorted(glob.glob(s + "/velo_0*.npz")) for s in self.dataSims]
        self.pathsFrc = [sorted(glob.glob(s + "/forc_0*.npz")) for s in self.dataSims]
        self.dataFrms = [np.arange(num_frames) for _ in self.dataSims]
        self.batchSize = batch_size
        self.epoch, self.epochIdx, self.batchIdx, self.stepIdx = None, 0, 0, 0
        self.printFn = print_fn
        self.numOfSims = len(self.dataSims) if num_sims is None else num_sims
        self.numOfBatchs = self.numOfSims // self.batchSize
        self.numOfFrames = self.numOfSteps = num_frames
        if not skip_preprocessing:
            for j, _ in enumerate(self.dataSims):
                for i in range(num_frames):
                    for paths in (self.pathsVel, self.pathsFrc):
                        formats.write_zipped_array(self.filenameToDownscaled(paths[j][i]),
                                                   formats.downsample(formats.read_zipped_array(paths[j][i]), scale, True))
        self.dataPreloaded = {
            s: [(formats.read_zipped_array(self.filenameToDownscaled(self.pathsVel[j][i])),
                 formats.read_zipped_array(self.filenameToDownscaled(self.pathsFrc[j][i]))) for i in range(num_frames)]
            for j, s in enumerate(self.dataSims)}
        self.resolution = [v - 1 for v in self.dataPreloaded[self.dataSims[0]][0][0].shape[1:3]]     # SMAC grid -> centred size
        cat = lambda k, c: np.concatenate([np.absolute(self.dataPreloaded[s][i][k][..., c].reshape(-1))
                                           for s in self.dataSims for i in range(num_frames)])
        self.dataStats = {"std": ((np.std(cat(0, 0)), np.std(cat(0, 1))), (np.std(cat(1, 0)), np.std(cat(1, 1))))}

    def getData(self, consecutive_frames, with_skip=1):
        """[v_frames, f_frames] (m+1 packed tensors each), burgers_train.py:288-311."""
        pick = lambda i, j, k: self.dataPreloaded[self.dataSims[self.epoch[self.batchIdx + i][self.stepIdx][0]]][
            self.epoch[self.batchIdx + i][self.stepIdx][1] + j * with_skip][k]
        v = [np.concatenate([pick(i, j, 0) for i in range(self.batchSize)], axis=0) for j in range(consecutive_frames + 1)]
        f = [np.concatenate([pick(i, j, 1) for i in range(self.batchSize)], axis=0) for j in range(consecutive_frames + 1)]
        return [v, f]

    @staticmethod
    def to_soa(adata):
        """(vy0, vx0, f_vy[m], f_vx[m], gt_vy[m], gt_vx[m]) float32 arrays for BurgersTrainer.train_step_host: state =
        frame 0, forces = frames 0..m-1, ground truth = frames 1..m (burgers_train.py:476-480)."""
        v, f = adata
        m = len(v) - 1
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        vy0, vx0 = formats.unpack_staggered(v[0])
        gts = [formats.unpack_staggered(t) for t in v[1:]]
        frc = [formats.unpack_staggered(t) for t in f[:m]]
        return (f32(vy0), f32(vx0), f32(np.stack([a[0] for a in frc])), f32(np.stack([a[1] for a in frc])),
                f32(np.stack([a[0] for a in gts])), f32(np.stack([a[1] for a in gts])))
