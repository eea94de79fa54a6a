"""Host-side formats and dataset logic of the reference pipelines, CPU only."""
import os
import random

import numpy as np

from solver_in_the_loop_b200 import formats
from solver_in_the_loop_b200.dataset import PhifDataset


def test_zipped_array_roundtrip_reverses_channels(tmp_path):
    v = np.random.default_rng(0).standard_normal((1, 9, 5, 2)).astype(np.float32)
    p = str(tmp_path / "sim_000000" / "velo_000003.npz")
    formats.write_zipped_array(p, v)
    raw = np.load(p)["arr_0"]
    assert raw.shape == (9, 5, 2) and np.array_equal(raw[..., 0], v[0, ..., 1])        # (u, v) on disk
    assert np.array_equal(formats.read_zipped_array(p), v)
    d = np.random.default_rng(1).standard_normal((1, 8, 4, 1)).astype(np.float32)
    formats.write_zipped_array(str(tmp_path / "d.npz"), d)
    assert np.array_equal(formats.read_zipped_array(str(tmp_path / "d.npz")), d)


def test_pack_unpack_and_downsample():
    rng = np.random.default_rng(2)
    vy = rng.standard_normal((2, 9, 4)); vx = rng.standard_normal((2, 8, 5))
    t = formats.pack_staggered(vy, vx)
    assert t.shape == (2, 9, 5, 2) and np.all(t[:, :, -1, 0] == 0) and np.all(t[:, -1, :, 1] == 0)
    y2, x2 = formats.unpack_staggered(t)
    assert np.array_equal(y2, vy) and np.array_equal(x2, vx)
    lo = formats.downsample2x_staggered(t)
    assert lo.shape == (2, 5, 3, 2)
    assert np.allclose(lo[:, :, :-1, 0], 0.5 * (vy[:, ::2, 0::2] + vy[:, ::2, 1::2]))
    # a divergence-free uniform flow stays uniform
    u = formats.pack_staggered(np.ones((1, 9, 4)), np.zeros((1, 8, 5)))
    assert np.allclose(formats.downsample(u, 4, True)[:, :, :-1, 0], 1.0)
    c = rng.standard_normal((1, 8, 4, 1))
    assert np.allclose(formats.downsample2x_centered(c)[0, 0, 0, 0], c[0, 0:2, 0:2, 0].mean())


def test_dataset_epoch_logic_matches_reference(tmp_path):
    root = str(tmp_path / "set")
    rng = np.random.default_rng(3)
    frames, sims = 6, 4
    for s in range(sims):
        sd = formats.sim_dir(root, s)
        formats.write_params(sd, {"re": 1.0e5 * 2 ** s})
        for f in range(frames):
            formats.write_zipped_array(os.path.join(sd, "dens_%06d.npz" % f), rng.standard_normal((1, 8, 4, 1)).astype(np.float32))
            formats.write_zipped_array(os.path.join(sd, "velo_%06d.npz" % f), rng.standard_normal((1, 9, 5, 2)).astype(np.float32))
    ds = PhifDataset(root, frames, num_sims=sims, batch_size=2, print_fn=lambda *_: None, scale=2)
    assert tuple(ds.resolution) == (4, 2) and ds.numOfBatchs == 2
    assert os.path.isfile(os.path.join(formats.sim_dir(root, 0), "ds_velo_000000.npz"))
    assert ds.dataStats["ext.std"][0] == np.std([1e5, 2e5, 4e5, 8e5])
    random.seed(0)
    ds.newEpoch(exclude_tail=2)
    assert ds.numOfSteps == 4 and len(ds.epoch) == 4 and all(len(e) == 4 for e in ds.epoch)
    flat = sorted(p for e in ds.epoch for p in e)
    assert flat == sorted((i, s) for i in range(sims) for s in range(4))               # a permutation of all (sim, step) pairs
    d, v, ext = ds.getData(consecutive_frames=2)
    assert len(v) == 3 and v[0].shape == (2, 5, 3, 2) and len(ext) == 2
    re, vy0, vx0, gy, gx = PhifDataset.to_soa([d, v, ext])
    assert vy0.shape == (2, 5, 2) and vx0.shape == (2, 4, 3) and gy.shape == (2, 2, 5, 2) and re.dtype == np.float32
    ds.nextStep(); ds.nextBatch()
    assert ds.batchIdx == 2 and ds.stepIdx == 0


def test_burgers_dataset_matches_reference_logic(tmp_path):
    """burgers/burgers_train.py:189-337: (velocity, force) frame pairs, SMAC resolution - 1, per-component stds, and the
    struct-of-arrays view BurgersTrainer consumes (state = frame 0, forces = frames 0..m-1, ground truth = frames 1..m)."""
    from solver_in_the_loop_b200.dataset import BurgersPhifDataset
    root = str(tmp_path / "bset")
    rng = np.random.default_rng(5)
    frames, sims, m = 6, 2, 2
    raw = {}
    for s in range(sims):
        sd = formats.sim_dir(root, s)
        for f in range(frames):
            for name in ("velo", "forc"):
                a = rng.standard_normal((1, 9, 9, 2)).astype(np.float32)
                raw[(s, f, name)] = a
                formats.write_zipped_array(os.path.join(sd, "%s_%06d.npz" % (name, f)), a)
    ds = BurgersPhifDataset(root, frames, num_sims=sims, batch_size=2, print_fn=lambda *_: None, scale=2)
    assert list(ds.resolution) == [4, 4] and ds.numOfBatchs == 1
    lo = formats.downsample(raw[(0, 0, "velo")], 2, True)
    assert np.allclose(ds.dataPreloaded[ds.dataSims[0]][0][0], lo)
    allv0 = np.concatenate([np.abs(formats.downsample(raw[(s, f, "velo")], 2, True)[..., 0]).reshape(-1) for s in range(sims) for f in range(frames)])
    assert np.isclose(ds.dataStats["std"][0][0], np.std(allv0))
    random.seed(1)
    ds.newEpoch(exclude_tail=m)
    v, f = ds.getData(consecutive_frames=m)
    assert len(v) == m + 1 and len(f) == m + 1 and v[0].shape == (2, 5, 5, 2)
    vy0, vx0, fy, fx, gy, gx = BurgersPhifDataset.to_soa([v, f])
    assert vy0.shape == (2, 5, 4) and vx0.shape == (2, 4, 5) and fy.shape == (m, 2, 5, 4) and gx.shape == (m, 2, 4, 5)
    assert np.array_equal(gy[0], v[1][:, :, :-1, 0]) and np.array_equal(fx[1], f[1][:, :-1, :, 1]) and np.array_equal(vy0, v[0][:, :, :-1, 0])
