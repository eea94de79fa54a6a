"""The reference's Makefile pipelines end to end on the GPU through the CLI drop-ins (same flags as karman.py /
karman_train.py / karman_apply.py / burgers_train.py / burgers_apply.py): data generation -> SOL training -> apply
(karman-2d/Makefile:20-23, 78-80, 119-128; burgers/Makefile:75-77), at toy sizes."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_karman_generate_train_apply(cuda_device, tmp_path):
    from solver_in_the_loop_b200 import formats
    from solver_in_the_loop_b200.scripts import karman, karman_apply, karman_train
    data, tf, run = str(tmp_path / "data"), str(tmp_path / "tf"), str(tmp_path / "run")
    for k, re in enumerate((1.6e5, 3.2e5)):      # two hi-res simulations (128x64 -> down-sampled 2x to 64x32)
        karman.main(["-o", data, "-r", "64", "-l", "100", "--re", str(re), "-t", "14", "--skipsteps", "5", "--sim-index", str(k)])
    frames = sorted(glob.glob(data + "/sim_000000/velo_0*.npz"))
    assert len(frames) == 8 and formats.read_zipped_array(frames[0]).shape == (1, 129, 65, 2)
    for model in ("mars_moon", "mercury"):
        out = tf + "_" + model
        karman_train.main(["--train", data, "--tf", out, "-s", "2", "-n", "2", "-b", "2", "-t", "8", "-m", "2", "-e", "1", "--lr", "1e-4",
                           "-l", "100", "--model", model, "--seed", "0"])
        assert os.path.isfile(out + "/model.npz") and os.path.isfile(out + "/dataStats.pickle")
        st = karman_apply.main(["-o", run + "_" + model, "-r", "32", "-l", "100", "--re", "3.2e5", "-t", "6", "-s", "2",
                                "--initvH", frames[0], "--stats", out + "/dataStats.pickle", "--model", out + "/model.npz"])
        v = st.velocity.staggered_tensor()
        assert v.shape == (1, 65, 33, 2) and torch.isfinite(v).all()
        assert len(glob.glob(run + "_" + model + "/sim_000000/velTf_0*.npz")) == 6
        # the same rollout in one library call (sol_unroll_rollout)
        st2 = karman_apply.main(["-r", "32", "-l", "100", "--re", "3.2e5", "-t", "6", "-s", "2", "--initvH", frames[0],
                                 "--stats", out + "/dataStats.pickle", "--model", out + "/model.npz", "--fused"])
        v2 = st2.velocity.staggered_tensor()
        assert float((v2 - v).norm() / v.norm()) < 1e-5


def test_karman_generated_trajectory_matches_oracle(cuda_device, tmp_path):
    """karman.py (data generation, karman-2d/Makefile:20-23): every written frame of the rollout against the oracle's trajectory from
    the same warm start (karman.py:107-110) and Reynolds number — not just shapes."""
    from oracle import sol_oracle as so
    from solver_in_the_loop_b200 import formats
    from solver_in_the_loop_b200.scripts import karman
    data = str(tmp_path / "traj")
    res, L, Re, T = 32, 100, 3.2e5, 13
    st = karman.main(["-o", data, "-r", str(res), "-l", str(L), "--re", str(Re), "-t", str(T), "--skipsteps", "0", "--sim-index", "0"])
    geom = so.KarmanGeom(2 * res, res, L)
    rho, vy, vx = so.warm_start(geom, 1)
    re_t = torch.tensor([Re], dtype=torch.float64)
    rel = lambda a, b: float((torch.as_tensor(a).double().cpu() - b).norm() / (b.norm() + 1e-30))
    Y, X = 2 * res, res
    worst = 0.0
    for i in range(T):
        if i > 0:
            rho, vy, vx = so.karman_step(rho, vy, vx, re_t, geom)
        v = formats.read_zipped_array(data + "/sim_000000/velo_%06d.npz" % i)
        d = formats.read_zipped_array(data + "/sim_000000/dens_%06d.npz" % i)
        assert v.shape == (1, Y + 1, X + 1, 2) and d.shape == (1, Y, X, 1)
        e = (rel(v[0, :, :X, 0], vy[0]), rel(v[0, :Y, :, 1], vx[0]), rel(d[0, ..., 0], rho[0]))
        worst = max(worst, *e)
        assert max(e) < 1e-5, (i, e)
    print("generated trajectory vs oracle, worst relative L2 over %d frames: %.2e" % (T, worst))
    vfin = st.velocity.staggered_tensor()
    assert rel(vfin[0, :, :X, 0], vy[0]) < 1e-5 and rel(vfin[0, :Y, :, 1], vx[0]) < 1e-5


def test_burgers_generate_train_apply(cuda_device, tmp_path):
    """burgers/Makefile:20-23, 75-77: burgers.py (20 travelling sine forces, random smooth initial state) -> burgers_train.py -> burgers_apply.py."""
    from solver_in_the_loop_b200 import formats
    from solver_in_the_loop_b200.scripts import burgers, burgers_apply, burgers_train
    data, tf, run = str(tmp_path / "bdata"), str(tmp_path / "btf"), str(tmp_path / "brun")
    frames, dt = 6, 0.1
    for s in range(2):       # "hi-res" 64x64 trajectories, down-sampled 2x by the dataset
        burgers.main(["-o", data, "-r", "64", "-l", "32", "--dt", str(dt), "--skipsteps", "3", "-t", str(frames), "--seed", str(s), "--sim-index", str(s)])
    v0 = formats.read_zipped_array(data + "/sim_000000/velo_000000.npz"); f5 = formats.read_zipped_array(data + "/sim_000001/forc_000005.npz")
    assert v0.shape == (1, 65, 65, 2) and f5.shape == (1, 65, 65, 2) and np.isfinite(v0).all() and 0.05 < np.abs(f5).max() < 3.0
    assert not np.allclose(formats.read_zipped_array(data + "/sim_000000/forc_000000.npz"), formats.read_zipped_array(data + "/sim_000000/forc_000005.npz"))
    tr = burgers_train.main(["--train", data, "--tf", tf, "-s", "2", "-n", "2", "-b", "2", "-t", str(frames), "-m", "2", "-e", "2", "--dt", str(dt),
                             "-l", "32", "--lr", "1e-4", "--seed", "0"])
    assert tr.t == 2 * (frames - 2) and os.path.isfile(tf + "/model.npz") and os.path.isfile(tf + "/model_epoch0001.pt")
    st = burgers_apply.main(["-o", run, "-r", "32", "-l", "32", "--dt", str(dt), "-t", "5", "-s", "2", "--initvH", data + "/sim_000000/velo_000000.npz",
                             "--loadfH", data + "/sim_000000/forc_0*.npz", "--stats", tf + "/dataStats.pickle", "--model", tf + "/model.npz"])
    v = st.velocity.staggered_tensor()
    assert v.shape == (1, 33, 33, 2) and torch.isfinite(v).all()
    assert len(glob.glob(run + "/sim_000000/velTf_0*.npz")) == 5
