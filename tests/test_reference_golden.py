"""Parity against REAL PhiFlow outputs — active only when tests/golden/phiflow_*.npz exist (written by
tests/golden/make_reference_golden.py on a machine where phiflow==1.5.1 is installed; absent in the offline build container, where
these tests skip and the parity status stays "unpinned")."""
import os

import numpy as np
import pytest
import torch

from oracle import sol_oracle as so

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KARMAN = os.path.join(HERE, "phiflow_karman_64x32.npz")
BURGERS = os.path.join(HERE, "phiflow_burgers_32x32.npz")


def rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.skipif(not os.path.exists(KARMAN), reason="no PhiFlow-generated karman fixture (see tests/golden/make_reference_golden.py)")
def test_oracle_matches_phiflow_karman_steps():
    g = np.load(KARMAN)
    res, L = int(g["res"]), float(g["L"])
    geom = so.KarmanGeom(2 * res, res, L)
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    vy, vx = t(g["v0"][:, :, :-1, 0]), t(g["v0"][:, :-1, :, 1])
    rho = t(g["d0"][..., 0])
    re = torch.tensor([float(g["re"])], dtype=torch.float64)
    for i in (1, 2, 3):
        rho, vy, vx = so.karman_step(rho, vy, vx, re, geom)
        ref = g["v%d" % i]
        assert rel(vy.numpy(), ref[:, :, :-1, 0]) < 1e-5 and rel(vx.numpy(), ref[:, :-1, :, 1]) < 1e-4, "step %d" % i
        assert rel(rho.numpy(), g["d%d" % i][..., 0]) < 1e-5, "density step %d" % i


@pytest.mark.skipif(not os.path.exists(BURGERS), reason="no PhiFlow-generated burgers fixture (see tests/golden/make_reference_golden.py)")
def test_oracle_matches_phiflow_burgers_steps():
    g = np.load(BURGERS)
    R, L, dt = int(g["R"]), float(g["L"]), float(g["dt"])
    t = lambda a: torch.tensor(np.asarray(a), dtype=torch.float64)
    vy, vx = t(g["v0"][:, :, :-1, 0]), t(g["v0"][:, :-1, :, 1])
    for i in (1, 2, 3):
        vy, vx = so.burgers_step(vy, vx, dt, L / R, 0.1)
        ref = g["v%d" % i]
        assert rel(vy.numpy(), ref[:, :, :-1, 0]) < 1e-5 and rel(vx.numpy(), ref[:, :-1, :, 1]) < 1e-5, "step %d" % i
