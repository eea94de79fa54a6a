"""On-disk formats of the reference pipelines (SURVEY.md §8b "On-disk", §8f row 2).

* ``sim_%06d/{dens,velo}_%06d.npz`` written by PhiFlow's ``Scene.write`` (karman-2d/karman.py:154-159):
  one array (key ``arr_0``), batch dim squeezed, vector channels stored in REVERSED order
  (on disk (u=x, v=y); in memory "v first", karman.py:104).
* ``params.pickle`` next to the frames (karman.py:136), read back for Re (karman_train.py:247-249).
* ``dataStats.pickle`` (karman_train.py:474).
Host-side numpy only (no GPU needed).
"""
from __future__ import annotations

import os
import pickle
from typing import Dict

import numpy as np


def read_zipped_array(filename: str) -> np.ndarray:
    """phi.data.fluidformat.read_zipped_array: returns [1, ..., C] with channels reversed back."""
    with np.load(filename) as f:
        array = f[f.files[-1]]
    if array.shape[0] != 1 or array.ndim == 1:
        array = np.expand_dims(array, axis=0)
    if array.shape[-1] != 1:
        array = array[..., ::-1]
    return np.ascontiguousarray(array)


def write_zipped_array(filename: str, array: np.ndarray) -> None:
    """phi.data.fluidformat.write_zipped_array."""
    array = np.asarray(array)
    if array.shape[0] == 1 and array.ndim > 1:
        array = array[0, ...]
    if array.shape[-1] != 1:
        array = array[..., ::-1]
    os.makedirs(os.path.dirname(os.path.abspath(filename)), exist_ok=True)
    # write-then-rename: a concurrent reader (another rank preloading the dataset) never sees a half-written zip
    tmp = "%s.tmp%d.npz" % (filename, os.getpid())
    np.savez_compressed(tmp, array)
    os.replace(tmp, filename if str(filename).endswith(".npz") else str(filename) + ".npz")


def frame_paths(sim_dir: str, step: int, names=("dens", "velo")):
    return [os.path.join(sim_dir, "%s_%06d.npz" % (n, step)) for n in names]


def sim_dir(root: str, index: int) -> str:
    return os.path.join(root, "sim_%06d" % index)


def write_params(sim_path: str, params: Dict) -> None:
    os.makedirs(sim_path, exist_ok=True)
    with open(os.path.join(sim_path, "params.pickle"), "wb") as f:
        pickle.dump(params, f)


def read_params(sim_path: str) -> Dict:
    with open(os.path.join(sim_path, "params.pickle"), "rb") as f:
        return pickle.load(f)


# ---- packed <-> struct-of-arrays -------------------------------------------------------------------
def pack_staggered(vy: np.ndarray, vx: np.ndarray) -> np.ndarray:
    """[B,Y+1,X], [B,Y,X+1] -> packed [B,Y+1,X+1,2], channel 0 = y (StaggeredGrid.staggered_tensor())."""
    B, Yp1, X = vy.shape
    t = np.zeros((B, Yp1, X + 1, 2), dtype=vy.dtype)
    t[:, :, :-1, 0] = vy
    t[:, :-1, :, 1] = vx
    return t


def unpack_staggered(t: np.ndarray):
    return np.ascontiguousarray(t[:, :, :-1, 0]), np.ascontiguousarray(t[:, :-1, :, 1])


# ---- down-sampling (karman_train.py:140-144) --------------------------------------------------------
def downsample2x_centered(d: np.ndarray) -> np.ndarray:
    """math.downsample2x on [B,Y,X,C]: mean of 2x2 blocks."""
    B, Y, X, C = d.shape
    return d.reshape(B, Y // 2, 2, X // 2, 2, C).mean(axis=(2, 4))


def downsample2x_staggered(t: np.ndarray) -> np.ndarray:
    """StaggeredGrid.downsample2x on the packed tensor: keep every second face along the component's
    own axis, average the two faces across it."""
    vy, vx = unpack_staggered(t)
    vy_lo = 0.5 * (vy[:, ::2, 0::2] + vy[:, ::2, 1::2])
    vx_lo = 0.5 * (vx[:, 0::2, ::2] + vx[:, 1::2, ::2])
    return pack_staggered(vy_lo, vx_lo)


def downsample(d: np.ndarray, scale: int, staggered: bool) -> np.ndarray:
    assert scale in (1, 2, 4, 8)
    while scale > 1:
        d = downsample2x_staggered(d) if staggered else downsample2x_centered(d)
        scale //= 2
    return d
