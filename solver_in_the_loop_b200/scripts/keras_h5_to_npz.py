"""Convert the reference's Keras checkpoints (`model.h5`, karman_train.py:514-517 / burgers_train.py:497-500) to the Keras-ordered
`.npz` this package loads (`CorrectionModel.load`, `--model` of the apply scripts, `--inittf` of the train scripts), and back.

    python -m solver_in_the_loop_b200.scripts.keras_h5_to_npz model.h5 model.npz
    python -m solver_in_the_loop_b200.scripts.keras_h5_to_npz --to-h5 model.npz template.h5 out.h5

Needs `h5py`, which is NOT available in the offline build container — run it where the reference's own environment
is installed.  The `.npz` holds `arr_0, arr_1, ...` = kernel `[5,5,Cin,Cout]`, bias `[Cout]` per Conv2D layer in model order, i.e.
exactly `model.get_weights()`.
"""
import sys

import numpy as np


def _h5():
    try:
        import h5py
        return h5py
    except ImportError:
        raise SystemExit("h5py is not installed here; run this converter in the reference's environment (pip install h5py)")


def _weight_datasets(f):
    """Weight datasets of a Keras HDF5 file in model order (attrs layer_names / weight_names)."""
    g = f["model_weights"] if "model_weights" in f else f
    out = []
    for ln in g.attrs["layer_names"]:
        ln = ln.decode() if isinstance(ln, bytes) else ln
        for wn in g[ln].attrs["weight_names"]:
            wn = wn.decode() if isinstance(wn, bytes) else wn
            out.append(g[ln][wn])
    return out


def h5_to_npz(src, dst):
    h5py = _h5()
    with h5py.File(src, "r") as f:
        ws = [np.asarray(d) for d in _weight_datasets(f)]
    np.savez(dst, *ws)
    print("wrote %s: %d arrays, %d parameters" % (dst, len(ws), sum(w.size for w in ws)))


def npz_to_h5(src, template, dst):
    """Write the arrays of `src` into a copy of `template` (an .h5 saved by the reference for the same architecture)."""
    import shutil
    h5py = _h5()
    z = np.load(src)
    ws = [z["arr_%d" % i] for i in range(len(z.files))]
    shutil.copyfile(template, dst)
    with h5py.File(dst, "r+") as f:
        ds = _weight_datasets(f)
        if len(ds) != len(ws):
            raise SystemExit("architecture mismatch: %d arrays vs %d weight datasets" % (len(ws), len(ds)))
        for d, w in zip(ds, ws):
            if tuple(d.shape) != tuple(w.shape):
                raise SystemExit("shape mismatch: %s vs %s" % (d.shape, w.shape))
            d[...] = w
    print("wrote", dst)


def main(argv=None):
    a = sys.argv[1:] if argv is None else argv
    if len(a) == 2:
        return h5_to_npz(a[0], a[1])
    if len(a) == 4 and a[0] == "--to-h5":
        return npz_to_h5(a[1], a[2], a[3])
    raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
