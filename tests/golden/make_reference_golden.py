"""Generates REFERENCE golden vectors — run this wherever `phiflow==1.5.1` (commit 4f5e678, numpy backend is enough) is importable:

    python tests/golden/make_reference_golden.py /path/to/Solver-in-the-Loop

It executes the reference's OWN `KarmanFlow` (the class body is read out of `karman-2d/karman_apply.py` at run time with `ast`,
nothing is copied into this repository) and its NumPy-backend `BurgersTest` for a few steps from the reference's warm start, and writes
`tests/golden/phiflow_karman_64x32.npz` / `phiflow_burgers_32x32.npz`.  `tests/test_reference_golden.py` then checks the oracle
(and, on a GPU box, the CUDA path) against them — which turns "parity unpinned" into a pinned statement.  In the build container
PhiFlow / TensorFlow are not installable (no network), so the fixtures are absent and those tests skip.
"""
import ast
import os
import sys

import numpy as np


def _class_source(path, name):
    src = open(path).read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise SystemExit("class %s not found in %s" % (name, path))


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    ref = sys.argv[1]
    try:
        import phi.flow as pf                                    # noqa: F401  (PhiFlow 1.5.1, numpy backend)
    except ImportError:
        raise SystemExit("phiflow is not importable here: pip install phiflow==1.5.1 (see the reference README.md:21-24)")
    here = os.path.dirname(os.path.abspath(__file__))
    ns = {}
    exec("from phi.flow import *", ns)
    # ---- karman: res 32 (64x32), the reference's warm start (karman_apply.py:96-101), Re = 1e6, 3 steps
    exec(_class_source(os.path.join(ref, "karman-2d", "karman_apply.py"), "KarmanFlow"), ns)
    res, L, Re, steps = 32, 100, 1.0e6, 3
    st = ns["Fluid"](ns["Domain"](resolution=[res * 2, res], box=ns["box"][0:L * 2, 0:L], boundaries=ns["OPEN"]), buoyancy_factor=0)
    vn = st.velocity.staggered_tensor()
    vn[..., 0] = 1.0
    vn[..., vn.shape[1] // 2 + 10:vn.shape[1] // 2 + 20, vn.shape[2] // 2 - 2:vn.shape[2] // 2 + 2, 1] = 1.0
    st = st.copied_with(velocity=ns["StaggeredGrid"](ns["unstack_staggered_tensor"](vn), st.velocity.box))
    bc = np.zeros(st.velocity.data[0].data.shape)
    bc[..., 0:2, 0:bc.shape[2] - 1, 0] = 1.0
    bc[..., 0:bc.shape[1], 0:1, 0] = 1.0
    bc[..., 0:bc.shape[1], -1:, 0] = 1.0
    sim = ns["KarmanFlow"]()
    out = {"res": res, "L": L, "re": Re, "v0": st.velocity.staggered_tensor(), "d0": st.density.data}
    for i in range(steps):
        st = sim.step(st, re=Re, res=res, velBCy=bc, velBCyMask=np.copy(bc))
        out["v%d" % (i + 1)] = st.velocity.staggered_tensor()
        out["d%d" % (i + 1)] = st.density.data
    np.savez_compressed(os.path.join(here, "phiflow_karman_64x32.npz"), **out)
    # ---- burgers: 32x32 periodic, dt 0.1, a smooth deterministic start, 3 unforced steps
    exec(_class_source(os.path.join(ref, "burgers", "burgers_apply.py"), "BurgersVelocitySMAC"), ns)
    exec(_class_source(os.path.join(ref, "burgers", "burgers_apply.py"), "BurgersTest"), ns)
    R, Lb, dt = 32, 32, 0.1
    dm = ns["Domain"]([R, R], box=ns["box"]([Lb] * 2), boundaries=ns["PERIODIC"])
    jj, ii = np.meshgrid(np.arange(R + 1), np.arange(R + 1), indexing="ij")
    v0 = np.stack([np.sin(2 * np.pi * jj / R) * np.cos(2 * np.pi * ii / R), 0.5 * np.cos(4 * np.pi * jj / R)], axis=-1)[None].astype(np.float32)
    sb = ns["BurgersVelocitySMAC"](dm, velocity=v0)
    bsim = ns["BurgersTest"]()
    outb = {"R": R, "L": Lb, "dt": dt, "v0": sb.velocity.staggered_tensor()}
    for i in range(3):
        sb = bsim.step(v=sb, dt=dt)
        outb["v%d" % (i + 1)] = sb.velocity.staggered_tensor()
    np.savez_compressed(os.path.join(here, "phiflow_burgers_32x32.npz"), **outb)
    print("wrote phiflow_karman_64x32.npz and phiflow_burgers_32x32.npz under", here)


if __name__ == "__main__":
    main()
