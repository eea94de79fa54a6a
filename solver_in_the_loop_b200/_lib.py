"""ctypes binding of libsol_b200.so (the C ABI declared in include/sol_b200.h).

The product path has NO CPU fallback: if the shared library is missing or the machine has no
CUDA device, importing the engine raises.  ``python -m solver_in_the_loop_b200.csrc.build`` (or
``__graft_entry__.build()``) compiles the library in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsol_b200.so")

SOL_OK = 0
SOL_BOUNDARY_OPEN, SOL_BOUNDARY_PERIODIC = 0, 1
SOL_MODEL_MARS_MOON, SOL_MODEL_MERCURY = 0, 1
SOL_ACT_NONE, SOL_ACT_LRELU, SOL_ACT_DLRELU = 0, 1, 2


class SolError(RuntimeError):
    pass


class UnrollCfg(C.Structure):
    _fields_ = [("model", C.c_int), ("cin0", C.c_int), ("msteps", C.c_int), ("B", C.c_int),
                ("dt", C.c_float), ("res", C.c_float),
                ("sig_vy", C.c_float), ("sig_vx", C.c_float), ("sig_ext", C.c_float),
                ("with_density", C.c_int), ("use_graph", C.c_int)]


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/sol_b200.h declares
SIGNATURES = {
    "sol_abi_version": (_i, []),
    "sol_last_error_string": (C.c_char_p, []),
    "sol_launch_count": (C.c_ulonglong, []),
    "sol_plan_create": (_i, [_i, _i, _i, _f, _i, _vp, _vp, _vp, _vp, C.POINTER(_vp)]),
    "sol_plan_destroy": (_i, [_vp]),
    "sol_plan_set_cg": (_i, [_vp, _f, _f, _i, _i]),
    "sol_plan_set_option": (_i, [_vp, C.c_char_p, _i]),
    "sol_plan_query": (_i, [_vp, C.c_char_p, C.POINTER(_i)]),
    "sol_set_option": (_i, [C.c_char_p, _i]),
    "sol_diffuse_bc": (_i, [_vp, _vp, _i, _vp, _f, _f, _vp, _vp, _vp, _vp]),
    "sol_diffuse_bc_bwd": (_i, [_vp, _vp, _i, _vp, _f, _f, _vp, _vp, _vp, _vp]),
    "sol_diffuse_advect": (_i, [_vp, _vp, _i, _vp, _f, _f] + [_vp] * 8),
    "sol_advect": (_i, [_vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sol_advect_bwd": (_i, [_vp, _vp, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sol_pressure_solve": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "sol_project": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sol_divergence": (_i, [_vp, _vp, _i, _vp, _vp, _vp]),
    "sol_step_fwd": (_i, [_vp, _vp, _i, _vp, _f, _f] + [_vp] * 12),
    "sol_step_bwd": (_i, [_vp, _vp, _i, _vp, _f, _f] + [_vp] * 9),
    "sol_burgers_step": (_i, [_vp, _vp, _i, _f, _f] + [_vp] * 10),
    "sol_burgers_step_bwd": (_i, [_vp, _vp, _i, _f, _f] + [_vp] * 10),
    "sol_conv5x5": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp]),
    "sol_conv5x5_split_floats": (_sz, []),
    "sol_conv5x5_split_weights": (_i, [_vp, _vp, _vp]),
    "sol_conv5x5_c32_presplit": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _f, _vp, _i]),
    "sol_conv5x5_flip_weights": (_i, [_vp, _i, _i, _vp, _vp]),
    "sol_conv5x5_wgrad_workspace": (_sz, [_i, _i]),
    "sol_conv5x5_wgrad": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "sol_model_param_count": (_sz, [_i, _i]),
    "sol_to_feature": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _f, _f, _f, _vp]),
    "sol_unroll_workspace_bytes": (_sz, [_vp, C.POINTER(UnrollCfg)]),
    "sol_unroll_create": (_i, [_vp, C.POINTER(UnrollCfg), _vp, _sz, C.POINTER(_vp)]),
    "sol_unroll_destroy": (_i, [_vp]),
    "sol_unroll_forward": (_i, [_vp] * 13),
    "sol_unroll_rollout": (_i, [_vp] * 7 + [_i] + [_vp] * 3),
    "sol_unroll_backward": (_i, [_vp] * 6),
    "sol_unroll_train_iter": (_i, [_vp] * 11),
    "sol_unroll_set_burgers": (_i, [_vp, _f, _vp, _vp, _vp, _vp, _f, _f]),
    "sol_unroll_cg_iters": (_i, [_vp, C.POINTER(_vp), C.POINTER(_i)]),
    "sol_adam_tf1": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _i, _f, _f, _f, _f, _f]),
}

_lib = None


def load():
    """Load the shared library and bind every declared symbol; raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SolError(
            "libsol_b200.so not found at %s — build it with `python -m solver_in_the_loop_b200.csrc.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here = ABI drift between header and library
        fn.restype = res
        fn.argtypes = args
    if lib.sol_abi_version() != 1:
        raise SolError("libsol_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc: int):
    if rc != SOL_OK:
        msg = load().sol_last_error_string()
        raise SolError("libsol_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))
