"""The C-ABI shared library loads and exports every symbol include/sol_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sol_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sol_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    import __graft_entry__ as ge
    ge.build()
    from solver_in_the_loop_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header drifted apart"
    lib.sol_abi_version.restype = ctypes.c_int
    assert lib.sol_abi_version() == 1


def test_param_count_matches_reference_model():
    from solver_in_the_loop_b200 import _lib
    lib = _lib.load()
    assert lib.sol_model_param_count(_lib.SOL_MODEL_MARS_MOON, 3) == 260354
    assert lib.sol_model_param_count(_lib.SOL_MODEL_MARS_MOON, 4) == 261154     # burgers


def test_engine_refuses_cpu_tensors():
    import pytest
    import torch
    from solver_in_the_loop_b200 import engine
    with pytest.raises(engine.SolError):
        engine._ptr(torch.zeros(4))


def test_header_is_plain_c(tmp_path):
    """include/sol_b200.h is the drop-in boundary: it must compile as C99 (no C++ / CUDA / torch types in the signatures) and a C
    translation unit that references every declared entry point must link against the shared library."""
    import subprocess
    from solver_in_the_loop_b200 import _lib
    names = _declared()
    src = tmp_path / "abi_check.c"
    body = "\n".join("    p[%d] = (fn)%s;" % (i, n) for i, n in enumerate(names))
    src.write_text('#include "sol_b200.h"\n#include <stdio.h>\ntypedef void (*fn)(void);\nint main(void) {\n    fn p[%d];\n%s\n    printf("%%d %%d\\n", sol_abi_version(), p[0] != 0);\n    return 0;\n}\n'
                   % (len(names), body))
    exe = tmp_path / "abi_check"
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libsol_b200.so", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split()[0] == "1", out.stderr
