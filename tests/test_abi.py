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


def test_fused_graph_lowering_and_nvrtc_on_cpu(lo):
    """A dry graph (no context) can be lowered and its specialised kernel NVRTC-compiled for sm_100a without a GPU."""
    from linearoperators_jl_b200 import _lib
    lib = _lib.load()
    g, node = ctypes.c_void_p(), ctypes.c_int()
    _lib.check(lib.b2o_graph_create(None, 12345, ctypes.byref(g)))

    def leaf(kind, ptr=None):
        _lib.check(lib.b2o_graph_leaf(g, kind, ptr, ctypes.byref(node)))
        return node.value

    H, D, E = leaf(4, ctypes.c_void_p(0x1000)), leaf(0, ctypes.c_void_p(0x2000)), leaf(1)
    _lib.check(lib.b2o_graph_binary(g, 11, H, D, ctypes.byref(node)))
    P = node.value
    _lib.check(lib.b2o_graph_unary(g, 12, E, 0.1, ctypes.byref(node)))
    S = node.value
    _lib.check(lib.b2o_graph_binary(g, 10, P, S, ctypes.byref(node)))
    _lib.check(lib.b2o_graph_compile(g, node.value))
    npass, nred, nbytes = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
    _lib.check(lib.b2o_graph_info(g, 0, 0.0, ctypes.byref(npass), ctypes.byref(nred), ctypes.byref(nbytes)))
    assert (npass.value, nred.value, nbytes.value) == (2, 1, 7 * 8 * 12345)          # cfg3: 2 passes, 7n*8 bytes
    _lib.check(lib.b2o_graph_info(g, 0, 0.5, ctypes.byref(npass), ctypes.byref(nred), ctypes.byref(nbytes)))
    assert (npass.value, nred.value, nbytes.value) == (2, 1, 8 * 8 * 12345)          # + res read when beta != 0
    _lib.check(lib.b2o_graph_info(g, 1, 0.0, ctypes.byref(npass), ctypes.byref(nred), ctypes.byref(nbytes)))
    assert (npass.value, nred.value, nbytes.value) == (2, 1, 6 * 8 * 12345)          # transpose = D*H: the dot reads h, v only
    buf, n = ctypes.create_string_buffer(1 << 16), ctypes.c_int64()
    _lib.check(lib.b2o_graph_jit_source(g, 0, 0.0, buf, 1 << 16, ctypes.byref(n)))
    src = buf.value.decode()
    assert "__dmul_rn(a0, __dmul_rn(a1, a2))" in src and "b2o_fused" in src          # dot(h, d .* v), no fma contraction
    cb = ctypes.c_int64()
    st = lib.b2o_graph_jit_check(g, ctypes.byref(cb))
    if st == _lib.B2O_EUNSUPPORTED:
        pytest.skip("NVRTC not installed")
    assert st == 0, lib.b2o_last_error()
    assert cb.value > 1000
    res = ctypes.c_void_p(0x3000)
    assert lib.b2o_graph_apply(g, 0, res, 12345, res, 12345, 1.0, 0.0) == _lib.B2O_EARG   # dry graphs cannot be applied
    lib.b2o_graph_destroy(g)


def test_julia_shim_binds_every_export_with_the_header_arity():
    """julia/B200LinearOperators.jl cannot run here (no Julia anywhere): at least every function include/b2o.h declares must
    be bound there, and every `ccall` must list as many argument types as the C prototype has parameters."""
    from linearoperators_jl_b200 import _lib
    decl = {n: len(a) for n, _, a in _lib.declared_functions()}
    src = open(os.path.join(ROOT, "julia", "B200LinearOperators.jl")).read()
    seen = {}
    for m in re.finditer(r"ccall\(\(:(b2o_\w+), libb2o\),\s*(\w+),\s*\(", src):
        name, i, depth = m.group(1), m.end(), 1
        j = i
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        types = src[i:j - 1]
        # count top-level commas (types like Ptr{Ptr{Cvoid}} contain none outside braces)
        level, count, has = 0, 0, bool(types.strip().strip(","))
        for ch in types:
            level += {"{": 1, "}": -1}.get(ch, 0)
            count += ch == "," and level == 0
        nargs = 0 if not has else count + (0 if types.strip().endswith(",") else 1)
        seen.setdefault(name, set()).add(nargs)
    missing = sorted(set(decl) - set(seen) - {"b2o_last_error"})
    assert not missing, missing
    wrong = {n: (sorted(a), decl[n]) for n, a in seen.items() if n in decl and a != {decl[n]}}
    assert not wrong, wrong


def test_fused_graph_aot_table_is_current_and_hits(lo):
    """csrc/b2o_graph_aot.cu (generated) must match what the library's code generator produces today, and BASELINE config 3
    must resolve to an ahead-of-time kernel key (so it needs no libnvrtc at run time)"""
    r = subprocess.run([os.sys.executable, os.path.join(ROOT, "tools", "gen_graph_aot.py"), "--check"], capture_output=True, text=True,
                       cwd=ROOT)
    assert r.returncode == 0, r.stdout + r.stderr
    from linearoperators_jl_b200 import _lib
    lib = _lib.load()
    g, node = ctypes.c_void_p(), ctypes.c_int()
    _lib.check(lib.b2o_graph_create(None, 777, ctypes.byref(g)))

    def leaf(kind, ptr=None):
        _lib.check(lib.b2o_graph_leaf(g, kind, ptr, ctypes.byref(node)))
        return node.value

    H, D, E = leaf(4, ctypes.c_void_p(0x5000)), leaf(0, ctypes.c_void_p(0x7000)), leaf(1)
    _lib.check(lib.b2o_graph_binary(g, 11, H, D, ctypes.byref(node)))
    P = node.value
    _lib.check(lib.b2o_graph_unary(g, 12, E, 0.25, ctypes.byref(node)))          # a different scalar: same source, same key
    S = node.value
    _lib.check(lib.b2o_graph_binary(g, 10, P, S, ctypes.byref(node)))
    _lib.check(lib.b2o_graph_compile(g, node.value))
    table = open(os.path.join(ROOT, "linearoperators.jl_b200", "csrc", "b2o_graph_aot.cu")).read()
    for tr in (0, 1):
        for beta in (0.0, 2.0):
            h = ctypes.c_uint64()
            _lib.check(lib.b2o_graph_variant(g, tr, beta, None, ctypes.byref(h)))
            assert "0x%016xULL" % h.value in table, (tr, beta)
    lib.b2o_graph_destroy(g)
