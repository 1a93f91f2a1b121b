"""CPU check of the dense-matrix leaf's kernels (csrc/b2o_dense_kernels.cuh) under the host SIMT emulator (tests/emu/):
the same kernel bodies, split planning and launch logic the product compiles with nvcc, run thread-for-thread on the
CPU and compared with the oracle's `mul!(res, M, v, α, β)` (src/constructors.jl:25-27).  This pins the INDEX LOGIC of
every path (vectorised / scalar, split / unsplit, ragged rows and column groups, leading dimension > nrow, α/β, β = 0 with
NaN-filled res) where no GPU is available; the `-m gpu` tests repeat the comparison on the B200 through the C ABI."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
SO = os.path.join(EMU, "_build", "libdense_emu.so")
F64, F32 = 0, 1


@pytest.fixture(scope="module")
def emu():
    src = [os.path.join(EMU, f) for f in ("dense_emu.cpp", "simt_emu.h")]
    src.append(os.path.join(HERE, "..", "linearoperators.jl_b200", "csrc", "b2o_dense_kernels.cuh"))
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fvisibility=hidden",
                        "-Wl,-Bsymbolic", "-o", SO, src[0]],
                       check=True)
    L = ctypes.CDLL(SO)
    vp, i64, i32, d = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
    L.emu_dense_apply.restype = i32
    L.emu_dense_apply.argtypes = [i32, i32, i64, i64, i64, vp, vp, vp, d, d, i32, i32, ctypes.POINTER(i64)]
    L.emu_dense_plan.restype = None
    L.emu_dense_plan.argtypes = [i32, i32, i64, i64, i32, ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i32),
                                 ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.emu_last_error.restype = ctypes.c_char_p
    return L


def _aligned(n, dtype, offset_elems=0):
    """1-D array of n elements whose first element sits `offset_elems` elements past a 64-byte boundary"""
    item = np.dtype(dtype).itemsize
    raw = np.zeros((n + offset_elems) * item + 64, dtype=np.uint8)
    start = (-raw.ctypes.data) % 64
    return raw[start:start + (n + offset_elems) * item].view(dtype)[offset_elems:]


def _run(emu, orc, dtype, trans, m, n, lda=None, alpha=1.0, beta=0.0, moff=0, voff=0, num_sms=4, force_scalar=0, seed=0):
    dt = np.float64 if dtype == F64 else np.float32
    lda = max(1, m) if lda is None else lda
    rng = np.random.default_rng(seed)
    store = _aligned(lda * max(n, 1), dt, moff)
    store[:] = rng.uniform(-1, 1, store.shape[0]).astype(dt)
    M = store[:lda * n].reshape(n, lda).T[:m, :]                      # column-major m x n view, leading dimension lda
    nin, nout = (m, n) if trans else (n, m)
    v = _aligned(nin, dt, voff)
    v[:] = rng.uniform(-1, 1, nin).astype(dt)
    res = _aligned(nout, dt)
    res[:] = rng.uniform(-1, 1, nout).astype(dt) if beta != 0 else np.nan
    ref = res.copy()
    orc.gemv_(ref, np.array(M), v, alpha, beta, trans)
    launches = ctypes.c_int64()
    rc = emu.emu_dense_apply(dtype, trans, m, n, lda, store.ctypes.data, res.ctypes.data, v.ctypes.data, alpha, beta, num_sms,
                             force_scalar, ctypes.byref(launches))
    assert rc == 0, emu.emu_last_error()
    tol = 1e-13 if dtype == F64 else 2e-6
    scale = max(1.0, float(np.abs(ref).max())) if nout else 1.0
    assert nout == 0 or np.abs(res.astype(np.float64) - ref.astype(np.float64)).max() <= tol * scale * max(1, nin) ** 0.5
    return launches.value


SHAPES = [(5, 5), (10, 10), (20, 20),                 # the reference's GPU test blocks (test/gpu/nvidia.jl:8-10)
          (1, 1), (1, 9), (9, 1), (10, 6),            # test/test_linop.jl:2 (nrow, ncol) = (10, 6)
          (515, 70), (70, 601), (2051, 13), (3, 900), (1030, 24), (1283, 67)]


@pytest.mark.parametrize("dtype", [F64, F32])
@pytest.mark.parametrize("trans", [0, 1])
def test_dense_kernels_match_oracle(emu, orc, dtype, trans):
    for k, (m, n) in enumerate(SHAPES):
        _run(emu, orc, dtype, trans, m, n, seed=k)
        _run(emu, orc, dtype, trans, m, n, alpha=2.0, beta=-0.5, seed=100 + k)


@pytest.mark.parametrize("dtype", [F64, F32])
@pytest.mark.parametrize("trans", [0, 1])
def test_dense_kernels_layout_variants(emu, orc, dtype, trans):
    W = 2 if dtype == F64 else 4
    m, n = 1030, 75
    _run(emu, orc, dtype, trans, m, n, lda=m + 2 * W, seed=1)                 # sub-matrix view, still vectorisable
    _run(emu, orc, dtype, trans, m, n, lda=m + 1, alpha=-1.5, beta=2.0, seed=2)   # odd leading dimension -> scalar kernels
    _run(emu, orc, dtype, trans, m, n, moff=1, seed=3)                        # matrix base not 16-byte aligned -> scalar kernels
    _run(emu, orc, dtype, trans, m, n, voff=1, seed=4)                        # input vector not 16-byte aligned (T: scalar)
    _run(emu, orc, dtype, trans, m, n, force_scalar=1, alpha=0.5, beta=1.0, seed=5)
    _run(emu, orc, dtype, trans, 4 * W * 256 + W - 1, 9, seed=6)              # ragged last rows of a vectorised launch
    _run(emu, orc, dtype, trans, 300, 8 * 5, seed=7)                          # only full 8-column groups
    _run(emu, orc, dtype, trans, 300, 8 * 5 + 3, seed=8)                      # + one partial column group
    # narrow kernels (few rows): ragged last row quad, sub-matrix views, scalar variants, many CTA iterations
    # (every emulated shuffle is two barriers of 32 OS threads, so the T cases are kept smaller)
    shapes = ((7, 700), (33, 1300), (2 * W * 16 - 1, 517), (W * 32, 64), (W * 128 - 1, 130), (1, 3000))
    for k, (mm, nn) in enumerate(shapes):
        nn = nn if not trans else max(40, nn // 6)
        _run(emu, orc, dtype, trans, mm, nn, num_sms=1, seed=20 + k)
        _run(emu, orc, dtype, trans, mm, nn, lda=mm + 3 * W, alpha=1.5, beta=-2.0, num_sms=2, seed=30 + k)
        _run(emu, orc, dtype, trans, mm, nn, lda=mm + 1, moff=1, num_sms=1, seed=40 + k)
        if k % 2 == 0:
            _run(emu, orc, dtype, trans, mm, nn, force_scalar=1, num_sms=1, seed=50 + k)


@pytest.mark.parametrize("dtype", [F64, F32])
def test_dense_kernels_split_products(emu, orc, dtype):
    """short-and-wide (N) and tall-and-skinny (T) matrices are split across grid.y; the partial sums are added in split
    order by the finish kernel (2 launches) and the result is the same as the unsplit one to rounding."""
    assert _run(emu, orc, dtype, 0, 40, 3000, num_sms=4, alpha=2.0, beta=0.25, seed=1) == 2
    assert _run(emu, orc, dtype, 1, 20000, 5, num_sms=4, alpha=2.0, beta=0.25, seed=2) == 2
    assert _run(emu, orc, dtype, 0, 40, 3000, num_sms=4, seed=3) == 2                  # β = 0: res (NaN) never read
    assert _run(emu, orc, dtype, 1, 20000, 5, num_sms=4, seed=4) == 2
    assert _run(emu, orc, dtype, 0, 3000, 40, num_sms=1, seed=5) == 1                  # enough row blocks: one launch
    assert _run(emu, orc, dtype, 1, 40, 300, num_sms=1, seed=6) == 1


@pytest.mark.parametrize("dtype", [F64, F32])
def test_dense_kernels_empty_shapes(emu, orc, dtype):
    for trans in (0, 1):
        for m, n in ((0, 5), (5, 0), (0, 0)):
            _run(emu, orc, dtype, trans, m, n)                                          # zero-length products: res = 0 (β = 0)
            _run(emu, orc, dtype, trans, m, n, alpha=3.0, beta=2.0, seed=9)             # ... or β res


def test_dense_plan_invariants(emu):
    """the split plan covers every column/row exactly once, keeps T-splits on 16-byte boundaries, fits the CUDA grid limits,
    and matrices with few rows go to the narrow kernels with enough row threads"""
    gx, chunk, ns = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
    narrow, txl = ctypes.c_int(), ctypes.c_int()
    dims = [0, 1, 2, 7, 8, 9, 63, 64, 65, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1000, 4096, 10**5, 10**6, 10**8]
    seen = set()
    for sms in (1, 4, 148):
        for W in (1, 2, 4):
            for trans in (0, 1):
                for m in dims:
                    for n in dims:
                        if m * n > 10**13:
                            continue
                        emu.emu_dense_plan(sms, trans, m, n, W, ctypes.byref(gx), ctypes.byref(chunk), ctypes.byref(ns),
                                           ctypes.byref(narrow), ctypes.byref(txl))
                        key = (sms, W, trans, m, n)
                        assert ns.value >= 1 and chunk.value >= 1 and gx.value >= 1, key
                        assert ns.value <= 1024 and gx.value < 2**31, key
                        seen.add((trans, narrow.value))
                        if narrow.value:
                            TX = 1 << txl.value
                            assert TX * W >= m and TX <= (32 if trans else 128), key      # every row has a thread
                            assert (m + W - 1) // W <= (32 if trans else 128), key
                            assert gx.value * chunk.value >= n, key                      # the CTAs cover the columns
                            assert (gx.value - 1) * chunk.value < max(n, 1), key         # no empty CTA
                            assert chunk.value % ((256 // TX) * 8) == 0, key             # whole CTA iterations
                            assert ns.value == (1 if trans else gx.value), key
                            continue
                        length = m if trans else n                     # the dimension that is split
                        assert ns.value * chunk.value >= length, key
                        assert (ns.value - 1) * chunk.value < max(length, 1), key   # no empty split
                        if trans:
                            assert chunk.value % (256 * W) == 0, key
                            assert gx.value * 8 >= n, key
                        else:
                            assert gx.value * 256 * W >= m, key
    assert seen == {(0, 0), (0, 1), (1, 0), (1, 1)}
