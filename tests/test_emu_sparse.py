"""CPU check of the sparse-matrix leaf's kernels (csrc/b2o_sparse_kernels.cuh) under the host SIMT emulator (tests/emu/):
the product's row kernel, lane-group sizing and launch logic run thread-for-thread on the CPU and are compared with the
oracle's restatement of SparseArrays' `mul!(res, M, v, α, β)` (the closures of LinearOperator(M::SparseMatrixCSC),
src/constructors.jl:25-27) and with scipy.sparse as an independent implementation.  The `-m gpu` tests repeat the
comparison on the B200 through the C ABI."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
SO = os.path.join(EMU, "_build", "libsparse_emu.so")
F64, F32 = 0, 1


@pytest.fixture(scope="module")
def emu():
    src = [os.path.join(EMU, f) for f in ("sparse_emu.cpp", "simt_emu.h")]
    src.append(os.path.join(HERE, "..", "linearoperators.jl_b200", "csrc", "b2o_sparse_kernels.cuh"))
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in src):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fvisibility=hidden",
                        "-Wl,-Bsymbolic", "-o", SO, src[0]], check=True)
    L = ctypes.CDLL(SO)
    vp, i64, i32, d = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_double
    L.emu_sparse_apply.restype = i32
    L.emu_sparse_apply.argtypes = [i32, i32, i64, i64, i64, vp, vp, vp, i32, vp, vp, d, d, i32, ctypes.POINTER(i64),
                                   ctypes.POINTER(i32)]
    return L


def random_sparse(m, n, density, seed, dtype, dense_row=None, empty_rows=()):
    rng = np.random.default_rng(seed)
    A = sp.random(m, n, density=density, random_state=rng, format="lil", dtype=np.float64)
    A = A.tolil()
    if dense_row is not None and m > 0:
        A[dense_row, :] = rng.uniform(-1, 1, n)
    for r in empty_rows:
        if r < m:
            A[r, :] = 0
    A = A.tocsc().astype(dtype)
    A.eliminate_zeros()
    A.sort_indices()
    return A


@pytest.mark.parametrize("dtype", [F64, F32])
@pytest.mark.parametrize("fmt", [0, 1])
def test_spmv_kernel_matches_oracle(emu, orc, dtype, fmt):
    dt = np.float64 if dtype == F64 else np.float32
    seen = set()
    cases = [(10, 6, 0.5), (6, 10, 0.5), (200, 300, 0.004), (300, 200, 0.02), (150, 150, 0.05), (64, 500, 0.12),
             (40, 700, 0.3), (33, 900, 0.9), (1, 50, 0.5), (50, 1, 0.5), (257, 129, 0.03)]
    for k, (m, n, dens) in enumerate(cases):
        A = random_sparse(m, n, dens, k, dt, dense_row=(m // 2 if k % 3 == 0 else None), empty_rows=(0, m - 1) if k % 2 else ())
        for trans in (0, 1):
            seen.add(run_checked(emu, orc, A, dtype, fmt, trans, 1.0, 0.0, seed=k))
            run_checked(emu, orc, A, dtype, fmt, trans, 2.0, -0.5, seed=50 + k)
    assert {0, 1, 2, 3, 4, 5} <= seen          # every lane-group width 1..32 was exercised


def run_checked(emu, orc, A, dtype, fmt, trans, alpha, beta, num_sms=2, seed=0):
    """run() + independent scipy check, with the starting res regenerated for β != 0"""
    dt = np.float64 if dtype == F64 else np.float32
    m, n = A.shape
    csc = A.tocsc()
    csc.sort_indices()
    S = csc if fmt == 0 else A.tocsr()
    S.sort_indices()
    ptr1 = np.ascontiguousarray(S.indptr.astype(np.int64) + 1)
    idx1 = np.ascontiguousarray(S.indices.astype(np.int64) + 1)
    vals = np.ascontiguousarray(S.data.astype(dt))
    rng = np.random.default_rng(1000 + seed)
    nin, nout = (m, n) if trans else (n, m)
    v = rng.uniform(-1, 1, nin).astype(dt)
    res0 = rng.uniform(-1, 1, nout).astype(dt)
    res = res0.copy() if beta != 0 else np.full(nout, np.nan, dtype=dt)
    ref = res.copy()
    orc.spmv_csc_(ref, m, n, csc.indptr.astype(np.int64) + 1, csc.indices.astype(np.int64) + 1, csc.data.astype(dt), v, alpha, beta,
                  trans)
    launches, lanes = ctypes.c_int64(), ctypes.c_int()
    rc = emu.emu_sparse_apply(dtype, fmt, m, n, S.nnz, ptr1.ctypes.data, idx1.ctypes.data, vals.ctypes.data, trans,
                              res.ctypes.data, v.ctypes.data, alpha, beta, num_sms, ctypes.byref(launches), ctypes.byref(lanes))
    assert rc == 0
    tol = 1e-13 if dtype == F64 else 2e-6
    if nout:
        B = (A.T if trans else A).astype(np.float64)
        rowsum = np.asarray(abs(B).sum(axis=1)).ravel()
        bound = tol * (abs(alpha) * np.maximum(rowsum, 1.0) + abs(beta) + 1.0)
        assert np.all(np.abs(res.astype(np.float64) - ref.astype(np.float64)) <= bound), (A.shape, fmt, trans)
        ind = alpha * (B @ v.astype(np.float64)) + (beta * res0.astype(np.float64) if beta != 0 else 0.0)
        assert np.all(np.abs(ref.astype(np.float64) - ind) <= 10 * bound)               # scipy, independent of both
    return lanes.value


@pytest.mark.parametrize("dtype", [F64, F32])
def test_spmv_kernel_edge_cases(emu, orc, dtype):
    dt = np.float64 if dtype == F64 else np.float32
    for fmt in (0, 1):
        for trans in (0, 1):
            for shape in ((0, 5), (5, 0), (0, 0), (7, 7)):
                Z = sp.csc_matrix(shape, dtype=dt)                       # nnz == 0: res = 0 (β = 0) or β res
                run_checked(emu, orc, Z, dtype, fmt, trans, 1.0, 0.0)
                run_checked(emu, orc, Z, dtype, fmt, trans, 3.0, 2.0, seed=3)
            D = sp.identity(300, dtype=dt, format="csc") * 2.5           # diagonal: one entry per row -> 1 lane per row
            assert run_checked(emu, orc, D, dtype, fmt, trans, 1.0, 0.0, num_sms=1) == 0
            R = sp.csc_matrix(np.ones((3, 4000), dtype=dt))              # three dense rows / 4000 tiny columns
            run_checked(emu, orc, R, dtype, fmt, trans, 1.0, 0.0, num_sms=1)
