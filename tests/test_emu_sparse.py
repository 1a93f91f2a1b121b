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
                                   ctypes.POINTER(i32), i32, ctypes.POINTER(i64), i32]
    L.emu_sparse_tiles.restype = i64
    L.emu_sparse_tiles.argtypes = [vp, i64, i64, vp, i64, ctypes.POINTER(i32), ctypes.POINTER(i32)]
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


def run_checked(emu, orc, A, dtype, fmt, trans, alpha, beta, num_sms=2, seed=0, kernel=1, want_tiles=None, force_lanes=-1, out=None):
    """one product through the emulated row kernel (kernel=1) or TMA-staged tile kernel (kernel=2), checked against the
    oracle and, independently, scipy; the starting res is regenerated for β != 0"""
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
    launches, lanes, ntiles = ctypes.c_int64(), ctypes.c_int(), ctypes.c_int64()
    rc = emu.emu_sparse_apply(dtype, fmt, m, n, S.nnz, ptr1.ctypes.data, idx1.ctypes.data, vals.ctypes.data, trans,
                              res.ctypes.data, v.ctypes.data, alpha, beta, num_sms, ctypes.byref(launches), ctypes.byref(lanes),
                              kernel, ctypes.byref(ntiles), force_lanes)
    assert rc == 0
    if out is not None:
        out.append(res.copy())
    if want_tiles is not None:
        assert ntiles.value >= want_tiles, (ntiles.value, want_tiles)
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


# ---------------------------------------------------------------- the TMA-staged tile kernel (spmv_tiles_kernel)
def tile_cut(emu, ptr0, nnz):
    ptr0 = np.ascontiguousarray(ptr0, dtype=np.int64)
    nrows = len(ptr0) - 1
    out = np.zeros(3 * (nrows + 2), dtype=np.int64)
    c, rt = ctypes.c_int(), ctypes.c_int()
    nt = emu.emu_sparse_tiles(ptr0.ctypes.data, nrows, nnz, out.ctypes.data, nrows + 2, ctypes.byref(c), ctypes.byref(rt))
    return out[:3 * (nt + 1)].reshape(nt + 1, 3), c.value, rt.value


def test_tile_cut_invariants(emu):
    """spmv_build_tiles: tiles cover the rows once and in order; staged ranges start on a quad at or before the first entry,
    cover the tile's entries, fit the stage; rows that cannot fit become direct tiles of one row; sentinel at the end"""
    rng = np.random.default_rng(0)
    for trial in range(30):
        nrows = int(rng.integers(1, 4000))
        kind = trial % 5
        if kind == 0:
            lens = rng.integers(0, 6, nrows)
        elif kind == 1:
            lens = np.where(rng.uniform(size=nrows) < 0.01, rng.integers(2000, 9000, nrows), rng.integers(0, 40, nrows))
        elif kind == 2:
            lens = np.zeros(nrows, dtype=np.int64)
            lens[rng.integers(0, nrows, 3)] = rng.integers(1, 5000, 3)
        elif kind == 3:
            lens = rng.integers(500, 1100, nrows % 50 + 1)
        else:
            lens = np.full(nrows, 1)
        ptr0 = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        nrows, nnz = len(lens), int(ptr0[-1])
        tiles, C, RT = tile_cut(emu, ptr0, nnz)
        assert tuple(tiles[-1]) == (nnz, nrows, 0) and tiles[0, 1] == 0
        for (e0, r0, ne), (_, r1, _) in zip(tiles[:-1], tiles[1:]):
            assert r1 > r0 and r1 - r0 <= RT
            if ne < 0:
                assert r1 == r0 + 1 and e0 == ptr0[r0] and ptr0[r1] - (ptr0[r0] & ~3) > C      # a row that cannot fit
            else:
                assert e0 % 4 == 0 and ne % 4 == 0 and 0 <= ptr0[r0] - e0 <= 3
                assert e0 + ne >= ptr0[r1] and ne <= C and e0 + ne - ptr0[r1] <= 3
                # greedy: the next row would not have fitted (or the row limit / end was hit)
                assert r1 == nrows or r1 - r0 == RT or ptr0[r1 + 1] - e0 > C


@pytest.mark.parametrize("dtype", [F64, F32])
def test_tile_kernel_matches_oracle(emu, orc, dtype):
    """the tile kernel on structures that exercise: several tiles per CTA and more CTAs than tiles, nnz not a multiple of 4
    (entries past the last whole quad come from global memory), odd first rows (offset slice starts one row early), long rows
    (direct tiles), long runs of empty rows (row limit per tile), every lane-group width"""
    dt = np.float64 if dtype == F64 else np.float32
    rng = np.random.default_rng(7)
    seen = set()
    cases = []
    A = random_sparse(700, 900, 0.02, 1, dt)                                   # ~18 entries per row, ~6 tiles
    cases.append((A, 3))
    B = random_sparse(300, 5000, 0.004, 2, dt, dense_row=150)                  # one 5000-entry row -> direct tile
    cases.append((B, 3))
    Cm = sp.identity(5000, dtype=dt, format="csc") * 1.5                       # 1 entry per row -> row limit binds (5 tiles)
    cases.append((Cm, 4))
    Dm = sp.lil_matrix((3000, 40), dtype=dt)                                   # long runs of empty rows, entries in rows 1000-1020
    Dm[1000:1021, :] = rng.uniform(-1, 1, (21, 40))
    cases.append((Dm.tocsc(), 2))
    Em = random_sparse(37, 2500, 0.5, 3, dt)                                   # ~1250 per row: one row per tile, 32 lanes
    cases.append((Em, 30))
    for k, (M, want) in enumerate(cases):
        M = sp.csc_matrix(M)
        M.eliminate_zeros()
        for trans in (0, 1):
            for fmt in (0, 1):
                seen.add(run_checked(emu, orc, M, dtype, fmt, trans, 1.0, 0.0, seed=k, kernel=2,
                                     want_tiles=want if (trans == 0) else None))
            run_checked(emu, orc, M, dtype, 0, trans, 2.0, -0.5, seed=60 + k, kernel=2, num_sms=1)
    assert {0, 3, 5} <= seen
    # nnz % 4 in {1, 2, 3}: drop entries from the end of the storage
    base = random_sparse(400, 300, 0.06, 9, dt).tocsr()
    for cut in (1, 2, 3):
        M = base.copy()
        M.data[-cut:] = 0
        M.eliminate_zeros()
        assert M.nnz % 4 != base.nnz % 4 or cut == 0
        for trans in (0, 1):
            run_checked(emu, orc, sp.csc_matrix(M), dtype, 1, trans, 1.0, 0.0, seed=cut, kernel=2)
    # edge: empty matrix / no rows -> nothing staged
    for shape in ((0, 5), (5, 0), (9, 9)):
        Z = sp.csc_matrix(shape, dtype=dt)
        for trans in (0, 1):
            run_checked(emu, orc, Z, dtype, 0, trans, 1.0, 0.0, kernel=2)
            run_checked(emu, orc, Z, dtype, 0, trans, 3.0, 2.0, seed=3, kernel=2)


@pytest.mark.parametrize("dtype", [F64, F32])
def test_pipelined_row_kernel_and_forced_lane_widths(emu, orc, dtype):
    """the software-pipelined row kernel (sparse_kernel = 3) gives the SAME BITS as the plain row kernel (same lane layout and
    summation order), also with more CTAs than rows, rows longer than one trip, empty rows; any forced lane-group width
    (sparse_lanes) stays within the oracle's tolerance for all three kernels"""
    dt = np.float64 if dtype == F64 else np.float32
    mats = [random_sparse(300, 200, 0.02, 4, dt, empty_rows=(0, 7, 299)), random_sparse(64, 500, 0.12, 5, dt, dense_row=10),
            random_sparse(2500, 40, 0.1, 6, dt), sp.identity(700, dtype=dt, format="csc") * 0.5, random_sparse(5, 3000, 0.6, 8, dt)]
    for k, M in enumerate(mats):
        M = sp.csc_matrix(M)
        M.eliminate_zeros()
        for trans in (0, 1):
            for (alpha, beta) in ((1.0, 0.0), (2.0, -0.5)):
                a, b = [], []
                run_checked(emu, orc, M, dtype, 0, trans, alpha, beta, seed=k, kernel=1, out=a)
                run_checked(emu, orc, M, dtype, 0, trans, alpha, beta, seed=k, kernel=3, out=b)
                assert np.array_equal(a[0], b[0], equal_nan=True), (k, trans)
        for lanes in (0, 2, 5):
            a, b = [], []
            run_checked(emu, orc, M, dtype, 1, 0, 1.0, 0.0, seed=k, kernel=1, force_lanes=lanes, out=a)
            run_checked(emu, orc, M, dtype, 1, 0, 1.0, 0.0, seed=k, kernel=3, force_lanes=lanes, out=b)
            assert np.array_equal(a[0], b[0], equal_nan=True)
            run_checked(emu, orc, M, dtype, 1, 0, 1.0, 0.0, seed=k, kernel=2, force_lanes=lanes)


def test_randomised_structures_all_kernels(emu, orc):
    """seeded random structures with power-law row lengths (a few very long rows, many empty ones), random shapes and both
    storage formats through all three kernels; the plain and the pipelined row kernel must agree bit for bit"""
    rng = np.random.default_rng(2024)
    for trial in range(12):
        m, n = int(rng.integers(1, 400)), int(rng.integers(1, 3000))
        lens = np.minimum(n, (rng.pareto(1.2, m) * 3).astype(np.int64))
        lens[rng.uniform(size=m) < 0.2] = 0
        rows = np.repeat(np.arange(m), lens)
        cols = np.concatenate([rng.choice(n, size=int(k), replace=False) for k in lens]) if lens.sum() else np.zeros(0, dtype=np.int64)
        vals = rng.uniform(-1, 1, rows.shape[0])
        dtype = F64 if trial % 2 == 0 else F32
        A = sp.csc_matrix((vals.astype(np.float64 if dtype == F64 else np.float32), (rows, cols)), shape=(m, n))
        A.sort_indices()
        fmt, trans = trial % 2, (trial // 2) % 2
        alpha, beta = (1.0, 0.0) if trial % 3 else (-1.5, 0.75)
        a, b = [], []
        run_checked(emu, orc, A, dtype, fmt, trans, alpha, beta, seed=trial, kernel=1, out=a)
        run_checked(emu, orc, A, dtype, fmt, trans, alpha, beta, seed=trial, kernel=3, out=b)
        assert np.array_equal(a[0], b[0], equal_nan=True), trial
        run_checked(emu, orc, A, dtype, fmt, trans, alpha, beta, seed=trial, kernel=2)
