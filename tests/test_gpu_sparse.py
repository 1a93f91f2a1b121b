"""GPU parity of the sparse-matrix leaf `LinearOperator(M::SparseMatrixCSC)` (src/constructors.jl:15-29, "dense or sparse"
:3-5) through the C ABI (b2o_sparse_*): the hand-written compressed-row kernels against the oracle's restatement of
SparseArrays' `mul!(res, M, v, α, β)` on the same inputs, the reference's sparse predicates (test/test_linop.jl:740-766:
BlockDiagonalOperator with a `sprand` block, issue #139 products; test/test_cat.jl:47 `[opEye(2); sparse(1.0I, 2, 2)]`) and,
far above L2, size-independent properties with cuSPARSE (torch.sparse.mm) as an independent implementation.
Bars: index work exact (structure transposition; products with 0/1 patterns are compared with ==), sums norm-wise relative
<= 1e-12 (Float64) / <= 1e-5 (Float32 storage, sums in double on both sides)."""
import numpy as np
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
TOL = {"float64": 1e-12, "float32": 1e-5}


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def host(t):
    return t.detach().cpu().numpy()


def tdtype(name):
    import torch
    return getattr(torch, name)


def rand_vec(ctx, n, dtype, seed):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1).to(dtype).to("cuda:%d" % ctx.device)


def random_sparse(m, n, density, seed, dt, dense_row=None, empty_rows=()):
    rng = np.random.default_rng(seed)
    A = sp.random(m, n, density=density, random_state=rng, format="lil", dtype=np.float64)
    if dense_row is not None and m > 0:
        A[dense_row, :] = rng.uniform(-1, 1, n)
    for r in empty_rows:
        if 0 <= r < m:
            A[r, :] = 0
    A = A.tocsc().astype(dt)
    A.eliminate_zeros()
    A.sort_indices()
    return A


def to_torch(ctx, A, fmt):
    """scipy matrix -> torch CUDA tensor in sparse_csc (Julia's SparseMatrixCSC) or sparse_csr layout"""
    import torch
    dev = "cuda:%d" % ctx.device
    if fmt == "csc":
        S = A.tocsc()
        S.sort_indices()
        return torch.sparse_csc_tensor(torch.from_numpy(S.indptr.astype(np.int64)), torch.from_numpy(S.indices.astype(np.int64)),
                                       torch.from_numpy(S.data), size=A.shape, device=dev)
    S = A.tocsr()
    S.sort_indices()
    return torch.sparse_csr_tensor(torch.from_numpy(S.indptr.astype(np.int64)), torch.from_numpy(S.indices.astype(np.int64)),
                                   torch.from_numpy(S.data), size=A.shape, device=dev)


def oracle_product(orc, A, v, alpha, beta, trans, res0):
    csc = A.tocsc()
    csc.sort_indices()
    ref = res0.copy()
    orc.spmv_csc_(ref, A.shape[0], A.shape[1], csc.indptr.astype(np.int64) + 1, csc.indices.astype(np.int64) + 1, csc.data, v,
                  alpha, beta, trans)
    return ref


CASES = [(10, 6, 0.5), (6, 10, 0.5), (200, 300, 0.004), (300, 200, 0.02), (150, 150, 0.05), (64, 500, 0.12), (40, 700, 0.3),
         (33, 900, 0.9), (1, 50, 0.5), (50, 1, 0.5), (257, 129, 0.03), (20011, 30011, 0.0004), (70001, 517, 0.01),
         (64, 3000, 0.01), (129, 2600, 0.5), (300, 5000, 0.004)]     # the last one gets a 5000-entry row: a direct tile


@pytest.fixture(params=[1, 2, 3], ids=["rowkernel", "tilekernel", "pipekernel"])
def kernel(request, ctx):
    """force the plain row kernel / the TMA-staged tile kernel / the software-pipelined row kernel (the default)"""
    ctx.set_option("sparse_kernel", request.param)
    yield request.param
    ctx.set_option("sparse_kernel", 0)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("fmt", ["csc", "csr"])
def test_sparse_apply_matches_oracle(lo, ctx, orc, dtype, fmt, kernel):
    """prod!/tprod!/ctprod! of LinearOperator(M) for every lane-group width (mean row length 1 ... >= 32), empty and dense rows,
    through both kernels"""
    import torch
    dt = tdtype(dtype)
    ndt = np.dtype(dtype)
    for k, (m, n, dens) in enumerate(CASES):
        A = random_sparse(m, n, dens, k, ndt, dense_row=(m // 2 if k % 3 == 0 else None), empty_rows=(0, m - 1) if k % 2 else ())
        M = to_torch(ctx, A, fmt)
        op = lo.LinearOperator(M)
        assert lo.size(op) == (m, n) and lo.eltype(op) == dt
        assert not lo.issymmetric(op) and not lo.ishermitian(op)
        v, u = rand_vec(ctx, n, dt, 100 + k), rand_vec(ctx, m, dt, 200 + k)
        for alpha, beta in ((1.0, 0.0), (2.0, -0.5)):
            r0 = rand_vec(ctx, m, dt, 300 + k)
            res = r0.clone() if beta != 0 else torch.full((m,), float("nan"), dtype=dt, device=v.device)
            lo.mul_(res, op, v, alpha, beta)
            ref = oracle_product(orc, A, host(v), alpha, beta, 0, host(r0))
            assert rel(host(res), ref) <= TOL[dtype], (m, n, fmt, alpha, beta)
            for wrap in (lo.transpose, lo.adjoint):
                t0 = rand_vec(ctx, n, dt, 400 + k)
                rt = t0.clone() if beta != 0 else torch.full((n,), float("nan"), dtype=dt, device=v.device)
                lo.mul_(rt, wrap(op), u, alpha, beta)
                reft = oracle_product(orc, A, host(u), alpha, beta, 1, host(t0))
                assert rel(host(rt), reft) <= TOL[dtype], (m, n, fmt, alpha, beta, wrap.__name__)
        assert lo.nprod(op) == 2 and lo.ntprod(op) == 2 and lo.nctprod(op) == 2


@pytest.mark.parametrize("fmt", ["csc", "csr"])
def test_sparse_index_work_is_exact(lo, ctx, fmt, kernel):
    """0/1 patterns and small-integer values: every product is exactly representable, so == must hold (structure
    transposition, pointer arithmetic, duplicate-free gather).  Includes Matrix(op) == A and Matrix(op') == A'."""
    import torch
    rng = np.random.default_rng(5)
    for m, n, dens in ((10, 10, 0.2), (37, 91, 0.1), (1000, 333, 0.02), (5, 4000, 0.5), (4001, 4003, 0.003)):
        A = sp.random(m, n, density=dens, random_state=rng, format="csc", data_rvs=lambda k: rng.integers(-8, 9, k).astype(np.float64))
        A.eliminate_zeros()
        M = to_torch(ctx, A, fmt)
        op = lo.LinearOperator(M)
        v = torch.from_numpy(rng.integers(-16, 17, n).astype(np.float64)).to(M.device)
        u = torch.from_numpy(rng.integers(-16, 17, m).astype(np.float64)).to(M.device)
        assert np.array_equal(host(op * v), A @ host(v))                    # opA * b == A * b        (test_linop.jl:763)
        assert np.array_equal(host(lo.transpose(op) * u), A.T @ host(u))    # transpose(opA) * b      (:764)
        assert np.array_equal(host(lo.adjoint(op) * u), A.T @ host(u))      # adjoint(opA) * b        (:765)
        if n <= 100:
            assert np.array_equal(host(lo.Matrix(op)), A.toarray())
            assert np.array_equal(host(lo.Matrix(lo.transpose(op))), A.T.toarray())


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_sparse_edge_cases_and_errors(lo, ctx, orc, dtype, kernel):
    """nnz == 0, empty dimensions, β == 0 never reads res; malformed structure, CPU / COO matrices, shape and dtype mismatch"""
    import ctypes
    import torch
    dt = tdtype(dtype)
    ndt = np.dtype(dtype)
    dev = "cuda:%d" % ctx.device
    for fmt in ("csc", "csr"):
        for shape in ((0, 5), (5, 0), (0, 0), (7, 7)):
            Z = sp.csc_matrix(shape, dtype=ndt)
            op = lo.LinearOperator(to_torch(ctx, Z, fmt))
            v = rand_vec(ctx, shape[1], dt, 1)
            res = torch.full((shape[0],), float("nan"), dtype=dt, device=dev)
            lo.mul_(res, op, v, 3.0, 0.0)
            assert torch.equal(res, torch.zeros_like(res))                  # 0 even from NaN-filled res
            r0 = rand_vec(ctx, shape[0], dt, 2)
            res = r0.clone()
            lo.mul_(res, op, v, 3.0, 2.0)
            assert torch.equal(res, 2.0 * r0)
            rt = torch.full((shape[1],), float("nan"), dtype=dt, device=dev)
            lo.mul_(rt, lo.transpose(op), rand_vec(ctx, shape[0], dt, 3), 1.0, 0.0)
            assert torch.equal(rt, torch.zeros_like(rt))
    D = sp.identity(3000, dtype=ndt, format="csc") * 2.5                     # one entry per row
    op = lo.LinearOperator(to_torch(ctx, D, "csc"))
    v = rand_vec(ctx, 3000, dt, 4)
    assert torch.equal(op * v, 2.5 * v) and torch.equal(lo.transpose(op) * v, 2.5 * v)
    # malformed structure through the raw C ABI (1-based host arrays, as a Julia caller passes colptr / rowval)
    lib, h = ctx.lib, ctypes.c_void_p()
    vals = torch.ones(3, dtype=dt, device=dev)
    code = 0 if dtype == "float64" else 1

    def create(ptr1, idx1, m=3, n=3, nnz=3):
        p = np.asarray(ptr1, dtype=np.int64)
        i = np.asarray(idx1, dtype=np.int64)
        return lib.b2o_sparse_create(ctx.handle, code, 0, m, n, nnz, ctypes.c_void_p(p.ctypes.data), ctypes.c_void_p(i.ctypes.data),
                                     ctypes.c_void_p(vals.data_ptr()), ctypes.byref(h))
    assert create([1, 2, 3, 4], [1, 2, 3]) == 0
    assert lib.b2o_sparse_destroy(h) == 0
    assert create([0, 1, 2, 3], [1, 2, 3]) != 0                             # 0-based pointers
    assert create([1, 3, 2, 4], [1, 2, 3]) != 0                             # not monotone
    assert create([1, 2, 3, 4], [1, 2, 4]) != 0                             # row index out of range
    assert create([1, 2, 3, 4], [0, 2, 3]) != 0
    assert create([1, 2, 3, 5], [1, 2, 3]) != 0                             # colptr[end] != nnz + 1
    assert b"sparse" in lib.b2o_last_error()
    A = random_sparse(20, 30, 0.2, 1, ndt)
    M = to_torch(ctx, A, "csc")
    with pytest.raises(lo.B2OError):
        lo.LinearOperator(M.cpu())                                          # no CPU fallback
    with pytest.raises(lo.B2OError):
        lo.SparseMatrixOperator(M.to_sparse_coo())                          # COO is not a SparseMatrixCSC
    op = lo.LinearOperator(M)
    with pytest.raises(lo.LinearOperatorException):
        op * rand_vec(ctx, 31, dt, 1)                                       # shape mismatch (operations.jl:23-24)
    with pytest.raises(lo.B2OError):
        op * rand_vec(ctx, 30, torch.float32 if dt == torch.float64 else torch.float64, 1)


def test_sparse_values_are_aliased_and_refresh(lo, ctx, orc):
    """the closures capture M (constructors.jl:25-27): changing nzval in place changes the operator.  The given orientation is
    read in place; refresh() re-gathers the transposed copy."""
    import torch
    for fmt in ("csc", "csr"):
        A = random_sparse(400, 300, 0.05, 9, np.float64)
        M = to_torch(ctx, A, fmt)
        op = lo.LinearOperator(M)
        v, u = rand_vec(ctx, 300, torch.float64, 1), rand_vec(ctx, 400, torch.float64, 2)
        y0, z0 = op * v, lo.transpose(op) * u
        M.values().mul_(3.0)                                                # in place: the tensor the operator aliases
        op.refresh()
        assert rel(host(op * v), 3.0 * host(y0)) <= 1e-15
        assert rel(host(lo.transpose(op) * u), 3.0 * host(z0)) <= 1e-15


def test_reference_sparse_predicates(lo, ctx, orc):
    """test/test_linop.jl:740-755 (BlockDiagonalOperator(A, B, C) with B dense 4x2 and C = sprand(2, 4, 0.5)), test/test_cat.jl:47
    (`[opEye(2); sparse(1.0I, 2, 2)]`), matrix promotion in + and * (operations.jl:159-160,218-219)."""
    import torch
    dev = "cuda:%d" % ctx.device
    rng = np.random.default_rng(3)
    dvals = torch.tensor([0.5, 0.25, 0.125], dtype=torch.float64, device=dev)
    Aop = lo.opDiagonal(dvals)                                              # the reference uses ldiv! by a diagonal Cholesky factor
    B = torch.from_numpy(rng.uniform(0, 1, (2, 4))).to(dev).t()             # column-major 4 x 2
    Cs = sp.random(2, 4, density=0.5, random_state=rng, format="csc")
    C = to_torch(ctx, Cs, "csc")
    D = np.zeros((9, 9))
    D[:3, :3] = np.diag(host(dvals))
    D[3:7, 3:5] = host(B)
    D[7:9, 5:9] = Cs.toarray()
    M = lo.BlockDiagonalOperator(Aop, B, C)
    assert lo.size(M) == (9, 9)
    assert np.linalg.norm(host(lo.Matrix(M)) - D) <= np.sqrt(np.finfo(float).eps) * np.linalg.norm(D)
    assert np.linalg.norm(host(lo.Matrix(lo.transpose(M))) - D.T) <= np.sqrt(np.finfo(float).eps) * np.linalg.norm(D)
    assert np.linalg.norm(host(lo.Matrix(lo.adjoint(M))) - D.T) <= np.sqrt(np.finfo(float).eps) * np.linalg.norm(D)
    I2 = to_torch(ctx, sp.identity(2, dtype=np.float64, format="csc"), "csc")
    K = lo.vcat(lo.opEye(2, ctx=ctx), I2)
    x = torch.tensor([1.5, -2.0], dtype=torch.float64, device=dev)
    assert np.array_equal(host(K * x), np.array([1.5, -2.0, 1.5, -2.0]))
    # promotion: op * M, M * op, op + M with a sparse M
    n = 500
    S = random_sparse(n, n, 0.02, 4, np.float64)
    Sm = to_torch(ctx, S, "csr")
    d = rand_vec(ctx, n, torch.float64, 5)
    Dg = lo.opDiagonal(d)
    v = rand_vec(ctx, n, torch.float64, 6)
    dn, vn = host(d), host(v)
    assert rel(host((Dg * Sm) * v), dn * (S @ vn)) <= 1e-13
    assert rel(host((Sm * Dg) * v), S @ (dn * vn)) <= 1e-13
    assert rel(host((Dg + Sm) * v), dn * vn + S @ vn) <= 1e-13
    assert rel(host(lo.transpose(Sm - Dg) * v), S.T @ vn - dn * vn) <= 1e-13
    hc = lo.hcat(Dg, Sm)
    v2 = rand_vec(ctx, 2 * n, torch.float64, 7)
    assert rel(host(hc * v2), dn * host(v2)[:n] + S @ host(v2)[n:]) <= 1e-13


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_sparse_large_matrix_properties(lo, ctx, orc, dtype):
    """a matrix far above L2 (2^21 rows, 24 entries per row: banded + random columns, 0.6 GB in Float64 CSR): product against
    the oracle on a strided row sample, against cuSPARSE (independent), linearity, <u, A v> == <Aᵀ u, v>, run-to-run bit
    determinism (fixed-order sums, no atomics)"""
    import torch
    dt = tdtype(dtype)
    dev = "cuda:%d" % ctx.device
    n, per_row = 1 << 21, 24
    g = torch.Generator(device=dev).manual_seed(11)
    rows = torch.arange(n, device=dev, dtype=torch.int64)
    band = (rows[:, None] + torch.arange(-6, 6, device=dev)[None, :]) % n                      # 12 banded columns
    rnd = torch.randint(0, n, (n, per_row - 12), generator=g, device=dev, dtype=torch.int64)   # 12 random columns
    cols = torch.sort(torch.cat([band, rnd], dim=1), dim=1).values
    # duplicate columns inside a row are legal for the kernels (entries are summed) but not canonical CSR: nudge them apart
    dup = torch.zeros_like(cols, dtype=torch.bool)
    dup[:, 1:] = cols[:, 1:] == cols[:, :-1]
    vals = (torch.rand((n, per_row), generator=g, device=dev, dtype=torch.float64) * 2 - 1).to(dt)
    vals[dup] = 0                                                                               # structural duplicates carry 0
    crow = torch.arange(0, n * per_row + 1, per_row, device=dev, dtype=torch.int64)
    M = torch.sparse_csr_tensor(crow, cols.reshape(-1), vals.reshape(-1), size=(n, n), device=dev)
    op = lo.LinearOperator(M)
    v, w, u = rand_vec(ctx, n, dt, 1), rand_vec(ctx, n, dt, 2), rand_vec(ctx, n, dt, 3)
    y = op * v
    z = lo.transpose(op) * u
    # oracle on a strided sample of rows
    pick = np.arange(0, n, 4099)
    csub = host(cols[pick]).astype(np.int64)
    vsub = host(vals[pick])
    sub = sp.csr_matrix((vsub.reshape(-1), csub.reshape(-1), np.arange(0, len(pick) * per_row + 1, per_row)), shape=(len(pick), n))
    ref = oracle_product(orc, sub, host(v), 1.0, 0.0, 0, np.empty(len(pick), dtype=np.dtype(dtype)))
    assert rel(host(y)[pick], ref) <= TOL[dtype]
    # cuSPARSE as an independent implementation of the full product
    yc = torch.sparse.mm(M, v[:, None])[:, 0]
    assert rel(host(y), host(yc)) <= (1e-12 if dtype == "float64" else 1e-5)
    tol = 1e-12 if dtype == "float64" else 1e-4
    assert rel(host(op * (v + w)), host(y + op * w)) <= tol
    lhs = float(torch.dot(u.double(), y.double()))
    rhs = float(torch.dot(z.double(), v.double()))
    assert abs(lhs - rhs) <= tol * max(abs(lhs), 1.0) * 10
    assert torch.equal(op * v, y) and torch.equal(lo.transpose(op) * u, z)
    # the TMA-staged tile kernel (opt-in) must agree with the default (pipelined row) kernel to rounding, and is deterministic too;
    # values whose storage is off a 16-byte boundary fall back to the row kernel
    ctx.set_option("sparse_kernel", 2)
    try:
        yr, zr = op * v, lo.transpose(op) * u
        assert torch.equal(op * v, yr)
    finally:
        ctx.set_option("sparse_kernel", 0)
    assert rel(host(yr), host(y)) <= (1e-14 if dtype == "float64" else 1e-6)
    assert rel(host(zr), host(z)) <= (1e-14 if dtype == "float64" else 1e-6)
    ctx.set_option("sparse_kernel", 1)                                                # plain row kernel: the same bits as the pipelined default
    try:
        assert torch.equal(op * v, y) and torch.equal(lo.transpose(op) * u, z)
        ctx.set_option("sparse_lanes", 1)                                             # another lane-group width: rounding only
        assert rel(host(op * v), host(y)) <= (1e-14 if dtype == "float64" else 1e-6)
    finally:
        ctx.set_option("sparse_lanes", -1)
        ctx.set_option("sparse_kernel", 0)
    off = torch.empty(n * per_row + 1, dtype=dt, device=dev)[1:]
    off.copy_(vals.reshape(-1))
    Moff = torch.sparse_csr_tensor(crow, cols.reshape(-1), off, size=(n, n), device=dev)
    if Moff.values().data_ptr() % 16 != 0:                                            # torch kept the view: unaligned storage
        ctx.set_option("sparse_kernel", 2)
        try:
            assert torch.equal(lo.LinearOperator(Moff) * v, y)                            # row kernel again: same bits
        finally:
            ctx.set_option("sparse_kernel", 0)
    assert op.apply_bytes() == n * per_row * (vals.element_size() + 4) + 8 * (n + 1) + 2 * n * vals.element_size()
