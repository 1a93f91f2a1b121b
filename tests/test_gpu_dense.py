"""GPU parity of the dense-matrix leaf `LinearOperator(M)` (src/constructors.jl:15-29) through the C ABI (b2o_dense_*):
the hand-written matrix-vector kernels against the oracle's `mul!(res, M, v, α, β)` on the same inputs, the reference's own
GPU test (test/gpu/nvidia.jl:8-21: BlockDiagonalOperator of three Float32 CUDA matrices) and the `LinearOperator(Matrix)`
predicates of test/test_linop.jl:41-75.  Bars: norm-wise relative <= 1e-12 (Float64), <= 1e-5 (Float32 storage; sums are
accumulated in double on both sides and rounded once)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = {"float64": 1e-12, "float32": 1e-5}


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def host(t):
    return t.detach().cpu().numpy()


def rand_matrix(ctx, m, n, dtype, seed, colmajor):
    """m x n CUDA matrix, column-major (Julia layout) or row-major (torch default), entries U[-1,1)"""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = (torch.rand((m, n), generator=g, dtype=torch.float64) * 2 - 1).to(dtype)
    A = A.to("cuda:%d" % ctx.device)
    if colmajor:
        A = A.t().contiguous().t()
        assert m <= 1 or n <= 1 or A.stride() == (1, m)
    return A


def rand_vec(ctx, n, dtype, seed):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1).to(dtype).to("cuda:%d" % ctx.device)


def tdtype(name):
    import torch
    return getattr(torch, name)


SHAPES = [(5, 5), (10, 6), (1, 9), (9, 1), (515, 70), (70, 1101), (2051, 13), (3, 40000), (100003, 5), (1283, 1031)]


@pytest.mark.parametrize("dtype", ["float64", "float32"])
@pytest.mark.parametrize("colmajor", [True, False])
def test_dense_apply_matches_oracle(lo, ctx, orc, dtype, colmajor):
    """prod!/tprod!/ctprod! of LinearOperator(M): every kernel path (N / T, split / unsplit, both layouts)"""
    import torch
    dt = tdtype(dtype)
    for k, (m, n) in enumerate(SHAPES):
        A = rand_matrix(ctx, m, n, dt, 10 + k, colmajor)
        op = lo.LinearOperator(A)
        assert lo.size(op) == (m, n) and lo.eltype(op) == dt
        assert not lo.issymmetric(op) and not lo.ishermitian(op)                # defaultsymmetric(M) = false
        v, u = rand_vec(ctx, n, dt, 100 + k), rand_vec(ctx, m, dt, 200 + k)
        An = host(A)
        for alpha, beta in ((1.0, 0.0), (2.0, -0.5)):
            r0 = rand_vec(ctx, m, dt, 300 + k)
            res = r0.clone() if beta != 0 else torch.full((m,), float("nan"), dtype=dt, device=A.device)
            lo.mul_(res, op, v, alpha, beta)                                    # mul!(res, M, v, α, β)
            ref = host(r0).copy()
            orc.gemv_(ref, An, host(v), alpha, beta, 0)
            assert rel(host(res), ref) <= TOL[dtype], (m, n, alpha, beta)
            for wrap in (lo.transpose, lo.adjoint):                             # mul!(res, transpose(M) / adjoint(M), u, α, β)
                t0 = rand_vec(ctx, n, dt, 400 + k)
                rt = t0.clone() if beta != 0 else torch.full((n,), float("nan"), dtype=dt, device=A.device)
                lo.mul_(rt, wrap(op), u, alpha, beta)
                reft = host(t0).copy()
                orc.gemv_(reft, An, host(u), alpha, beta, 1)
                assert rel(host(rt), reft) <= TOL[dtype], (m, n, alpha, beta, wrap.__name__)
        assert lo.nprod(op) == 2 and lo.ntprod(op) == 2 and lo.nctprod(op) == 2     # counters (test_linop.jl:634-673)


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_dense_unaligned_views_and_scalar_kernels(lo, ctx, orc, dtype):
    """sub-matrix views (leading dimension > nrow, base off a 16-byte boundary), unaligned vectors, forced scalar kernels"""
    import torch
    dt = tdtype(dtype)
    big = rand_matrix(ctx, 1200, 90, dt, 1, True)
    cases = [big[3:1033, 2:77], big[4:1034, :75], big[:, 1:], big[1:, :]]
    for k, A in enumerate(cases):
        m, n = A.shape
        op = lo.LinearOperator(A)
        v = rand_vec(ctx, n + 1, dt, 50 + k)[1:]                                 # odd offset
        u = rand_vec(ctx, m + 1, dt, 60 + k)[1:]
        for forced in (0, 1):
            ctx.set_option("dense_scalar", forced)
            try:
                ref = np.empty(m, dtype=host(A).dtype)
                orc.gemv_(ref, host(A), host(v), 1.0, 0.0, 0)
                assert rel(host(op * v), ref) <= TOL[dtype]
                reft = np.empty(n, dtype=host(A).dtype)
                orc.gemv_(reft, host(A), host(u), 1.0, 0.0, 1)
                assert rel(host(lo.transpose(op) * u), reft) <= TOL[dtype]
            finally:
                ctx.set_option("dense_scalar", 0)
    with pytest.raises(lo.B2OError):
        lo.LinearOperator(big[::2, ::2])                                        # no unit stride
    with pytest.raises(lo.B2OError):
        lo.LinearOperator(big.cpu())                                            # no CPU fallback
    op = lo.LinearOperator(big)
    with pytest.raises(lo.LinearOperatorException):
        op * rand_vec(ctx, 91, dt, 1)                                           # shape mismatch (test_linop.jl:27)
    with pytest.raises(lo.B2OError):
        op * rand_vec(ctx, 90, torch.float32 if dt == torch.float64 else torch.float64, 1)


def test_reference_gpu_test_block_diagonal_of_float32_matrices(lo, ctx, orc):
    """test/gpu/nvidia.jl:8-21: M = BlockDiagonalOperator(A, B, C) with CUDA.rand (Float32) 5x5, 10x10, 20x20 blocks,
    y = M * v is a Float32 device vector; storage_type is preserved by adjoint/transpose."""
    import torch
    A, B, C = (rand_matrix(ctx, k, k, torch.float32, k, True) for k in (5, 10, 20))
    M = lo.BlockDiagonalOperator(A, B, C)
    assert lo.size(M) == (35, 35)
    v = rand_vec(ctx, 35, torch.float32, 7)
    y = M * v
    assert y.is_cuda and y.dtype == torch.float32 and y.shape == (35,)
    ref = np.concatenate([host(X).astype(np.float64) @ host(v)[a:b].astype(np.float64)
                          for X, a, b in ((A, 0, 5), (B, 5, 15), (C, 15, 35))])
    assert rel(host(y), ref) <= 1e-6
    yo = np.empty(35, dtype=np.float32)
    for X, a, b in ((A, 0, 5), (B, 5, 15), (C, 15, 35)):
        orc.gemv_(yo[a:b], host(X), host(v)[a:b], 1.0, 0.0, 0)
    assert rel(host(y), yo) <= 1e-6
    yt = lo.transpose(M) * v
    reft = np.concatenate([host(X).astype(np.float64).T @ host(v)[a:b].astype(np.float64)
                           for X, a, b in ((A, 0, 5), (B, 5, 15), (C, 15, 35))])
    assert rel(host(yt), reft) <= 1e-6
    opA = lo.LinearOperator(A)
    assert lo.storage_type(opA) == lo.storage_type(lo.adjoint(opA)) == lo.storage_type(lo.transpose(opA))
    assert lo.storage_type(opA).dtype == torch.float32


def test_linear_operator_of_matrix_predicates(lo, ctx, orc):
    """test/test_linop.jl:41-75 in real arithmetic: Matrix(op) == A, transposes, products with vectors and with matrix
    right-hand sides, `Constructor with specified structure` (:101-124)."""
    import torch
    nrow, ncol = 10, 6
    A = rand_matrix(ctx, nrow, ncol, torch.float64, 3, True)
    An = host(A)
    op = lo.LinearOperator(A)
    assert (op.nrow, op.ncol) == (nrow, ncol)
    assert rel(host(lo.Matrix(op)), An) <= 1e-15
    assert rel(host(lo.Matrix(lo.transpose(op))), An.T) <= 1e-15
    assert rel(host(lo.Matrix(lo.adjoint(op))), An.T) <= 1e-15
    v, u = rand_vec(ctx, ncol, torch.float64, 4), rand_vec(ctx, nrow, torch.float64, 5)
    assert rel(host(op * v), An @ host(v)) <= 1e-14
    assert rel(host(lo.transpose(op) * u), An.T @ host(u)) <= 1e-14
    mv = torch.stack([v, -2 * v]).t()                                           # column-major ncol x 2 (hcat(v, -2v))
    mu = torch.stack([u, -2 * u]).t()
    res_mat, res_trans = torch.empty((2, nrow), dtype=torch.float64, device=A.device).t(), \
        torch.empty((2, ncol), dtype=torch.float64, device=A.device).t()
    lo.mul_(res_mat, op, mv)
    lo.mul_(res_trans, lo.transpose(op), mu)
    assert rel(host(res_mat), An @ host(mv)) <= 1e-14
    assert rel(host(res_trans), An.T @ host(mu)) <= 1e-14
    # specified structure: a symmetric matrix declared symmetric/hermitian routes transpose/adjoint to prod!
    S = A[:6, :6] + A[:6, :6].t()
    S = S.t().contiguous().t()
    ops = lo.LinearOperator(S, symmetric=True, hermitian=True)
    w = rand_vec(ctx, 6, torch.float64, 6)
    for o in (ops, lo.transpose(ops), lo.adjoint(ops)):
        assert rel(host(o * w), host(S) @ host(w)) <= 1e-14
    assert lo.nprod(ops) == 3 and lo.ntprod(ops) == 0


def test_matrix_promotion_in_operator_algebra(lo, ctx, orc):
    """op * M, M * op, op ± M, M ± op (src/operations.jl:159-160,218-219,229-230), hcat/vcat with matrices (src/cat.jl:3-5,61-63)"""
    import torch
    n = 300
    A = rand_matrix(ctx, n, n, torch.float64, 1, True)
    Brm = rand_matrix(ctx, n, n, torch.float64, 2, False)                        # row-major operand
    d = rand_vec(ctx, n, torch.float64, 3)
    D = lo.opDiagonal(d)
    v = rand_vec(ctx, n, torch.float64, 4)
    An, Bn, dn, vn = host(A), host(Brm), host(d), host(v)
    assert rel(host((D * A) * v), dn * (An @ vn)) <= 1e-13
    assert rel(host((A * D) * v), An @ (dn * vn)) <= 1e-13
    assert rel(host((D + Brm) * v), dn * vn + Bn @ vn) <= 1e-13
    assert rel(host((Brm + D) * v), dn * vn + Bn @ vn) <= 1e-13
    assert rel(host((D - A) * v), dn * vn - An @ vn) <= 1e-13
    assert rel(host((A - D) * v), An @ vn - dn * vn) <= 1e-13
    assert rel(host(lo.transpose(D * Brm) * v), Bn.T @ (dn * vn)) <= 1e-13
    hc = lo.hcat(D, A)
    v2 = rand_vec(ctx, 2 * n, torch.float64, 5)
    assert rel(host(hc * v2), dn * host(v2)[:n] + An @ host(v2)[n:]) <= 1e-13
    vc = lo.vcat(Brm, D)
    assert rel(host(vc * v), np.concatenate([Bn @ vn, dn * vn])) <= 1e-13
    assert rel(host(lo.transpose(vc) * v2), Bn.T @ host(v2)[:n] + dn * host(v2)[n:]) <= 1e-13


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_dense_large_matrix_properties(lo, ctx, orc, dtype):
    """a matrix far larger than L2 (8192 x 16384: 1 GiB in Float64): the product against a strided sample of oracle rows,
    linearity, <u, A v> == <Aᵀ u, v>, run-to-run bit determinism"""
    import torch
    dt = tdtype(dtype)
    m, n = 8192, 16384
    A = (torch.rand((n, m), dtype=dt, device="cuda:%d" % ctx.device) * 2 - 1).t()      # column-major m x n
    op = lo.LinearOperator(A)
    v, w = rand_vec(ctx, n, dt, 1), rand_vec(ctx, n, dt, 2)
    u = rand_vec(ctx, m, dt, 3)
    y = op * v
    rows = np.arange(0, m, 257)
    sample = host(A[rows, :]).astype(np.float64) @ host(v).astype(np.float64)
    assert rel(host(y)[rows], sample) <= TOL[dtype]
    cols = np.arange(0, n, 509)
    z = lo.transpose(op) * u
    sample_t = host(A[:, cols]).astype(np.float64).T @ host(u).astype(np.float64)
    assert rel(host(z)[cols], sample_t) <= TOL[dtype]
    tol = 1e-12 if dtype == "float64" else 1e-4
    assert rel(host(op * (v + w)), host(y + op * w)) <= tol                       # linearity
    lhs = float(torch.dot(u.double(), y.double()))
    rhs = float(torch.dot(z.double(), v.double()))
    assert abs(lhs - rhs) <= tol * max(abs(lhs), 1.0) * 10                        # <u, A v> == <Aᵀ u, v>
    assert torch.equal(op * v, y) and torch.equal(lo.transpose(op) * u, z)       # deterministic (fixed-order split sums)
