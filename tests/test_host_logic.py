"""Host-side dispatch layer (linearoperators.jl_b200/abstract.py, cat.py) driven on CPU with user closures over
numpy arrays -- the same way the reference tests its own dispatch with Julia closures:
  test/test_linop.jl:634-673 (counters), :768-891 (3-arg closures / prod3!), test/test_adjtrans.jl:10-37,
  test/test_cat.jl:4-50, test/test_linop.jl:25-26,201-202 (shape errors)."""
import numpy as np
import pytest


def dense_op(lo, M, args5=True, with_t=True):
    M = np.asarray(M, dtype=float)
    S = lo.Storage("numpy")
    if args5:
        def prod(res, v, a, b):
            res[:] = a * (M @ v) + (b * res if b != 0 else 0)

        def tprod(res, u, a, b):
            res[:] = a * (M.T @ u) + (b * res if b != 0 else 0)
    else:
        def prod(res, v):
            res[:] = M @ v

        def tprod(res, u):
            res[:] = M.T @ u
    return lo.LinearOperator(np.float64, M.shape[0], M.shape[1], False, False, prod, tprod if with_t else None,
                             tprod if with_t else None, S=S)


def test_mul_shape_and_counters(lo):
    rng = np.random.default_rng(0)
    A = rng.random((5, 3))
    op = dense_op(lo, A)
    v, u = rng.random(3), rng.random(5)
    assert np.allclose(op * v, A @ v)
    assert lo.nprod(op) == 1 and lo.ntprod(op) == 0 and lo.nctprod(op) == 0
    assert np.allclose(lo.transpose(op) * u, A.T @ u)
    assert lo.ntprod(op) == 1
    assert np.allclose(lo.adjoint(op) * u, A.T @ u)
    assert lo.nctprod(op) == 1
    # wrappers remap counters (src/adjtrans.jl:46-58)
    assert lo.nprod(lo.adjoint(op)) == lo.nctprod(op)
    assert lo.nprod(lo.transpose(op)) == lo.ntprod(op)
    lo.reset_(op)
    assert (lo.nprod(op), lo.ntprod(op), lo.nctprod(op)) == (0, 0, 0)
    with pytest.raises(lo.LinearOperatorException, match="shape mismatch"):
        op * u
    with pytest.raises(lo.LinearOperatorException, match="shape mismatch"):
        lo.mul_(np.empty(4), op, v)
    with pytest.raises(lo.LinearOperatorException, match="shape mismatch"):
        lo.adjoint(op) * v
    assert lo.size(op) == (5, 3) and lo.size(lo.adjoint(op)) == (3, 5) and lo.size(op, 2) == 3
    with pytest.raises(lo.LinearOperatorException):
        lo.size(op, 3)


def test_five_arg_mul(lo):
    rng = np.random.default_rng(1)
    A = rng.random((4, 4))
    op = dense_op(lo, A)
    v, r0 = rng.random(4), rng.random(4)
    res = r0.copy()
    lo.mul_(res, op, v, 2.0, -3.0)
    assert np.allclose(res, 2 * A @ v - 3 * r0)
    res = np.full(4, np.nan)
    lo.mul_(res, op, v, 2.0, 0.0)                 # beta == 0: res is never read
    assert np.allclose(res, 2 * A @ v)


def test_three_arg_closures_prod3(lo):
    # src/operations.jl:10-20: Mv is allocated lazily only when beta != 0
    rng = np.random.default_rng(2)
    A = rng.random((4, 6))
    op = dense_op(lo, A, args5=False)
    assert not lo.has_args5(op) and not lo.isallocated5(op)
    v, r0 = rng.random(6), rng.random(4)
    res = np.empty(4)
    lo.mul_(res, op, v)
    assert np.allclose(res, A @ v) and not lo.isallocated5(op)
    lo.mul_(res, op, v, 3.0, 0.0)
    assert np.allclose(res, 3 * A @ v) and not lo.isallocated5(op)
    res = r0.copy()
    lo.mul_(res, op, v, 3.0, 0.5)
    assert np.allclose(res, 3 * A @ v + 0.5 * r0) and lo.isallocated5(op)
    u, t0 = rng.random(4), rng.random(6)
    res = t0.copy()
    lo.mul_(res, lo.transpose(op), u, 2.0, 2.0)
    assert np.allclose(res, 2 * A.T @ u + 2 * t0)


def test_adjoint_algebra(lo):
    # test/test_adjtrans.jl:10-37
    op = dense_op(lo, np.arange(6.0).reshape(2, 3))
    A, T, C = lo.adjoint(op), lo.transpose(op), lo.conj(op)
    assert lo.adjoint(A) is op and lo.transpose(T) is op and lo.conj(C) is op
    assert isinstance(lo.adjoint(T), lo.ConjugateLinearOperator) and lo.adjoint(T).parent is op
    assert isinstance(lo.transpose(A), lo.ConjugateLinearOperator)
    assert isinstance(lo.conj(A), lo.TransposeLinearOperator)
    assert isinstance(lo.conj(T), lo.AdjointLinearOperator)
    assert isinstance(lo.adjoint(C), lo.TransposeLinearOperator)
    assert isinstance(lo.transpose(C), lo.AdjointLinearOperator)
    v = np.array([1.0, -1.0, 2.0])
    assert np.allclose(C * v, op * v)


def test_adjoint_inference(lo):
    M = np.array([[2.0, 1.0], [1.0, 3.0]])
    S = lo.Storage("numpy")
    prod = lambda res, v, a, b: res.__setitem__(slice(None), a * (M @ v) + (b * res if b != 0 else 0))
    sym = lo.LinearOperator(np.float64, 2, 2, True, False, prod, None, None, S=S)
    v = np.array([1.0, -2.0])
    assert np.allclose(lo.transpose(sym) * v, M @ v)       # symmetric shortcut
    assert np.allclose(lo.adjoint(sym) * v, M @ v)         # inferred through prod! (conj sandwich is a no-op for real)
    assert lo.nprod(sym) == 2
    nosym = lo.LinearOperator(np.float64, 2, 2, False, False, prod, None, None, S=S)
    with pytest.raises(lo.LinearOperatorException, match="unable to infer conjugate transpose"):
        lo.adjoint(nosym) * v
    with pytest.raises(lo.LinearOperatorException, match="unable to infer transpose"):
        lo.transpose(nosym) * v
    herm = lo.LinearOperator(np.float64, 2, 2, False, True, prod, None, None, S=S)
    assert np.allclose(lo.adjoint(herm) * v, M @ v)
    assert np.allclose(lo.transpose(herm) * v, M @ v)      # ctprod inferred from hermitian


def test_operator_algebra_against_dense(lo):
    rng = np.random.default_rng(3)
    A, B, C = rng.random((4, 5)), rng.random((4, 5)), rng.random((5, 3))
    a, b, c = dense_op(lo, A), dense_op(lo, B), dense_op(lo, C)
    v5, v3, u4 = rng.random(5), rng.random(3), rng.random(4)
    for op, M, v in [(a + b, A + B, v5), (a - b, A - B, v5), (a * c, A @ C, v3), (2.5 * a, 2.5 * A, v5), (a * 2.5, 2.5 * A, v5),
                     (-a, -A, v5), (a / 4, A / 4, v5), ((a + b) * c, (A + B) @ C, v3), (+a, A, v5)]:
        assert np.allclose(op * v, M @ v)
        assert np.allclose(lo.Matrix(op, like=v), M)
    for op, M in [(lo.transpose(a * c), (A @ C).T), (lo.adjoint(a + b), (A + B).T), (lo.transpose(-a), -A.T),
                  (lo.transpose(3 * a), 3 * A.T), (lo.adjoint(a) * 2, 2 * A.T)]:
        assert np.allclose(op * u4, M @ u4)
    with pytest.raises(lo.LinearOperatorException, match="shape mismatch"):
        a * b
    with pytest.raises(lo.LinearOperatorException, match="shape mismatch"):
        a + c
    sq = dense_op(lo, rng.random((4, 4)))
    Hs = lo.Hermitian(sq)
    assert np.allclose(lo.Matrix(Hs, like=u4), lo.Matrix(Hs, like=u4).T)
    with pytest.raises(lo.LinearOperatorException, match="not square"):
        lo.Symmetric(a)
    assert lo.opEye() * v5 is v5 and (lo.opEye() * a) is a


def test_nested_counters(lo):
    # every node of the closure tree counts its own applies
    rng = np.random.default_rng(4)
    a, b = dense_op(lo, rng.random((3, 3))), dense_op(lo, rng.random((3, 3)))
    s = a * b + a
    s * rng.random(3)
    assert lo.nprod(s) == 1 and lo.nprod(a) == 2 and lo.nprod(b) == 1


def test_cat(lo):
    # test/test_cat.jl:4-50
    rng = np.random.default_rng(5)
    A, B, C = rng.random((4, 3)), rng.random((4, 2)), rng.random((6, 3))
    a, b, c = dense_op(lo, A), dense_op(lo, B), dense_op(lo, C)
    H = lo.hcat(a, b)
    V = lo.vcat(a, c)
    v5, v3, u4, u10 = rng.random(5), rng.random(3), rng.random(4), rng.random(10)
    assert np.allclose(H * v5, np.hstack([A, B]) @ v5)
    assert np.allclose(lo.transpose(H) * u4, np.hstack([A, B]).T @ u4)
    assert np.allclose(V * v3, np.vstack([A, C]) @ v3)
    assert np.allclose(lo.adjoint(V) * u10, np.vstack([A, C]).T @ u10)
    r0 = rng.random(4)
    res = r0.copy()
    lo.mul_(res, H, v5, 2.0, 3.0)
    assert np.allclose(res, 2 * np.hstack([A, B]) @ v5 + 3 * r0)
    with pytest.raises(lo.LinearOperatorException, match="hcat: inconsistent row sizes"):
        lo.hcat(a, c)
    with pytest.raises(lo.LinearOperatorException, match="vcat: inconsistent column sizes"):
        lo.vcat(a, b)
    D = rng.random((6, 2))
    HV = lo.hvcat((2, 2), a, b, c, dense_op(lo, D))
    assert np.allclose(HV * v5, np.block([[A, B], [C, D]]) @ v5)


def test_block_diagonal_host(lo):
    # test/test_linop.jl:718-756
    rng = np.random.default_rng(6)
    A, B = rng.random((3, 2)), rng.random((2, 4))
    bd = lo.BlockDiagonalOperator(dense_op(lo, A), dense_op(lo, B))
    M = np.block([[A, np.zeros((3, 4))], [np.zeros((2, 2)), B]])
    v, u = rng.random(6), rng.random(5)
    assert lo.size(bd) == (5, 6)
    assert np.allclose(bd * v, M @ v)
    assert np.allclose(lo.transpose(bd) * u, M.T @ u)
    assert np.allclose(lo.adjoint(bd) * u, M.T @ u)


def test_storage_promotion(lo):
    a = dense_op(lo, np.eye(2))
    b = dense_op(lo, np.eye(2))
    b.S = lo.Storage("cuda", 0)
    with pytest.raises(lo.LinearOperatorException, match="cannot be promoted"):
        a * b
    with pytest.raises(lo.LinearOperatorException, match="cannot be promoted"):
        a + b


def test_dense_matrix_layout_mapping(lo):
    """LinearOperator(M): how a strided 2-D tensor maps onto the column-major matrix of the C ABI (constructors.colmajor_view);
    checked by addressing: element (i, j) of M must be element (i, j) [or (j, i) when swapped] of the column-major view."""
    import torch
    from linearoperators_jl_b200.constructors import colmajor_view
    base = torch.arange(40 * 30, dtype=torch.float64).reshape(40, 30)
    views = [base, base.t(), base[3:20, 5:9], base.t()[2:7, 1:30], base[:, 4:5], base[7:8, :], base.t()[:, 3:4],
             base[:1, :1], base[5:5, :], base[:, 2:2], base.t().contiguous().t()]
    for M in views:
        m, n = M.shape
        flat = torch.as_strided(M, (M.untyped_storage().nbytes() // 8,), (1,), 0)     # the tensor's whole storage
        (rows, cols, lda), swap = colmajor_view(m, n, *M.stride())
        assert (rows, cols) == ((n, m) if swap else (m, n)) and lda >= max(1, rows)
        off = M.storage_offset()
        for i in range(m):
            for j in range(n):
                r, c = (j, i) if swap else (i, j)
                assert flat[off + r + c * lda] == M[i, j]
    with pytest.raises(lo.B2OError):
        colmajor_view(20, 15, 2, 60)
    with pytest.raises(lo.B2OError):
        lo.LinearOperator(base)                      # CPU tensor: no CPU fallback
    # the closure-based constructor is untouched by the matrix overload
    op = lo.LinearOperator(np.float64, 2, 2, True, True, lambda res, v, a, b: None)
    assert type(op) is lo.LinearOperator and lo.size(op) == (2, 2)


def test_sparse_matrix_leaf_refuses_cpu_and_coo(lo):
    """LinearOperator(M) with a sparse M dispatches to the sparse leaf (constructors.jl:3-5 "dense or sparse"); there is no CPU
    fallback and only the compressed layouts (Julia's SparseMatrixCSC, or CSR) are accepted"""
    import torch
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        S = torch.sparse_csc_tensor(torch.tensor([0, 1, 2]), torch.tensor([0, 1]), torch.tensor([1.0, 2.0]), size=(2, 2))
        with pytest.raises(lo.B2OError, match="CUDA"):
            lo.LinearOperator(S)
        with pytest.raises(lo.B2OError, match="sparse_csc or sparse_csr"):
            lo.SparseMatrixOperator(S.to_sparse_coo())


def test_kron_of_operators_against_dense_kron(lo):
    """kron(A, B) for two arbitrary operators (src/kron.jl:10-49) driven with host closures: `Matrix(K)` against np.kron
    (test/test_kron.jl:3-39: norm(Matrix(K) - kron(A, B)) <= eps * norm), transpose / adjoint, 5-arg form, repeated applies,
    symmetry flags, shape errors; scaling (test/test_kron.jl:50-58: kron(2A, B) == 2 kron(A, B))"""
    rng = np.random.default_rng(5)
    for (m, n, p, q) in ((2, 3, 4, 2), (3, 3, 2, 5), (1, 4, 3, 1), (10, 10, 3, 3)):
        A, B = rng.uniform(-1, 1, (m, n)), rng.uniform(-1, 1, (p, q))
        K = lo.kron(dense_op(lo, A), dense_op(lo, B))
        D = np.kron(A, B)
        assert lo.size(K) == (m * p, n * q)
        x, u = rng.uniform(-1, 1, n * q), rng.uniform(-1, 1, m * p)
        assert np.linalg.norm(lo.Matrix(K, like=x) - D, 1) <= 10 * np.finfo(float).eps * np.linalg.norm(D, 1)
        assert np.linalg.norm(lo.Matrix(lo.transpose(K), like=u) - D.T, 1) <= 10 * np.finfo(float).eps * np.linalg.norm(D, 1)
        assert np.allclose(lo.adjoint(K) * u, D.T @ u, rtol=1e-13, atol=1e-14)
        r0 = rng.uniform(-1, 1, m * p)
        res = r0.copy()
        lo.mul_(res, K, x, 2.0, -0.5)
        assert np.allclose(res, 2.0 * (D @ x) - 0.5 * r0, rtol=1e-13, atol=1e-14)
        res = np.full(m * p, np.nan)
        lo.mul_(res, K, x)                                                    # β == 0 never reads res
        assert np.allclose(res, D @ x, rtol=1e-13, atol=1e-14)
        y = x.copy()
        if m * p == n * q:
            for _ in range(100):                                             # 100 applies stay accurate (test_kron.jl:26-38)
                y = K * y
                y /= np.linalg.norm(y)
            z = x.copy()
            for _ in range(100):
                z = D @ z
                z /= np.linalg.norm(z)
            assert np.linalg.norm(y - z) <= 1e-10
        K2 = lo.kron(2.0 * dense_op(lo, A), dense_op(lo, B))
        assert np.allclose(K2 * x, 2.0 * (K * x), rtol=1e-13, atol=1e-14)
        with pytest.raises(lo.LinearOperatorException, match="shape mismatch"):
            K * np.zeros(n * q + 1)
    S1, S2 = rng.uniform(-1, 1, (3, 3)), rng.uniform(-1, 1, (2, 2))
    sym = lambda M: lo.LinearOperator(np.float64, M.shape[0], M.shape[0], True, True,
                                      lambda res, v, a, b: res.__setitem__(slice(None), a * ((M + M.T) @ v) + (b * res if b != 0 else 0)),
                                      S=lo.Storage("numpy"))
    Ks = lo.kron(sym(S1), sym(S2))
    assert lo.issymmetric(Ks) and lo.ishermitian(Ks)
    assert not lo.issymmetric(lo.kron(dense_op(lo, S1), sym(S2)))
    xs = rng.uniform(-1, 1, 6)
    assert np.allclose(lo.transpose(Ks) * xs, np.kron(S1 + S1.T, S2 + S2.T) @ xs, rtol=1e-13)


def test_kron_of_operators_torch_tensor_plumbing(lo):
    """the same kron(A, B) composition with torch tensors (CPU here) instead of numpy arrays: exercises the reshape / t() /
    contiguous() plumbing the CUDA path takes, with host closures standing in for the kernels"""
    import torch

    class TorchCPU(lo.Storage):
        def __init__(self):
            super().__init__("torchcpu", 0, torch.float64)

        def alloc(self, n, zero=False):
            return (torch.zeros if zero else torch.empty)(int(n), dtype=torch.float64)

    def top(M):
        Mt = torch.as_tensor(M)

        def prod(res, v, a, b):
            assert v.dim() == 1 and (v.numel() <= 1 or v.stride(0) == 1), "closures need unit-stride vectors"
            assert res.numel() <= 1 or res.stride(0) == 1
            res[:] = a * (Mt @ v) + (b * res if b != 0 else 0)

        def tprod(res, u, a, b):
            assert u.numel() <= 1 or u.stride(0) == 1
            assert res.numel() <= 1 or res.stride(0) == 1
            res[:] = a * (Mt.t() @ u) + (b * res if b != 0 else 0)
        return lo.LinearOperator(torch.float64, M.shape[0], M.shape[1], False, False, prod, tprod, tprod, S=TorchCPU())

    rng = np.random.default_rng(6)
    for (m, n, p, q) in ((2, 3, 4, 2), (5, 1, 1, 6), (7, 7, 3, 4)):
        A, B = rng.uniform(-1, 1, (m, n)), rng.uniform(-1, 1, (p, q))
        K = lo.kron(top(A), top(B))
        D = np.kron(A, B)
        x, u = torch.as_tensor(rng.uniform(-1, 1, n * q)), torch.as_tensor(rng.uniform(-1, 1, m * p))
        assert np.allclose((K * x).numpy(), D @ x.numpy(), rtol=1e-13, atol=1e-14)
        assert np.allclose((lo.transpose(K) * u).numpy(), D.T @ u.numpy(), rtol=1e-13, atol=1e-14)
        r0 = torch.as_tensor(rng.uniform(-1, 1, m * p))
        res = r0.clone()
        lo.mul_(res, K, x, 1.5, 0.25)
        assert np.allclose(res.numpy(), 1.5 * (D @ x.numpy()) + 0.25 * r0.numpy(), rtol=1e-13, atol=1e-14)


def test_quasi_newton_constructor_forms_and_element_types():
    """LBFGSOperator(T, n; ...) / LBFGSOperator(n; ...) (src/lbfgs.jl:168, 208; src/lsr1.jl:86, 115): the positional element type of the
    reference, the keyword form of the mirror, Float64 / Float32 built, anything else refused -- pure host logic, no GPU"""
    import torch
    from linearoperators_jl_b200 import _lib, qn
    assert qn._split_T(10, (), None) == (10, None)
    assert qn._split_T(torch.float32, (10,), None) == (10, torch.float32)
    assert qn._split_T(10, (), torch.float32) == (10, torch.float32)
    with pytest.raises(TypeError):
        qn._split_T(10, (5,), None)                         # mem is a keyword argument in the reference too
    with pytest.raises(TypeError):
        qn._split_T(torch.float32, (10,), torch.float64)     # element type given twice
    with pytest.raises(TypeError):
        qn._split_T(torch.float32, (10, 5), None)
    assert qn._eltype_code(None) == (torch.float64, _lib.B2O_F64)
    assert qn._eltype_code(float) == (torch.float64, _lib.B2O_F64)
    assert qn._eltype_code(np.float32) == (torch.float32, _lib.B2O_F32)
    assert qn._eltype_code(torch.float32) == (torch.float32, _lib.B2O_F32)
    for bad in (torch.float16, torch.bfloat16, torch.complex128):   # Float16 / BigFloat / complex of the reference's precision test: not built
        with pytest.raises(_lib.B2OError):
            qn._eltype_code(bad)
