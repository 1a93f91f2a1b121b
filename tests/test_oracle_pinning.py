"""Pin the CPU oracle (oracle/b2o_oracle.c + oracle/oracle.py) against the reference's OWN test predicates
and literal known answers for the apply path.  The reference is Julia and cannot run here, so
these are transcriptions of what its test-suite asserts:

  test/test_linop.jl:437-461   Restriction/Extension, exact ==
  test/test_linop.jl:229-344   Eye / Ones / Zeros / Diagonal / rectangular Diagonal
  test/test_linop.jl:511-518   Householder
  test/test_lbfgs.jl:13-159    identity at start, rejection, H*B≈I, diag, reset!, dense BFGS, damped, norm bound
  test/test_lsr1.jl:7-72       same for L-SR1 against dense SR1
  test/test_kron.jl:3-39       kron operator against dense kron
"""
import json
import os

import numpy as np
import pytest

EPS = np.finfo(np.float64).eps
RTOL = np.sqrt(EPS)


def simple_vector(n):
    """test/test_aux.jl:35  T[-(-one(T))^i for i = 1:n] = [1, -1, 1, ...]"""
    return np.array([-((-1.0) ** i) for i in range(1, n + 1)])


# ---------------------------------------------------------------- index operators: exact equality
@pytest.mark.parametrize("idx", [[1, 2, 4, 7], list(range(3, 7)), list(range(1, 8, 2)), [4]])
def test_restriction_extension_exact(orc, idx):
    n = 10
    v = simple_vector(n)
    idx = np.array(idx)
    P, Z = orc.opRestriction(idx, n), orc.opExtension(idx, n)
    w = v[idx - 1]
    vz = np.zeros(n)
    vz[idx - 1] = v[idx - 1]
    assert np.array_equal(P(v), w)
    assert np.array_equal(P.T(w), vz)
    assert np.array_equal(Z(w), vz)
    assert np.array_equal(Z.T(v), w)
    assert np.array_equal((P * Z)(w), w)
    assert np.array_equal((Z * P)(v), vz)


def test_extension_duplicates_last_wins(orc):
    # Q4: res[I] = u with duplicate I keeps the last occurrence (src/special-operators.jl:173)
    res = np.full(5, 7.0)
    orc.extend_(res, [2, 4, 2, 2], np.array([1.0, 2.0, 3.0, 4.0]))
    assert np.array_equal(res, [0, 4.0, 0, 2.0, 0])


def test_restriction_ignores_alpha_beta(orc):
    # Q1 (src/special-operators.jl:167-169)
    P = orc.opRestriction([2, 3], 4)
    res = np.array([9.0, 9.0])
    P.mul(res, np.array([1.0, 2.0, 3.0, 4.0]), 5.0, 3.0)
    assert np.array_equal(res, [2.0, 3.0])


# ---------------------------------------------------------------- leaf operators
def test_eye_and_rect_eye_quirk(orc):
    v = simple_vector(5)
    assert np.array_equal(orc.opEye(5)(v), v)
    E = orc.opEye(7, 5)
    res = np.full(7, 3.0)
    E.mul(res, v, 2.0, 0.0)
    assert np.array_equal(res, np.r_[2 * v, 0, 0])
    res = np.full(7, 3.0)
    E.mul(res, v, 2.0, 0.5)                     # Q2: tail becomes β, not β*res
    assert np.array_equal(res, np.r_[2 * v + 1.5, 0.5, 0.5])


def test_ones_zeros(orc):
    u = simple_vector(6) * 1.5
    assert np.allclose(orc.opOnes(4, 6)(u), np.sum(u) * np.ones(4), rtol=0, atol=RTOL * np.linalg.norm(u))
    assert np.linalg.norm(orc.opZeros(4, 6)(u)) <= EPS
    res = np.full(4, 2.0)
    orc.opZeros(4, 6).mul(res, u, 1.0, 3.0)
    assert np.array_equal(res, np.full(4, 6.0))


def test_diagonal_known_answers(orc):
    # test/test_linop.jl:308-319 (real restriction of the complex test) + the literal `[2;1]` check (:565,:594)
    v, u = simple_vector(8) * 2, simple_vector(8) + 0.5
    D = orc.opDiagonal(v)
    assert np.linalg.norm(D(u) - v * u) <= EPS * np.linalg.norm(u)
    res = simple_vector(8).copy()
    res2 = v * u * 2.0 + 2.0 * res
    D.mul(res, u, 2.0, 2.0)
    assert np.linalg.norm(res - res2) <= EPS * np.linalg.norm(u)
    assert np.array_equal(orc.opDiagonal(np.array([2.0, 1.0]))(np.ones(2)), [2.0, 1.0])


def test_rect_diagonal(orc):
    # test/test_linop.jl:321-344; Q3: tail rows are zeroed even when β != 0
    nmin, nmax = 4, 7
    v, u = simple_vector(nmin) * 3, simple_vector(nmin)
    A = np.zeros((nmax, nmin))
    A[np.arange(nmin), np.arange(nmin)] = v
    D = orc.opDiagonal(v, nmax, nmin)
    assert np.linalg.norm(A @ u - D(u)) <= EPS * np.linalg.norm(u)
    w = simple_vector(nmax)
    assert np.linalg.norm(A.T @ w - D.T(w)) <= EPS * np.linalg.norm(w)
    res = np.full(nmax, 5.0)
    D.mul(res, u, 1.0, 1.0)
    assert np.array_equal(res[nmin:], np.zeros(nmax - nmin))


def test_householder(orc):
    # test/test_linop.jl:511-518
    n = 9
    h = simple_vector(n) / 3.0
    u = np.arange(1.0, n + 1)
    H = orc.opHouseholder(h)
    assert np.linalg.norm(H(u) - (u - 2 * np.dot(h, u) * h)) <= RTOL * np.linalg.norm(u)
    assert np.linalg.norm(H.T(u) - H(u)) == 0


def test_composition_against_dense(orc):
    # test/test_linop.jl:139-226: (A+B), (A*B), scalar, unary minus against dense algebra
    rng = np.random.default_rng(0)
    d1, d2, h = rng.random(6), rng.random(6) + 0.5, rng.random(6)
    h /= np.linalg.norm(h)
    D1, D2, H = orc.opDiagonal(d1), orc.opDiagonal(d2), orc.opHouseholder(h)
    Hm = np.eye(6) - 2 * np.outer(h, h)
    for op, M in [(D1 + D2, np.diag(d1 + d2)), (H * D1, Hm @ np.diag(d1)), (2.5 * H, 2.5 * Hm), (-(H * D2), -Hm @ np.diag(d2)),
                  (H * D1 + 0.1 * orc.opEye(6), Hm @ np.diag(d1) + 0.1 * np.eye(6)), ((H * D1).T, np.diag(d1) @ Hm)]:
        assert np.linalg.norm(op.matrix() - M) <= RTOL * np.linalg.norm(M)


# ---------------------------------------------------------------- L-BFGS (test/test_lbfgs.jl)
def bfgs_dense(B, s, y, damped=False):
    ys = y @ s
    Bs = B @ s
    tol = 0.2 * (s @ Bs) if damped else 1e-20
    if ys > tol:
        B = B - np.outer(Bs, Bs) / (s @ Bs) + np.outer(y, y) / ys
    return B


def test_lbfgs_reference_predicates(orc):
    n, mem = 10, 5
    B = orc.LBFGS(n, mem=mem, scaling=False)
    H = orc.LBFGS(n, mem=mem, scaling=False, inverse=True)
    for _ in range(2):                                                # "Run again after reset!"
        assert np.linalg.norm(B.diag() - np.diag(B.matrix())) <= RTOL
        assert B.insert == 1 and H.insert == 1
        assert np.linalg.norm(B.matrix() - np.eye(n)) <= EPS
        assert np.linalg.norm(H.matrix() - np.eye(n)) <= EPS
        s, z = simple_vector(n), np.zeros(n)
        for op in (B, H):                                             # nonpositive curvature is rejected
            assert not op.push(s, -s) and op.insert == 1
            assert not op.push(s, z) and op.insert == 1
        insert = 0
        for i in range(1, mem + 3):
            s = np.ones(n) * i
            y = np.r_[i, np.ones(n - 1)]
            if s @ y > 1e-20:
                insert += 1
                B.push(s, y)
                H.push(s, y)
        assert B.insert == insert % mem + 1 and H.insert == insert % mem + 1
        Bm, Hm = B.matrix(), H.matrix()
        assert np.all(np.linalg.eigvalsh((Bm + Bm.T) / 2) > 0) and np.all(np.linalg.eigvalsh((Hm + Hm.T) / 2) > 0)
        assert np.linalg.norm(Bm - Bm.T) <= RTOL * np.linalg.norm(Bm)
        assert np.linalg.norm(B.diag() - np.diag(Bm)) <= RTOL
        assert np.linalg.norm(Hm @ Bm - np.eye(n)) <= RTOL            # Matrix(H*B) ≈ I
        v = simple_vector(n)
        assert np.linalg.norm(B.apply(v) - v) > RTOL
        assert np.linalg.norm(np.linalg.norm(Bm, 2)) <= B.opnorm_upper_bound
        B.reset()
        H.reset()
        assert B.scaling_factor == 1.0 and H.scaling_factor == 1.0
        assert np.linalg.norm(B.apply(v) - v) < RTOL and np.linalg.norm(H.apply(v) - v) < RTOL


@pytest.mark.parametrize("damped", [False, True])
def test_lbfgs_equals_dense_bfgs(orc, damped):
    n = mem = 10
    LB = orc.LBFGS(n, mem=mem, scaling=False, damped=damped)
    B = np.eye(n)
    rng = np.random.default_rng(3)
    assert np.linalg.norm(LB.matrix() - B) < RTOL * np.linalg.norm(B)
    for k in range(mem):
        s = simple_vector(n) if k == 0 else rng.random(n)
        y = simple_vector(n) if k == 0 else s + 0.1 * rng.random(n)
        B = bfgs_dense(B, s, y, damped)
        LB.push(s, y)
        assert np.linalg.norm(LB.matrix() - B) < RTOL * np.linalg.norm(B)
        assert np.linalg.norm(LB.diag() - np.diag(B)) < RTOL * np.linalg.norm(np.diag(B))
    assert np.linalg.norm(B, 2) <= LB.opnorm_upper_bound * (1 + 1e-12)


def test_lbfgs_damped_pair(orc):
    # test/test_lbfgs.jl:104-137
    n, mem = 10, 5
    B = orc.LBFGS(n, mem=mem, damped=True, scaling=False, sigma2=0.8, sigma3=np.inf)
    H = orc.LBFGS(n, mem=mem, damped=True, scaling=False, sigma2=0.8, sigma3=np.inf, inverse=True)
    ins = 0
    for i in range(1, mem + 3):
        y = simple_vector(n)
        g = simple_vector(n)
        d = -H.apply(g)
        a = i / mem
        s = a * d
        if y @ simple_vector(n) > 0.2 * (s @ B.apply(s)):
            ins += 1
            B.push(s, y)
            H.push_damped_inv(s, y.copy(), a, g)
    assert B.insert == ins % mem + 1 and H.insert == ins % mem + 1
    Bm, Hm = B.matrix(), H.matrix()
    assert np.linalg.norm(Hm @ Bm - np.eye(n)) <= RTOL
    assert np.linalg.norm(B.diag() - np.diag(Bm)) <= RTOL
    assert np.linalg.norm(Bm, 2) <= B.opnorm_upper_bound


def test_lbfgs_scaling_and_5arg(orc):
    n, mem = 12, 4
    rng = np.random.default_rng(5)
    for inverse in (False, True):
        op = orc.LBFGS(n, mem=mem, scaling=True, inverse=inverse)
        for _ in range(6):
            s = rng.random(n)
            op.push(s, s + 0.1 * rng.random(n))
        M = op.matrix()
        x, r0 = rng.random(n), rng.random(n)
        res = r0.copy()
        op.apply(x, 2.0, -0.5, res=res)
        assert np.linalg.norm(res - (2.0 * M @ x - 0.5 * r0)) <= 1e-13 * np.linalg.norm(res)


def test_solve_shifted_system_reference_predicates(orc):
    """test/test_solve_shifted_system.jl:5-63 and the docstring check of src/utilities.jl:191"""
    rng = np.random.default_rng(9)
    for scaling, sigma in [(False, 0.1), (True, 0.1), (True, 0.0)]:
        n, M = 100, 5
        B = orc.LBFGS(n, mem=M, scaling=scaling)
        H = orc.LBFGS(n, mem=M, scaling=scaling, inverse=True)
        for _ in range(10):
            s, y = rng.random(n), rng.random(n)
            B.push(s, y)
            H.push(s, y)
        x_true = rng.standard_normal(n)
        b = B.apply(x_true) + sigma * x_true
        x = B.solve_shifted(b, sigma)
        assert np.all(np.isfinite(x))
        assert np.allclose(x, x_true, atol=1e-6, rtol=1e-6)
        assert np.linalg.norm(B.apply(x) + sigma * x - b) / np.linalg.norm(b) < 1e-8
        if sigma == 0.0:
            assert np.allclose(x, H.apply(b), atol=1e-6, rtol=1e-6)            # ldiv! agrees with the inverse operator
    with pytest.raises(ValueError):
        B.solve_shifted(b, -0.1)


# ---------------------------------------------------------------- L-SR1 (test/test_lsr1.jl)
def sr1_dense(B, s, y):
    r = y - B @ s
    den = r @ s
    if abs(den) >= 1e-8 + 1e-8 * np.linalg.norm(s) * np.linalg.norm(r):
        B = B + np.outer(r, r) / den
    return B


def test_lsr1_reference_predicates(orc):
    n, mem = 10, 5
    B = orc.LSR1(n, mem=mem, scaling=False)
    for _ in range(2):
        assert np.linalg.norm(B.diag() - np.diag(B.matrix())) <= RTOL
        assert B.insert == 1
        assert np.linalg.norm(B.matrix() - np.eye(n)) <= EPS
        s = simple_vector(n)
        assert not B.push(s, B.apply(s)) and B.insert == 1            # y = B s is rejected
        for i in range(1, mem + 3):
            B.push(np.ones(n) * i, np.r_[i, np.ones(n - 1)])
        Bm = B.matrix()
        assert np.linalg.norm(Bm - Bm.T) <= RTOL * np.linalg.norm(Bm)
        assert np.linalg.norm(B.diag() - np.diag(Bm)) <= RTOL
        v = simple_vector(n)
        assert np.linalg.norm(B.apply(v) - v) > RTOL
        B.reset()
        assert B.scaling_factor == 1.0
        assert np.linalg.norm(B.apply(v) - v) < RTOL
        assert np.linalg.norm(B.matrix(), 2) <= B.opnorm_upper_bound


def test_lsr1_equals_dense_sr1(orc):
    n = mem = 10
    LB = orc.LSR1(n, mem=mem, scaling=False)
    B = np.eye(n)
    rng = np.random.default_rng(11)
    for k in range(mem):
        s = simple_vector(n) if k == 0 else rng.random(n) - 0.5
        y = simple_vector(n) if k == 0 else rng.random(n) - 0.5
        B = sr1_dense(B, s, y)
        LB.push(s, y)
        assert np.linalg.norm(LB.matrix() - B) < RTOL * np.linalg.norm(B)
        assert np.linalg.norm(LB.diag() - np.diag(B)) < RTOL * np.linalg.norm(np.diag(B))
    assert np.linalg.norm(B, 2) <= LB.opnorm_upper_bound * (1 + 1e-12)


# ---------------------------------------------------------------- diagonal quasi-Newton (test/test_diag.jl:37-106): LITERAL known answers
X0 = np.array([-1.0, 1.0, -1.0])
X1 = X0 + np.array([1.0, 0.0, 1.0])
GRADS = {
    "f": lambda x: 2 * np.array([x[0], x[1], x[2]]),
    "g": lambda x: np.array([np.exp(x[0]), 1.0, -np.sin(x[2])]),
    "h": lambda x: np.array([2 * x[0] * x[1] * x[2]**3, x[0]**2 * x[2]**3, 3 * x[0]**2 * x[1] * x[2]**2]),
}
# Bref of test/test_diag.jl:76-88 (kind 0 = DiagonalPSB, 1 = DiagonalAndrei) and Bref_spg (:85-88)
BREF = {
    ("f", 0): [2, -1, 2], ("f", 1): [2, -2, 2],
    ("g", 0): [1 + (np.sin(-1) - np.exp(-1) - 1) / 2, -1, 1 + (np.sin(-1) - np.exp(-1) - 1) / 2],
    ("g", 1): [(1 + np.sin(-1) - np.exp(-1)) / 2, -2, (1 + np.sin(-1) - np.exp(-1)) / 2],
    ("h", 0): [-5 / 2, -1, -5 / 2], ("h", 1): [-5 / 2, -2, -5 / 2],
}
BREF_SPG = {"f": 2.0, "g": (1 - np.exp(-1) + np.sin(-1)) / 2, "h": -5 / 2}


@pytest.mark.parametrize("fun", ["f", "g", "h"])
def test_diagonal_qn_hard_coded_reference_values(orc, fun):
    s, y = X1 - X0, GRADS[fun](X1) - GRADS[fun](X0)
    for kind in (0, 1):
        d = np.array([1.0, -1.0, 1.0])
        orc.diagqn_push(kind, d, s, y)
        assert np.linalg.norm(d - np.array(BREF[(fun, kind)], dtype=float)) <= 1e-10
        assert abs(s @ (d * s) - s @ y) <= 1e-10                                # weak secant equation (:50-72)
    sig = np.array([1.0])
    orc.diagqn_push(3, sig, s, y)
    assert abs(sig[0] - BREF_SPG[fun]) <= 1e-10
    with pytest.raises(ZeroDivisionError):
        orc.diagqn_push(0, np.ones(3), np.zeros(3), y)


# ---------------------------------------------------------------- kron (test/test_kron.jl:3-39)
@pytest.mark.parametrize("shapeA,shapeB", [((2, 3), (2, 3)), ((4, 4), (3, 5))])
def test_kron_against_dense(orc, shapeA, shapeB):
    rng = np.random.default_rng(7)
    A, B = rng.random(shapeA), rng.random(shapeB)
    K = np.kron(A, B)
    normK = np.abs(K).sum(axis=0).max()
    x = simple_vector(K.shape[1])
    res = np.empty(K.shape[0])
    orc.kron_(res, A, B, x)
    assert np.abs(K @ x - res).sum() < 1e-12 * normK
    xt = simple_vector(K.shape[0])
    rt = np.empty(K.shape[1])
    orc.kron_(rt, A, B, xt, trans=1)
    assert np.abs(K.T @ xt - rt).sum() < 1e-12 * normK
    r2 = np.ones(K.shape[0])
    orc.kron_(r2, A, B, x, alpha=2.0, beta=-1.0)
    assert np.abs(2 * K @ x - 1 - r2).sum() < 1e-12 * normK


# ---------------------------------------------------------------- frozen golden vectors
def test_golden_vectors(orc):
    """tests/golden/golden_v1.json was produced by tests/golden/make_golden.py from this oracle after it passed
    every predicate above; it freezes seeded inputs -> outputs so that later edits cannot drift silently."""
    path = os.path.join(os.path.dirname(__file__), "golden", "golden_v1.json")
    G = json.load(open(path))
    from golden.make_golden import compute_cases
    fresh = compute_cases(orc)
    assert set(fresh) == set(G["cases"])
    for name, val in fresh.items():
        ref = np.array(G["cases"][name])
        assert np.allclose(val, ref, rtol=1e-14, atol=0), name


# ---------------------------------------------------------------- LinearOperator(Matrix) -- test/test_linop.jl:41-75, :101-124
def test_dense_matrix_operator_predicates(orc):
    """oracle gemv (the closures of LinearOperator(M), src/constructors.jl:25-27) against dense algebra: A*v, transpose(A)*u,
    A'*u, Matrix(op) == A (column by column with unit vectors, src/abstract.jl:282-292), matrix right-hand sides hcat(v, -2v),
    5-arg α/β form, β = 0 never reads res; Float32 storage as in test/gpu/nvidia.jl:8-15."""
    nrow, ncol = 10, 6                                                # test/test_linop.jl:2
    rng = np.random.default_rng(11)
    for dt, rtol in ((np.float64, 1e-15), (np.float32, 1e-6)):
        A = np.asfortranarray(rng.uniform(-1, 1, (nrow, ncol)).astype(dt))
        v, u = rng.uniform(-1, 1, ncol).astype(dt), rng.uniform(-1, 1, nrow).astype(dt)
        res = np.full(nrow, np.nan, dtype=dt)
        orc.gemv_(res, A, v)
        assert np.linalg.norm(res - A @ v) <= 10 * rtol * np.linalg.norm(v) * np.sqrt(ncol)
        rt = np.full(ncol, np.nan, dtype=dt)
        orc.gemv_(rt, A, u, trans=1)
        assert np.linalg.norm(rt - A.T @ u) <= 10 * rtol * np.linalg.norm(u) * np.sqrt(nrow)
        full = np.empty((nrow, ncol), dtype=dt)                       # Matrix(op): op * e_i
        for i in range(ncol):
            e = np.zeros(ncol, dtype=dt)
            e[i] = 1
            col = np.empty(nrow, dtype=dt)
            orc.gemv_(col, A, e)
            full[:, i] = col
        assert np.array_equal(full, A)                                # norm(A - Matrix(op)) <= ϵ * norm(A), exactly 0 here
        mv = np.stack([v, -2 * v], axis=1)                            # mul!(res_mat, op, hcat(v, -2v))
        out = np.empty((nrow, 2), dtype=dt)
        for j in range(2):
            c = np.empty(nrow, dtype=dt)
            orc.gemv_(c, A, np.ascontiguousarray(mv[:, j]))
            out[:, j] = c
        assert np.linalg.norm(out - A @ mv) <= 10 * rtol * np.linalg.norm(mv) * np.sqrt(ncol)
        r0 = rng.uniform(-1, 1, nrow).astype(dt)
        r5 = r0.copy()
        orc.gemv_(r5, A, v, 2.0, -0.5)
        assert np.linalg.norm(r5 - (2 * (A @ v) - 0.5 * r0)) <= 20 * rtol * (np.linalg.norm(v) * np.sqrt(ncol) + np.linalg.norm(r0))
        sub = A[2:9, 1:5]                                             # a view with leading dimension > nrow is passed column-major
        rs = np.empty(7, dtype=dt)
        orc.gemv_(rs, sub, v[1:5])
        assert np.linalg.norm(rs - sub @ v[1:5]) <= 10 * rtol * np.linalg.norm(v) * 2


# ---------------------------------------------------------------- LinearOperator(SparseMatrixCSC) -- test/test_linop.jl:740-766
def test_sparse_matrix_operator_predicates(orc):
    """oracle spmv (SparseArrays' mul! restated: the closures of LinearOperator(M::SparseMatrixCSC), src/constructors.jl:25-27)
    against dense algebra and scipy.sparse (independent): `opA * b == A * b`, transpose, adjoint (test_linop.jl:759-766, exact
    on integer-valued data), Matrix(op) == A column by column, 5-arg α/β form, β = 0 never reads res, empty columns/rows,
    the block-diagonal predicate of :740-755 with a sprand(2, 4, 0.5) block."""
    import scipy.sparse as sp
    rng = np.random.default_rng(21)
    for dt, rtol in ((np.float64, 1e-15), (np.float32, 1e-6)):
        for (m, n, dens) in ((10, 10, 0.2), (2, 4, 0.5), (40, 25, 0.1), (7, 300, 0.3), (300, 7, 0.3), (5, 5, 0.0)):
            A = sp.random(m, n, density=dens, random_state=rng, format="csc").astype(dt)
            A.sort_indices()
            cp, rv, nz = A.indptr.astype(np.int64) + 1, A.indices.astype(np.int64) + 1, A.data
            Ad = A.toarray().astype(np.float64)
            v, u = rng.uniform(-1, 1, n).astype(dt), rng.uniform(-1, 1, m).astype(dt)
            res = np.full(m, np.nan, dtype=dt)
            orc.spmv_csc_(res, m, n, cp, rv, nz, v)
            assert np.linalg.norm(res - Ad @ v) <= 10 * rtol * max(np.linalg.norm(Ad), 1.0) * np.linalg.norm(v)
            assert np.linalg.norm(res - A.astype(np.float64) @ v.astype(np.float64)) <= 10 * rtol * max(np.linalg.norm(Ad), 1.0) * np.linalg.norm(v)
            rt = np.full(n, np.nan, dtype=dt)
            orc.spmv_csc_(rt, m, n, cp, rv, nz, u, trans=1)
            assert np.linalg.norm(rt - Ad.T @ u) <= 10 * rtol * max(np.linalg.norm(Ad), 1.0) * np.linalg.norm(u)
            full = np.empty((m, n), dtype=dt)                      # Matrix(op): op * e_i, exact
            for i in range(n):
                e = np.zeros(n, dtype=dt)
                e[i] = 1
                col = np.empty(m, dtype=dt)
                orc.spmv_csc_(col, m, n, cp, rv, nz, e)
                full[:, i] = col
            assert np.array_equal(full, A.toarray())
            r0 = rng.uniform(-1, 1, m).astype(dt)
            r5 = r0.copy()
            orc.spmv_csc_(r5, m, n, cp, rv, nz, v, 2.0, -0.5)
            assert np.linalg.norm(r5 - (2 * (Ad @ v) - 0.5 * r0)) <= 20 * rtol * (max(np.linalg.norm(Ad), 1.0) * np.linalg.norm(v) + np.linalg.norm(r0))
            t0 = rng.uniform(-1, 1, n).astype(dt)
            t5 = t0.copy()
            orc.spmv_csc_(t5, m, n, cp, rv, nz, u, 2.0, -0.5, trans=1)
            assert np.linalg.norm(t5 - (2 * (Ad.T @ u) - 0.5 * t0)) <= 20 * rtol * (max(np.linalg.norm(Ad), 1.0) * np.linalg.norm(u) + np.linalg.norm(t0))
    # issue #139 (test_linop.jl:759-766) asserts ==; with integer-valued entries every sum is exact in any order
    A = sp.random(10, 10, density=0.2, random_state=rng, format="csc", data_rvs=lambda k: rng.integers(-8, 9, k).astype(np.float64))
    b = np.where(rng.uniform(size=10) < 0.2, rng.integers(-8, 9, 10), 0).astype(np.float64)     # sprand(10, 0.2) as a dense vector
    cp, rv = A.indptr.astype(np.int64) + 1, A.indices.astype(np.int64) + 1
    for trans, ref in ((0, A @ b), (1, A.T @ b)):
        out = np.empty(10)
        orc.spmv_csc_(out, 10, 10, cp, rv, A.data, b, trans=trans)
        assert np.array_equal(out, ref)


def test_golden_vectors_v2_matrix_leaves_and_quirks(orc):
    """tests/golden/golden_v2.json (make_golden.compute_cases_v2): dense / sparse LinearOperator(M) products in Float64 and
    Float32, index operators and the α/β quirks Q1-Q4, frozen after the predicates above passed"""
    path = os.path.join(os.path.dirname(__file__), "golden", "golden_v2.json")
    G = json.load(open(path))["cases"]
    from golden.make_golden import compute_cases_v2
    fresh = compute_cases_v2(orc)
    assert set(fresh) == set(G)
    for name, val in fresh.items():
        exact = name.startswith(("restrict", "extend", "eye", "diag"))
        assert (np.array_equal(val, G[name]) if exact else np.allclose(val, G[name], rtol=1e-14, atol=0)), name
    # independent of the oracle: the frozen dense / sparse products against plain numpy on the same seeded inputs
    from golden.make_golden import sparse_pattern
    A = (2.0 * orc.uniform(37 * 23, 11) - 1.0).reshape(23, 37).T
    assert np.allclose(G["dense_f64_N"], A @ orc.uniform(23, 12), rtol=1e-13)
    assert np.allclose(G["dense_f64_T"], A.T @ orc.uniform(37, 13), rtol=1e-13)
    cp, rv, nz = sparse_pattern(orc, 60, 45, 0.2, 21)
    S = np.zeros((60, 45))
    for j in range(45):
        for k in range(cp[j] - 1, cp[j + 1] - 1):
            S[rv[k] - 1, j] = nz[k]
    assert 400 < len(nz) < 700
    assert np.allclose(G["sparse_f64_N"], S @ orc.uniform(45, 23), rtol=1e-12, atol=1e-15)
    assert np.allclose(G["sparse_f64_T"], S.T @ orc.uniform(60, 24), rtol=1e-12, atol=1e-15)


def test_external_column_oracle_is_the_resident_one():
    """full-size parity runs stream the state columns from the GPU: same apply code, same bits, same inner products"""
    import oracle as orc
    orc.set_mode(True, 1)
    n, mem = 1003, 4
    for inverse in (False, True):
        o = orc.LBFGS(n, mem=mem, inverse=inverse)
        for i in range(6):
            s = orc.uniform(n, 100 + i)
            o.push(s, s + 0.1 * orc.uniform(n, 200 + i))
        x = orc.uniform(n, 7)
        ref = o.apply(x)
        bufs = [np.empty(n), np.empty(n)]
        calls = []

        def fetch(which, k0, slot):
            calls.append((which, k0, slot))
            bufs[slot][:] = o.col("syab"[which], k0)
            return bufs[slot].ctypes.data

        e = orc.ExternalLBFGS(n, mem, inverse, fetch)
        e.ys[:] = o.ys
        e.set_state(o.insert, o.scaling_factor)
        assert np.array_equal(e.apply(x), ref)
        d = e.last_dots()
        assert len(d) == 2 * mem and len(calls) == 4 * mem if inverse else len(calls) == 2 * mem
        if not inverse:      # a_k.x, b_k.x oldest -> newest
            k = (o.insert - 1) % mem
            assert d[0] == orc.dot(o.col("a", k), x) and d[1] == orc.dot(o.col("b", k), x)


# ---------------------------------------------------------------- Float32 quasi-Newton restatement (oracle/oracle_f32.py)
def _f32_pairs(n, k, lsr1=False):
    rng = np.random.default_rng(5)
    out = []
    for _ in range(k):
        s = rng.random(n).astype(np.float32)
        y = (s + 0.1 * rng.random(n)).astype(np.float32) if not lsr1 else (rng.random(n) * 1.5 - 0.5).astype(np.float32)
        out.append((s, y))
    return out


def test_f32_precision_predicates_of_the_reference():
    """test/test_lbfgs.jl:162-178, test/test_lsr1.jl:74-86 for T = Float32: s = y = ones, v = (-1)^i; eltype(B*v) == T, and
    the operator is the identity before the first push (test/test_lbfgs.jl:21-27)."""
    import oracle_f32 as o32
    n, mem = 10, 5
    v = np.array([-(-1.0) ** i for i in range(1, n + 1)], dtype=np.float32)
    for op in (o32.LBFGS32(n, mem), o32.LBFGS32(n, mem, inverse=True), o32.LSR1_32(n, mem)):
        assert np.array_equal(op.apply(v), v)
        op.push(np.ones(n, np.float32), np.ones(n, np.float32))
        assert op.apply(v).dtype == np.float32


def test_f32_secant_equation_and_inverse_pair():
    """B_{k+1} s_k = y_k, H_{k+1} y_k = s_k and H·B = I (test/test_lbfgs.jl:45-52 predicate) hold to Float32 accuracy"""
    import oracle_f32 as o32
    n, mem = 40, 5
    B, H, L = o32.LBFGS32(n, mem), o32.LBFGS32(n, mem, inverse=True), o32.LSR1_32(n, mem)
    for (s, y), (sl, yl) in zip(_f32_pairs(n, 8), _f32_pairs(n, 8, lsr1=True)):
        assert B.push(s, y) and H.push(s, y)
        L.push(sl, yl)
        assert np.linalg.norm(B.apply(s) - y) <= 2e-5 * np.linalg.norm(y)
        assert np.linalg.norm(H.apply(y) - s) <= 2e-5 * np.linalg.norm(s)
    x = np.random.default_rng(6).random(n).astype(np.float32)
    assert np.linalg.norm(H.apply(B.apply(x)) - x) <= 1e-4 * np.linalg.norm(x)
    assert not B.push(x, -x)                                                 # non-positive curvature is rejected  :281


def test_f32_restatement_agrees_with_the_float64_oracle():
    import oracle as orc
    import oracle_f32 as o32
    orc.build()
    n, mem = 200, 4
    for inverse in (False, True):
        a, b = o32.LBFGS32(n, mem, inverse=inverse), orc.LBFGS(n, mem=mem, inverse=inverse)
        for s, y in _f32_pairs(n, 6):
            a.push(s, y)
            b.push(s.astype(np.float64), y.astype(np.float64))
        x = np.random.default_rng(7).random(n).astype(np.float32)
        r0 = np.random.default_rng(8).random(n).astype(np.float32)
        ref = r0.astype(np.float64)
        b.apply(x.astype(np.float64), -0.5, 2.0, res=ref)
        got = a.apply(x, -0.5, 2.0, res=r0)
        assert np.linalg.norm(got - ref) <= 2e-5 * np.linalg.norm(ref)
    a, b = o32.LSR1_32(n, mem), orc.LSR1(n, mem=mem)
    for s, y in _f32_pairs(n, 6, lsr1=True):
        a.push(s, y)
        b.push(s.astype(np.float64), y.astype(np.float64))
    x = np.random.default_rng(7).random(n).astype(np.float32)
    ref = np.empty(n)
    b.apply(x.astype(np.float64), 1.0, 0.0, res=ref)
    assert np.linalg.norm(a.apply(x) - ref) <= 1e-3 * np.linalg.norm(ref)     # SR1 recurrences amplify Float32 rounding


def test_golden_vectors_v3_float32_quasi_newton(orc):
    """tests/golden/golden_v3.json (make_golden.compute_cases_v3): the numpy Float32 restatement reproduces its frozen outputs bit for
    bit (regression pin of oracle/oracle_f32.py; the GPU test compares the CUDA path with the same file)"""
    import json
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_v3.json")))["cases"]
    now = make_golden.compute_cases_v3(orc)
    assert set(now) == set(G)
    for k in G:
        assert np.array_equal(np.asarray(now[k]), np.asarray(G[k])), k


def test_compact_forward_form_accuracy_study():
    """tools/compact_accuracy_study.py (profiles/r2_compact_accuracy.md): over every family of (s, y) pairs the compact forward form stays
    within a small factor of the reference recursion's own rounding error, measured against a long-double ground truth -- the written
    basis for offering LBFGSOperator(n, compact=True)"""
    import importlib.util
    import io
    import json
    from contextlib import redirect_stdout
    path = os.path.join(os.path.dirname(__file__), "..", "tools", "compact_accuracy_study.py")
    spec = importlib.util.spec_from_file_location("compact_accuracy_study", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    buf = io.StringIO()
    with redirect_stdout(buf):
        mod.main()
    rows = [json.loads(l) for l in buf.getvalue().splitlines() if l.strip()]
    assert len(rows) >= 10
    for r in rows:
        assert r["rel_err_compact_form"] <= max(20.0 * r["rel_err_reference_form"], 5e-15), r
        assert r["rel_diff_compact_vs_reference"] <= max(1e-12, 5.0 * r["rel_err_reference_form"]), r
