"""Drop-in check: a Krylov solver written ONLY against the operator interface the reference exposes (mul!, size, transpose) --
what Krylov.jl's cg/minres use -- runs unchanged on the CUDA operators."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def cg(lo, A, b, ctx, tol=1e-12, maxit=200):
    """conjugate gradients using only mul!(res, A, v) and dot; x0 = 0"""
    n = lo.size(A, 1)
    x = ctx.zeros(n)
    r = b.clone()
    p = r.clone()
    Ap = ctx.empty(n)
    rs = ctx.dot(r, r)
    b2 = rs
    for it in range(maxit):
        lo.mul_(Ap, A, p)
        alpha = rs / ctx.dot(p, Ap)
        x.add_(p, alpha=alpha)
        r.add_(Ap, alpha=-alpha)
        rs_new = ctx.dot(r, r)
        if rs_new <= tol * tol * b2:
            return x, it + 1
        p.mul_(rs_new / rs).add_(r)
        rs = rs_new
    return x, maxit


def test_cg_on_shifted_lbfgs_matches_solve_shifted_system(lo, ctx):
    n, mem, sigma = 200003, 6, 0.5
    B = lo.LBFGSOperator(n, mem=mem, ctx=ctx)
    for i in range(9):
        s = ctx.uniform(n, 100 + i)
        y = s + 0.1 * ctx.uniform(n, 200 + i)
        lo.push_(B, s, y)
    A = lo.ShiftedOperator(B, sigma)                       # B + σI: SPD, mem+1 distinct eigenvalue clusters -> CG converges fast
    b = ctx.uniform(n, 51, -1.0, 1.0)
    x_cg, its = cg(lo, A, b, ctx)
    assert its <= 4 * mem + 10
    x_direct = lo.solve_shifted_system_(ctx.zeros(n), B, b, sigma)
    d = x_cg - x_direct
    assert np.sqrt(ctx.dot(d, d) / ctx.dot(x_direct, x_direct)) <= 1e-9
    res = A * x_cg - b
    assert np.sqrt(ctx.dot(res, res) / ctx.dot(b, b)) <= 1e-10
    assert lo.nprod(A) == its + 1 and lo.nprod(B) >= its    # counters tick like the reference's (src/operations.jl:25)


def test_cg_on_fused_spd_chain(lo, ctx):
    """CG on a fused static tree: D + H*D2*H (SPD) evaluated in one launch per iteration"""
    n = 100001
    h = ctx.uniform(n, 3)
    h /= float(np.sqrt(ctx.dot(h, h)))
    d1, d2 = ctx.uniform(n, 4, 1.0, 2.0), ctx.uniform(n, 5, 0.5, 1.5)
    H = lo.opHouseholder(h)
    tree = lo.opDiagonal(d1) + H * lo.opDiagonal(d2) * H
    A = lo.fuse(tree)
    b = ctx.uniform(n, 6, -1.0, 1.0)
    l0 = ctx.launch_count()
    x, its = cg(lo, A, b, ctx, tol=1e-11, maxit=300)
    assert its < 300
    res = tree * x - b
    assert np.sqrt(ctx.dot(res, res) / ctx.dot(b, b)) <= 1e-9
    assert lo.nprod(A) == its


def test_reference_check_predicates_and_normest(lo, ctx):
    """check_hermitian / check_positive_definite / check_ctranspose / normest (src/utilities.jl:20-149) -- the predicates the
    reference's own test-suite applies to its operators (test/test_lbfgs.jl:48-52) -- on the CUDA operators"""
    n = 50001
    B = lo.LBFGSOperator(n, mem=5, ctx=ctx)
    H = lo.InverseLBFGSOperator(n, mem=5, ctx=ctx)
    for i in range(7):
        s = ctx.uniform(n, 100 + i)
        y = s + 0.1 * ctx.uniform(n, 200 + i)
        lo.push_(B, s, y)
        lo.push_(H, s, y)
    d = ctx.uniform(n, 1, 0.5, 2.0)
    D = lo.opDiagonal(d)
    h = ctx.uniform(n, 3)
    h /= float(np.sqrt(ctx.dot(h, h)))
    P = lo.opRestriction(np.arange(1, n + 1, 3), n)
    for op in (B, H, D, lo.opHouseholder(h) * D * lo.opHouseholder(h)):
        assert lo.check_hermitian(op) and lo.check_ctranspose(op)
    assert lo.check_positive_definite(B) and lo.check_positive_definite(H) and lo.check_positive_definite(D)
    assert lo.check_ctranspose(P) and lo.check_ctranspose(lo.opHouseholder(h) * D)
    assert not lo.check_positive_definite(-D)
    e, cnt = lo.normest(D)
    assert abs(e - float(d.max())) <= 1e-6 * float(d.max()) or cnt > 100
    e, _ = lo.normest(lo.opHouseholder(h))
    assert abs(e - 1.0) <= 1e-8
    with pytest.raises(lo.LinearOperatorException):
        lo.check_hermitian(P)
