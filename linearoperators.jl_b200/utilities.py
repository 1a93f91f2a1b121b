"""Consumers of the apply path -- mirror of the cheap checks and the norm estimator of src/utilities.jl:20-149.  They use nothing
but `mul!`/`*` and `dot`, so they run on any operator of this package (vectors: float64 CUDA tensors)."""
import numpy as np

from ._lib import LinearOperatorException
from .abstract import adjoint, apply, mul_, size
from .special_operators import _ctx_of

_EPS = float(np.finfo(np.float64).eps)


def _rand(ctx, n):
    import torch
    return torch.rand(int(n), dtype=torch.float64, device="cuda:%d" % ctx.device)


def check_ctranspose(op):
    """cheap check that op and op' are related (src/utilities.jl:66-84): |y'(A x) - (A'y)'x| small"""
    ctx = _ctx_of(op)
    m, n = size(op)
    x, y = _rand(ctx, n), _rand(ctx, m)
    yAx = ctx.dot(y, apply(op, x))
    xAty = ctx.dot(x, apply(adjoint(op), y))
    return abs(yAx - xAty) < (abs(yAx) + _EPS) * _EPS ** (1 / 3)


def check_hermitian(op):
    """cheap check that op is Hermitian (src/utilities.jl:91-116): v'A'Av == v'AAv"""
    ctx = _ctx_of(op)
    m, n = size(op)
    if m != n:
        raise LinearOperatorException("shape mismatch")
    v = _rand(ctx, n)
    w = apply(op, v).clone()
    s = ctx.dot(w, w)
    t = ctx.dot(v, apply(op, w))
    return abs(s - t) < (abs(s) + _EPS) * _EPS ** (1 / 3)


def check_positive_definite(op, semi=False):
    """cheap check that op is positive (semi-)definite (src/utilities.jl:123-149)"""
    ctx = _ctx_of(op)
    m, n = size(op)
    if m != n:
        raise LinearOperatorException("shape mismatch")
    v = _rand(ctx, n)
    vw = ctx.dot(v, apply(op, v))
    return vw >= 0 if semi else vw > 0


def normest(S, tol=-1, maxiter=100):
    """normest(S): estimate of the matrix 2-norm by power iteration on S'S (src/utilities.jl:20-59). Returns (e, cnt)."""
    import torch
    ctx = _ctx_of(S)
    m, n = size(S)
    cnt = 0
    if tol == -1:
        tol = _EPS
    v = torch.ones(m, dtype=torch.float64, device="cuda:%d" % ctx.device)
    v[torch.randn(m, device=v.device) < 0] = -1
    x = ctx.zeros(n)
    mul_(x, adjoint(S), v)
    e = float(np.sqrt(ctx.dot(x, x)))
    if e == 0:
        return e, cnt
    x /= e
    e_0 = 0.0
    Sx = ctx.zeros(m)
    while abs(e - e_0) > tol * e:
        e_0 = e
        mul_(Sx, S, x)
        if int(torch.count_nonzero(Sx)) == 0:
            Sx.normal_()
        mul_(x, adjoint(S), Sx)
        normx = float(np.sqrt(ctx.dot(x, x)))
        e = normx / float(np.sqrt(ctx.dot(Sx, Sx)))
        x /= normx
        cnt += 1
        if cnt > maxiter:
            break
    return e, cnt
