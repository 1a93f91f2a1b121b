"""hcat / vcat / hvcat of operators -- mirror of src/cat.jl.  Blocks work on views of the caller's vectors."""
from ._lib import LinearOperatorException
from .abstract import (LinearOperator, _as_op, adjoint, eltype, mul_, promote_storage, size, storage_type, transpose)


def _hcat_prod_(res, A, B, Ancol, nV, v, alpha, beta):
    """src/cat.jl:7-19"""
    mul_(res, A, v[:Ancol], alpha, beta)
    mul_(res, B, v[Ancol:nV], alpha, 1.0)


def _hcat_ctprod_(res, A, B, Ancol, nV, u, alpha, beta):
    """src/cat.jl:21-33"""
    mul_(res[:Ancol], A, u, alpha, beta)
    mul_(res[Ancol:nV], B, u, alpha, beta)


def _hcat2(A, B):
    """src/cat.jl:35-51"""
    if size(A, 1) != size(B, 1):
        raise LinearOperatorException("hcat: inconsistent row sizes")
    nrow = size(A, 1)
    Ancol, Bncol = size(A, 2), size(B, 2)
    ncol = Ancol + Bncol
    prod_ = lambda res, v, a, b: _hcat_prod_(res, A, B, Ancol, ncol, v, a, b)
    tprod_ = lambda res, u, a, b: _hcat_ctprod_(res, transpose(A), transpose(B), Ancol, ncol, u, a, b)
    ctprod_ = lambda res, w, a, b: _hcat_ctprod_(res, adjoint(A), adjoint(B), Ancol, ncol, w, a, b)
    S = promote_storage(storage_type(A), storage_type(B))
    return LinearOperator(eltype(A), nrow, ncol, False, False, prod_, tprod_, ctprod_, S=S)


def hcat(*ops):
    """[A B ...]  (src/cat.jl:53-59); matrices are promoted with LinearOperator(M) (:3-5)"""
    op = _as_op(ops[0])
    for o in ops[1:]:
        op = _hcat2(op, _as_op(o))
    return op


def _vcat_prod_(res, A, B, Anrow, nV, u, alpha, beta):
    """src/cat.jl:65-77"""
    mul_(res[:Anrow], A, u, alpha, beta)
    mul_(res[Anrow:nV], B, u, alpha, beta)


def _vcat_ctprod_(res, A, B, Anrow, nV, v, alpha, beta):
    """src/cat.jl:79-91"""
    mul_(res, A, v[:Anrow], alpha, beta)
    mul_(res, B, v[Anrow:nV], alpha, 1.0)


def _vcat2(A, B):
    """src/cat.jl:93-109"""
    if size(A, 2) != size(B, 2):
        raise LinearOperatorException("vcat: inconsistent column sizes")
    Anrow, Bnrow = size(A, 1), size(B, 1)
    nrow = Anrow + Bnrow
    ncol = size(A, 2)
    prod_ = lambda res, v, a, b: _vcat_prod_(res, A, B, Anrow, nrow, v, a, b)
    tprod_ = lambda res, u, a, b: _vcat_ctprod_(res, transpose(A), transpose(B), Anrow, nrow, u, a, b)
    ctprod_ = lambda res, w, a, b: _vcat_ctprod_(res, adjoint(A), adjoint(B), Anrow, nrow, w, a, b)
    S = promote_storage(storage_type(A), storage_type(B))
    return LinearOperator(eltype(A), nrow, ncol, False, False, prod_, tprod_, ctprod_, S=S)


def vcat(*ops):
    """[A; B; ...]  (src/cat.jl:111-117); matrices are promoted with LinearOperator(M) (:61-63)"""
    op = _as_op(ops[0])
    for o in ops[1:]:
        op = _vcat2(op, _as_op(o))
    return op


def hvcat(rows, *ops):
    """hvcat((r1, r2, ...), ops...)  (src/cat.jl:120-129)"""
    rs, a = [], 0
    for r in rows:
        rs.append(hcat(*ops[a:a + r]))
        a += r
    return vcat(*rs)
