"""Leaf operators -- mirror of src/special-operators.jl and opHouseholder of src/linalg.jl.
Every closure body is one call across the C ABI (include/b2o.h) into a hand-written sm_100a kernel."""
import ctypes

import numpy as np

from . import _lib
from ._lib import LinearOperatorException
from .abstract import (AbstractLinearOperator, LinearOperator, Storage, adjoint, eltype, ishermitian, issymmetric,
                       mul_, op_times_op, size, storage_type, transpose)
from .context import default_context

F64 = _lib.B2O_F64


def _vp(t, what="vector"):
    """raw device pointer of a 1-D unit-stride float64 CUDA tensor (views with an offset are fine)."""
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.B2OError("%s must be a torch CUDA tensor (no CPU fallback)" % what)
    if t.dtype != torch.float64:
        raise _lib.B2OError("%s must be float64 (got %s)" % (what, t.dtype))
    if t.dim() != 1 or (t.numel() > 1 and t.stride(0) != 1):
        raise _lib.B2OError("%s must be a unit-stride 1-D tensor" % what)
    return ctypes.c_void_p(t.data_ptr())


def _cp(t, what="vector"):
    """raw device pointer of a 1-D unit-stride complex128 CUDA tensor (interleaved re, im)"""
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.B2OError("%s must be a torch CUDA tensor (no CPU fallback)" % what)
    if t.dtype != torch.complex128:
        raise _lib.B2OError("%s must be complex128 (got %s)" % (what, t.dtype))
    if t.dim() != 1 or (t.numel() > 1 and t.stride(0) != 1):
        raise _lib.B2OError("%s must be a unit-stride 1-D tensor" % what)
    return ctypes.c_void_p(t.resolve_conj().data_ptr() if t.is_conj() else t.data_ptr())


def _is_c128(t):
    import torch
    return isinstance(t, torch.Tensor) and t.dtype == torch.complex128


def _ri(z):
    z = complex(z)
    return float(z.real), float(z.imag)


def _ctx_of(like=None, ctx=None):
    if ctx is not None:
        return ctx
    if like is not None:
        c = getattr(like, "ctx", None)
        if c is not None:
            return c
        if isinstance(like, AbstractLinearOperator):
            p = like
            while hasattr(p, "parent"):
                p = p.parent
            c = getattr(p, "ctx", None)
            if c is not None:
                return c
    return default_context()


def _float_type():
    import torch
    return torch.float64


class _Leaf(LinearOperator):
    def __init__(self, ctx, T, nrow, ncol, symmetric, hermitian, prod_, tprod_, ctprod_):
        # storage_type: CuVector{T}; temporaries of composed operators (vtmp, Mv) must carry a complex element type
        cplx = getattr(T, "is_complex", False)
        super().__init__(T, nrow, ncol, symmetric, hermitian, prod_, tprod_, ctprod_,
                         S=Storage("cuda", ctx.device, T if cplx else None))
        self.ctx = ctx


class _OpEyeSingleton(AbstractLinearOperator):
    """opEye(): identity singleton; `opEye() * v is v` (src/special-operators.jl:14-34)."""
    is_identity_singleton = True
    nrow = ncol = None
    eltype = None
    symmetric = hermitian = True

    def __mul__(self, other):
        return other

    def __rmul__(self, other):
        return other

    def __repr__(self):
        return "Identity operator\n"


def opEye(*args, T=None, ctx=None, like=None):
    """opEye() | opEye(n) | opEye(nrow, ncol)  (src/special-operators.jl:14-77)"""
    if len(args) == 0:
        return _OpEyeSingleton()
    ctx = _ctx_of(like, ctx)
    T = T or _float_type()
    nrow = int(args[0])
    ncol = int(args[1]) if len(args) > 1 else nrow
    lib, h = ctx.lib, ctx.handle

    def prod_(res, v, a, b):
        # mulOpEye!(res, v, α, β, n_min): n_min = min(nrow, ncol) is the same for prod!/tprod!/ctprod!
        if _is_c128(res):
            _lib.check(lib.b2o_ceye_apply(h, res.shape[0], v.shape[0], _cp(res), res.shape[0], _cp(v), v.shape[0], *_ri(a), *_ri(b)))
            return
        _lib.check(lib.b2o_eye_apply(h, F64, res.shape[0], v.shape[0], _vp(res), res.shape[0], _vp(v), v.shape[0],
                                     float(a), float(b)))

    if nrow == ncol:
        op = _Leaf(ctx, T, nrow, ncol, True, True, prod_, prod_, prod_)
        op._expr = ("eye",)
        return op
    return _Leaf(ctx, T, nrow, ncol, False, False, prod_, prod_, prod_)


def opOnes(nrow, ncol, T=None, ctx=None, like=None):
    """opOnes(T, nrow, ncol) (src/special-operators.jl:79-100)"""
    ctx = _ctx_of(like, ctx)
    lib, h = ctx.lib, ctx.handle
    nrow, ncol = int(nrow), int(ncol)

    def prod_(res, v, a, b):
        _lib.check(lib.b2o_ones_apply(h, F64, res.shape[0], v.shape[0], _vp(res), res.shape[0], _vp(v), v.shape[0],
                                      float(a), float(b)))

    op = _Leaf(ctx, T or _float_type(), nrow, ncol, nrow == ncol, nrow == ncol, prod_, prod_, prod_)
    if nrow == ncol:
        op._expr = ("ones",)
    return op


def opZeros(nrow, ncol, T=None, ctx=None, like=None):
    """opZeros(T, nrow, ncol) (src/special-operators.jl:102-123)"""
    ctx = _ctx_of(like, ctx)
    lib, h = ctx.lib, ctx.handle
    nrow, ncol = int(nrow), int(ncol)

    def prod_(res, v, a, b):
        if _is_c128(res):
            _lib.check(lib.b2o_czeros_apply(h, res.shape[0], v.shape[0], _cp(res), res.shape[0], v.shape[0], *_ri(b)))
            return
        _lib.check(lib.b2o_zeros_apply(h, F64, res.shape[0], v.shape[0], _vp(res), res.shape[0], v.shape[0],
                                       float(a), float(b)))

    op = _Leaf(ctx, T or _float_type(), nrow, ncol, nrow == ncol, nrow == ncol, prod_, prod_, prod_)
    if nrow == ncol:
        op._expr = ("zeros",)
    return op


def opDiagonal(*args, ctx=None):
    """opDiagonal(d) | opDiagonal(nrow, ncol, d)  (src/special-operators.jl:125-165).
    `d` is aliased, not copied: mutating it later changes the operator, as in the reference."""
    if len(args) == 1:
        d = args[0]
        nrow = ncol = d.shape[0]
    else:
        nrow, ncol, d = int(args[0]), int(args[1]), args[2]
        if nrow == ncol <= d.shape[0]:
            return opDiagonal(d[:nrow].clone(), ctx=ctx)          # opDiagonal(d[1:nrow]) copies  :159
    ctx = _ctx_of(None, ctx)
    lib, h = ctx.lib, ctx.handle
    if _is_c128(d):
        # complex T (src/special-operators.jl:137-141): prod! = tprod! use d, ctprod! uses conj.(d); hermitian only if isreal(d)
        def cprod(conj_d):
            def f(res, v, a, b):
                _lib.check(lib.b2o_cdiag_apply(h, res.shape[0], v.shape[0], _cp(d, "diagonal"), d.shape[0], int(conj_d), _cp(res),
                                               res.shape[0], _cp(v), v.shape[0], *_ri(a), *_ri(b)))
            return f
        herm = bool((d.imag == 0).all().item())
        if nrow == ncol:
            return _Leaf(ctx, d.dtype, nrow, ncol, True, herm, cprod(False), cprod(False), cprod(True))
        return _Leaf(ctx, d.dtype, nrow, ncol, False, False, cprod(False), cprod(False), cprod(True))
    dp = _vp(d, "diagonal")

    def prod_(res, v, a, b):
        # real T: conj.(d) == d, so prod!/tprod!/ctprod! coincide
        _lib.check(lib.b2o_diag_apply(h, F64, res.shape[0], v.shape[0], _vp(d, "diagonal"), d.shape[0], _vp(res),
                                      res.shape[0], _vp(v), v.shape[0], float(a), float(b)))

    del dp
    if nrow == ncol:
        op = _Leaf(ctx, d.dtype, nrow, ncol, True, True, prod_, prod_, prod_)
        op._expr = ("diag", d)
        return op
    return _Leaf(ctx, d.dtype, nrow, ncol, False, False, prod_, prod_, prod_)


def opHouseholder(h, ctx=None):
    """opHouseholder(h): x -> (I - 2 h hᵀ) x  (src/linalg.jl:77-95); tprod! is inferred (None)."""
    ctx = _ctx_of(None, ctx)
    lib, hd = ctx.lib, ctx.handle
    n = h.shape[0]
    if _is_c128(h):
        # complex h: symmetric only if isreal(h), always hermitian; tprod! is inferred through the conj-sandwich (src/linalg.jl:91-95)
        def cprod_(res, v, a, b):
            _lib.check(lib.b2o_chouseholder_apply(hd, n, _cp(h, "h"), _cp(res), res.shape[0], _cp(v), v.shape[0], *_ri(a), *_ri(b)))
        return _Leaf(ctx, h.dtype, n, n, bool((h.imag == 0).all().item()), True, cprod_, None, cprod_)
    _vp(h, "h")

    def prod_(res, v, a, b):
        _lib.check(lib.b2o_householder_apply(hd, F64, n, _vp(h, "h"), _vp(res), res.shape[0], _vp(v), v.shape[0],
                                             float(a), float(b)))

    op = _Leaf(ctx, h.dtype, n, n, True, True, prod_, None, prod_)
    op._expr = ("house", h)
    return op


def _expand_index(idx, ncol):
    """LinearOperatorIndexType: vector of ints, range, or a single int; 1-based as in the reference."""
    if isinstance(idx, slice):
        start = 1 if idx.start is None else idx.start
        stop = ncol if idx.stop is None else idx.stop
        step = 1 if idx.step is None else idx.step
        return np.arange(start, stop + 1, step, dtype=np.int64)     # Julia ranges are inclusive
    if isinstance(idx, range):
        return np.asarray(list(idx), dtype=np.int64)
    if isinstance(idx, (int, np.integer)):
        return np.asarray([idx], dtype=np.int64)
    return np.ascontiguousarray(np.asarray(idx, dtype=np.int64))


class _IndexHandle:
    def __init__(self, ctx, idx1, ncol):
        self.ctx = ctx
        self.h = ctypes.c_void_p()
        idx1 = np.ascontiguousarray(idx1, dtype=np.int64)
        _lib.check(ctx.lib.b2o_index_create(ctx.handle, idx1.ctypes.data_as(ctypes.c_void_p), idx1.shape[0], int(ncol),
                                            ctypes.byref(self.h)))

    def __del__(self):
        try:
            if self.h and self.ctx.handle:
                self.ctx.lib.b2o_index_destroy(self.h)
        except Exception:
            pass


def opRestriction(Idx, ncol, ctx=None):
    """opRestriction(I, ncol): Z*v == v[I]  (src/special-operators.jl:167-201).  Indices are 1-based;
    `slice(None)` (Julia's `:`) gives the identity (:203)."""
    if isinstance(Idx, slice) and Idx == slice(None):
        return opEye(ncol, ctx=ctx)
    ctx = _ctx_of(None, ctx)
    idx1 = _expand_index(Idx, ncol)
    ix = _IndexHandle(ctx, idx1, ncol)      # raises "indices should be between 1 and ncol"
    lib = ctx.lib
    nrow = idx1.shape[0]

    def prod_(res, v, a, b):                 # mulRestrict!: α, β ignored (Q1)
        _lib.check(lib.b2o_restrict_apply(ix.h, F64, _vp(res), res.shape[0], _vp(v), v.shape[0]))

    def tprod_(res, u, a, b):                # multRestrict!
        _lib.check(lib.b2o_extend_apply(ix.h, F64, _vp(res), res.shape[0], _vp(u), u.shape[0]))

    op = _Leaf(ctx, "Int64", nrow, int(ncol), False, False, prod_, tprod_, tprod_)
    op._index = ix
    return op


def opExtension(Idx, ncol, ctx=None):
    """opExtension(I, ncol) = opRestriction(I, ncol)'  (src/special-operators.jl:216-221)"""
    if isinstance(Idx, slice) and Idx == slice(None):
        return opEye(ncol, ctx=ctx)
    return adjoint(opRestriction(Idx, ncol, ctx=ctx))


def getindex(op, rows, cols):
    """op[rows, cols] = R * op * E  (src/special-operators.jl:225-233)"""
    ctx = _ctx_of(op)
    R = opRestriction(rows, size(op, 1), ctx=ctx)
    E = opExtension(cols, size(op, 2), ctx=ctx)
    return op_times_op(op_times_op(R, op), E)


def BlockDiagonalOperator(*ops, S=None):
    """BlockDiagonalOperator(M1, ..., Mn) (src/special-operators.jl:249-294): block k maps the k-th slab of x
    to the k-th slab of y; slabs are views, no copies.  Blocks may be matrices (the reference calls `mul!` on them
    directly, :258-267; here they become LinearOperator(M) leaves) -- test/gpu/nvidia.jl:8-15."""
    from .abstract import _as_op
    ops = tuple(_as_op(op) for op in ops)
    nrow = ncol = 0
    for op in ops:
        m, n = size(op)
        nrow += m
        ncol += n
    if S is None:
        S = storage_type(ops[0])
        for op in ops[1:]:
            from .abstract import promote_storage
            S = promote_storage(S, storage_type(op))

    def prod_(y, x, a, b):
        k = j = 0
        for op in ops:
            m, n = size(op)
            mul_(y[k:k + m], op, x[j:j + n], a, b)
            k += m
            j += n

    def tprod_(y, x, a, b):
        k = j = 0
        for op in ops:
            m, n = size(op)
            mul_(y[k:k + n], transpose(op), x[j:j + m], a, b)
            k += n
            j += m

    def ctprod_(y, x, a, b):
        k = j = 0
        for op in ops:
            m, n = size(op)
            mul_(y[k:k + n], adjoint(op), x[j:j + m], a, b)
            k += n
            j += m

    symm = all(issymmetric(op) for op in ops)
    herm = all(ishermitian(op) for op in ops)
    out = LinearOperator(eltype(ops[0]), nrow, ncol, symm, herm, prod_, tprod_, ctprod_, S=S)
    out.ctx = _ctx_of(ops[0]) if S.kind == "cuda" else None
    return out


def LocalBlockOfDiagonal(op, rank, world):
    """BlockDiagonalOperator with ONE BLOCK PER GPU (one process per GPU): rank r owns block r and the r-th slab of x and y,
    so `mul!` of the global block-diagonal operator is just the local block applied to the local slab -- no collective
    (src/special-operators.jl:258-267 evaluates the blocks one after the other on one device).  The block must live on a
    context WITHOUT a communicator (its inner products are local to the block); this helper checks that and returns it."""
    ctx = _ctx_of(op)
    if ctx is not None and getattr(ctx, "nranks", 1) != 1:
        raise LinearOperatorException("a block of a block-diagonal operator must use a local (non row-partitioned) context")
    op.block_rank, op.block_world = int(rank), int(world)
    return op
