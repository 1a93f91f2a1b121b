"""fuse(op): collapse a STATIC composed operator (sums, products, scalar multiples, unary minus, adjoint/transpose of
square elementwise leaves: opDiagonal, opEye, opZeros, opOnes, opHouseholder) into ONE cooperative kernel launch
(csrc/b2o_graph.cu).  The lowering repeats mul!'s own recursion (src/operations.jl:117-128,163-177,187-197), so the
result equals the closure tree's: same α/β threading, same statement-level rounding, every dot/sum taken once, no
temporaries in HBM.  Trees containing anything else (quasi-Newton operators, index operators, rectangular leaves,
user closures) are not static-elementwise: fuse() raises and the closure tree (still all-CUDA) remains the path."""
import ctypes

from . import _lib
from ._lib import LinearOperatorException
from .abstract import (AdjointLinearOperator, ConjugateLinearOperator, LinearOperator, Storage, TransposeLinearOperator, size)
from .special_operators import _ctx_of, _vp

_LEAF = {"diag": 0, "eye": 1, "zeros": 2, "ones": 3, "house": 4}
_SUM, _PROD, _SCALE, _NEG, _TRANS = 10, 11, 12, 13, 14


class _Graph:
    def __init__(self, ctx, n):
        self.ctx = ctx
        self.h = ctypes.c_void_p()
        _lib.check(ctx.lib.b2o_graph_create(ctx.handle, int(n), ctypes.byref(self.h)))
        self.keep = []          # leaf tensors stay alive as long as the fused operator does (they are aliased)

    def __del__(self):
        try:
            if self.h and self.ctx.handle:
                self.ctx.lib.b2o_graph_destroy(self.h)
        except Exception:
            pass


def _build(g, op, n):
    lib = g.ctx.lib
    node = ctypes.c_int()
    if isinstance(op, (TransposeLinearOperator, AdjointLinearOperator)):     # real element type: adjoint == transpose
        c = _build(g, op.parent, n)
        _lib.check(lib.b2o_graph_unary(g.h, _TRANS, c, 0.0, ctypes.byref(node)))
        return node.value
    if isinstance(op, ConjugateLinearOperator):
        return _build(g, op.parent, n)
    e = getattr(op, "_expr", None)
    if e is None or size(op) != (n, n):
        raise LinearOperatorException("operator tree is not a static chain of square elementwise leaves; cannot fuse")
    kind = e[0]
    if kind in _LEAF:
        vec = None
        if kind in ("diag", "house"):
            vec = _vp(e[1], kind)
            g.keep.append(e[1])
        _lib.check(lib.b2o_graph_leaf(g.h, _LEAF[kind], vec, ctypes.byref(node)))
    elif kind in ("sum", "prod"):
        a, b = _build(g, e[1], n), _build(g, e[2], n)
        _lib.check(lib.b2o_graph_binary(g.h, _SUM if kind == "sum" else _PROD, a, b, ctypes.byref(node)))
    elif kind == "scale":
        if isinstance(e[2], complex):
            raise LinearOperatorException("complex scalars cannot be fused (Float64 kernels)")
        c = _build(g, e[1], n)
        _lib.check(lib.b2o_graph_unary(g.h, _SCALE, c, float(e[2]), ctypes.byref(node)))
    elif kind == "neg":
        c = _build(g, e[1], n)
        _lib.check(lib.b2o_graph_unary(g.h, _NEG, c, 0.0, ctypes.byref(node)))
    else:
        raise LinearOperatorException("cannot fuse node %r" % (kind,))
    return node.value


class FusedOperator(LinearOperator):
    """a composed operator evaluated by one launch; prod!/tprod!/ctprod! all go to b2o_graph_apply"""

    def info(self, transposed=False, beta=0.0):
        np_, nr, by = ctypes.c_int(), ctypes.c_int(), ctypes.c_double()
        _lib.check(self.ctx.lib.b2o_graph_info(self._graph.h, int(transposed), float(beta), ctypes.byref(np_), ctypes.byref(nr),
                                               ctypes.byref(by)))
        j = ctypes.c_int()
        _lib.check(self.ctx.lib.b2o_graph_uses_jit(self._graph.h, int(transposed), float(beta), ctypes.byref(j)))
        ex, hs = ctypes.c_int(), ctypes.c_uint64()
        _lib.check(self.ctx.lib.b2o_graph_variant(self._graph.h, int(transposed), float(beta), ctypes.byref(ex), ctypes.byref(hs)))
        return {"passes": np_.value, "reductions": nr.value, "alg_bytes": by.value, "jit": bool(j.value),
                "executor": ("interpreter", "nvrtc", "aot")[ex.value], "source_hash": "%016x" % hs.value}

    def source(self, transposed=False, beta=0.0):
        """CUDA source of the NVRTC-specialised kernel for this variant"""
        n = ctypes.c_int64()
        buf = ctypes.create_string_buffer(1 << 18)
        _lib.check(self.ctx.lib.b2o_graph_jit_source(self._graph.h, int(transposed), float(beta), buf, 1 << 18, ctypes.byref(n)))
        return buf.value.decode()


def fuse(op, ctx=None):
    ctx = _ctx_of(op, ctx)
    n, m = size(op)
    if n != m:
        raise LinearOperatorException("only square operator trees can be fused")
    g = _Graph(ctx, n)
    root = _build(g, op, n)
    _lib.check(ctx.lib.b2o_graph_compile(g.h, root))
    lib = ctx.lib

    def prod_(res, v, a, b):
        _lib.check(lib.b2o_graph_apply(g.h, 0, _vp(res), res.shape[0], _vp(v), v.shape[0], float(a), float(b)))

    def tprod_(res, v, a, b):
        _lib.check(lib.b2o_graph_apply(g.h, 1, _vp(res), res.shape[0], _vp(v), v.shape[0], float(a), float(b)))

    from .abstract import eltype, ishermitian, issymmetric
    out = FusedOperator(eltype(op), n, n, issymmetric(op), ishermitian(op), prod_, tprod_, tprod_, S=Storage("cuda", ctx.device))
    out.ctx = ctx
    out._graph = g
    return out
