"""LBFGSOperator / InverseLBFGSOperator / LSR1Operator -- mirror of src/lbfgs.jl and src/lsr1.jl.

State (the {s,y,a,b} columns, ring index, scaling factor) lives in HBM inside a libb2o handle; the apply is
ONE persistent sm_100a kernel (TMA-staged columns, all dots and axpys in a single launch)."""
import ctypes

import numpy as np

from . import _lib
from ._lib import ErrorException, LinearOperatorException
from .abstract import AbstractLinearOperator, Storage
from .context import default_context
from .special_operators import _vp

F64 = _lib.B2O_F64
F32 = _lib.B2O_F32


def _qvp(op, t, what="vector"):
    """raw device pointer of a unit-stride 1-D CUDA tensor whose dtype is the operator's element type"""
    import torch
    if op.eltype == torch.float64:
        return _vp(t, what)
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise _lib.B2OError("%s must be a torch CUDA tensor (no CPU fallback)" % what)
    if t.dtype != op.eltype:
        raise _lib.B2OError("%s must be %s (got %s)" % (what, op.eltype, t.dtype))
    if t.dim() != 1 or (t.numel() > 1 and t.stride(0) != 1):
        raise _lib.B2OError("%s must be a unit-stride 1-D tensor" % what)
    return ctypes.c_void_p(t.data_ptr())


def _split_T(first, rest, T):
    """the reference's positional form Op(T, n) beside Op(n, T=...)"""
    if len(rest) > 1:
        raise TypeError("expected (n) or (T, n)")
    if len(rest) == 1:
        if isinstance(first, (int, np.integer)):
            raise TypeError("mem, scaling, ... are keyword arguments (src/lbfgs.jl:168); positional forms are Op(n) and Op(T, n)")
        if T is not None:
            raise TypeError("element type given twice")
        return rest[0], first
    return first, T


def _eltype_code(T):
    """LBFGSOperator(T, n; ...) (src/lbfgs.jl:168, src/lsr1.jl:86): Float64 (default) and Float32 are built"""
    import torch
    if T is None or T == torch.float64 or T is float or T == np.float64:
        return torch.float64, F64
    if T == torch.float32 or T == np.float32:
        return torch.float32, F32
    raise _lib.B2OError("quasi-Newton operators are built for Float64 and Float32 (got %r)" % (T,))


class _QNData:
    """read-only view of op.data (LBFGSData src/lbfgs.jl:4-24 / LSR1Data src/lsr1.jl:4-17)."""

    def __init__(self, op):
        self._op = op

    def _scalars(self):
        op = self._op
        ins, g, ub = ctypes.c_int(), ctypes.c_double(), ctypes.c_double()
        ys = (ctypes.c_double * op.mem)()
        aux = (ctypes.c_double * op.mem)()
        _lib.check(op.ctx.lib.b2o_qn_get_scalars(op.handle, ctypes.byref(ins), ctypes.byref(g), ctypes.byref(ub), ys, aux))
        return ins.value, g.value, ub.value, np.array(ys[:]), np.array(aux[:])

    mem = property(lambda self: self._op.mem)
    scaling = property(lambda self: self._op.scaling)
    damped = property(lambda self: self._op.damped)
    insert = property(lambda self: self._scalars()[0])
    scaling_factor = property(lambda self: self._scalars()[1])
    opnorm_upper_bound = property(lambda self: self._scalars()[2])
    ys = property(lambda self: self._scalars()[3])
    aux = property(lambda self: self._scalars()[4])

    def col(self, which, k0):
        """device copy of column `which` ('s','y','a','b') in 0-based ring slot k0"""
        op = self._op
        out = op.ctx.empty(op.nrow, dtype=op.eltype)
        _lib.check(op.ctx.lib.b2o_qn_get_col(op.handle, "syab".index(which), int(k0), _qvp(op, out)))
        return out

    def set_col(self, which, k0, src):
        op = self._op
        _lib.check(op.ctx.lib.b2o_qn_set_col(op.handle, "syab".index(which), int(k0), _qvp(op, src)))

    def set_scalars(self, insert, scaling_factor, opnorm_upper_bound, ys, aux):
        op = self._op
        ysb = (ctypes.c_double * op.mem)(*[float(v) for v in ys])
        axb = (ctypes.c_double * op.mem)(*[float(v) for v in aux])
        _lib.check(op.ctx.lib.b2o_qn_set_scalars(op.handle, int(insert), float(scaling_factor),
                                                 float(opnorm_upper_bound), ysb, axb))


class AbstractQuasiNewtonOperator(AbstractLinearOperator):
    """AbstractQuasiNewtonOperator{T} (src/qn.jl)."""
    always_allocated5 = True          # has_args5 / isallocated5 are hard-wired true (src/lbfgs.jl:101-102)

    def _common(self, ctx, n, mem, T=None):
        self.ctx = ctx
        self.eltype, self._dt = _eltype_code(T)
        self.nrow = self.ncol = int(n)
        self.symmetric = self.hermitian = True
        self.nprod = self.ntprod = self.nctprod = 0
        self.mem = max(int(mem), 1)
        self.S = Storage("cuda", ctx.device, dtype=None if self._dt == F64 else self.eltype)
        self.Mv = self.Mtu = None
        self.data = _QNData(self)
        lib, op = ctx.lib, self

        def prod_(res, x, a, b):
            _lib.check(lib.b2o_qn_apply(op.handle, _qvp(op, res), res.shape[0], _qvp(op, x), x.shape[0], float(a), float(b)))

        self.prod_ = prod_

    def __del__(self):
        try:
            if getattr(self, "handle", None) and self.ctx.handle:
                self.ctx.lib.b2o_qn_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def apply_host(self, res_host, x_host, alpha=1.0, beta=0.0):
        """end-to-end apply with HOST buffers (numpy / pinned torch CPU tensors): H2D, apply, D2H inside."""
        def hp(t):
            return ctypes.c_void_p(t.data_ptr() if hasattr(t, "data_ptr") else t.ctypes.data)
        _lib.check(self.ctx.lib.b2o_qn_apply_host(self.handle, hp(res_host), hp(x_host), int(x_host.shape[0]),
                                                  float(alpha), float(beta)))
        self.nprod += 1
        return res_host

    def _mul_block(self, res, X, alpha, beta):
        """mul!(Res::Matrix, op, X::Matrix, α, β) (src/operations.jl:34-36) in one launch per 8 right-hand sides; returns False
        (caller falls back to the column loop) unless both matrices are column-major float64 CUDA tensors."""
        import torch
        for t in (res, X):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == self.eltype and t.dim() == 2):
                return False
            if t.shape[0] > 1 and t.stride(0) != 1:
                return False
            if t.shape[1] > 1 and t.stride(1) < t.shape[0]:
                return False
        k = X.shape[1]
        _lib.check(self.ctx.lib.b2o_qn_apply_multi(self.handle, ctypes.c_void_p(res.data_ptr()), int(res.stride(1)) if k > 1 else self.nrow,
                                                   ctypes.c_void_p(X.data_ptr()), int(X.stride(1)) if k > 1 else self.nrow,
                                                   int(X.shape[0]), int(k), float(alpha), float(beta)))
        return True

    def apply_bytes(self, beta=0.0):
        out = ctypes.c_double()
        _lib.check(self.ctx.lib.b2o_qn_apply_bytes(self.handle, float(beta), ctypes.byref(out)))
        return out.value

    def set_option(self, key, value):
        """e.g. set_option("inverse_mode", 1): compact-representation apply for InverseLBFGSOperator"""
        _lib.check(self.ctx.lib.b2o_qn_set_option(self.handle, key.encode(), int(value)))
        return self

    def _reset_state(self):
        _lib.check(self.ctx.lib.b2o_qn_reset(self.handle))


class LBFGSOperator(AbstractQuasiNewtonOperator):
    """LBFGSOperator(T, n; mem=5, scaling=true, damped=false, σ₂=0.99, σ₃=10.0) -- forward form, src/lbfgs.jl:168-208.
    InverseLBFGSOperator builds the same type with inverse=True (:112-160).  T (keyword here): torch.float64 (default) or
    torch.float32 -- the Float32 operator keeps its state, x and res in Float32 (test/test_lbfgs.jl:162-178)."""

    def __init__(self, n, *args, mem=5, scaling=True, damped=False, sigma2=0.99, sigma3=10.0, inverse=False, compact=False, ctx=None, T=None):
        n, T = _split_T(n, args, T)                          # LBFGSOperator(T, n; ...) or LBFGSOperator(n; ..., T=...)
        ctx = ctx or default_context()
        self._common(ctx, n, mem, T)
        self.scaling, self.damped, self.inverse = bool(scaling), bool(damped), bool(inverse)
        self.handle = ctypes.c_void_p()
        _lib.check(ctx.lib.b2o_lbfgs_create(ctx.handle, self._dt, int(n), int(mem), int(scaling), int(damped), float(sigma2),
                                            float(sigma3), int(inverse), ctypes.byref(self.handle)))
        self.tprod_ = self.prod_
        self.ctprod_ = self.prod_
        if compact:
            # extension (not the reference algorithm): compact representation -- same operator; for the forward form push! then
            # costs O(m) dots instead of O(m²) vector passes, for the inverse form the apply moves half the bytes
            self.set_option("inverse_mode" if inverse else "forward_mode", 1)


def InverseLBFGSOperator(n, *args, compact=False, **kw):
    """InverseLBFGSOperator(n; mem, scaling, damped, σ₂, σ₃) (src/lbfgs.jl:112-160).  compact=True switches the apply from the
    reference's two-loop recursion to the mathematically identical compact representation (half the DRAM traffic, one
    all-reduce instead of 2m): an extension, not the reference algorithm -- rounding differs."""
    kw.pop("inverse", None)
    return LBFGSOperator(n, *args, inverse=True, compact=compact, **kw)


class LSR1Operator(AbstractQuasiNewtonOperator):
    """LSR1Operator(T, n; mem=5, scaling=true) -- src/lsr1.jl:86-113 (tprod!/ctprod! are `nothing`: inferred).
    T (keyword): torch.float64 (default) or torch.float32 (test/test_lsr1.jl:74-86)."""

    def __init__(self, n, *args, mem=5, scaling=True, ctx=None, T=None):
        n, T = _split_T(n, args, T)
        ctx = ctx or default_context()
        self._common(ctx, n, mem, T)
        self.scaling, self.damped, self.inverse = bool(scaling), False, False
        self.handle = ctypes.c_void_p()
        _lib.check(ctx.lib.b2o_lsr1_create(ctx.handle, self._dt, int(n), int(mem), int(scaling), ctypes.byref(self.handle)))
        self.tprod_ = None
        self.ctprod_ = None


def push_(op, s, y, *rest):
    """push!(op, s, y) | push!(op, s, y, Bs) | push!(op, s, y, α, g) | push!(op, s, y, α, g, Bs)
    (src/lbfgs.jl:269-367, src/lsr1.jl:119-184).  Returns op; a rejected pair leaves the state unchanged."""
    from .diagqn import AbstractDiagonalQuasiNewtonOperator
    if isinstance(op, AbstractDiagonalQuasiNewtonOperator):
        if rest:
            raise TypeError("no such push! method for diagonal quasi-Newton operators")
        return op._push(s, y)
    lib = op.ctx.lib
    acc = ctypes.c_int(0)
    n = s.shape[0]
    if isinstance(op, LSR1Operator) or len(rest) == 0:
        if len(rest) != 0:
            raise TypeError("no such push! method for LSR1Operator")
        _lib.check(lib.b2o_qn_push(op.handle, _qvp(op, s), _qvp(op, y), n, ctypes.byref(acc)))
        if isinstance(op, LSR1Operator) or (op.damped and not op.inverse):
            op.nprod += 1          # push! runs mul!(Bs, op, s) first (src/lsr1.jl:125, src/lbfgs.jl:305 through :273-277)
    elif len(rest) == 1:
        (Bs,) = rest
        _lib.check(lib.b2o_lbfgs_push_damped_fwd(op.handle, _qvp(op, s), _qvp(op, y), _qvp(op, Bs), n, ctypes.byref(acc)))
        op.nprod += 1              # mul!(Bs, op, s)  src/lbfgs.jl:305
    elif len(rest) in (2, 3):
        alpha, g = rest[0], rest[1]
        Bs = rest[2] if len(rest) == 3 else op.ctx.empty(n, dtype=op.eltype)      # similar(g)  src/lbfgs.jl:366
        _lib.check(lib.b2o_lbfgs_push_damped_inv(op.handle, _qvp(op, s), _qvp(op, y), float(alpha), _qvp(op, g), _qvp(op, Bs), n,
                                                 ctypes.byref(acc)))
    else:
        raise TypeError("no such push! method")
    op.last_push_accepted = bool(acc.value)
    return op


def diag_(op, d):
    """diag!(op, d) (src/lbfgs.jl:379-395, src/lsr1.jl:196-211)"""
    _lib.check(op.ctx.lib.b2o_qn_diag(op.handle, _qvp(op, d), d.shape[0]))
    return d


def diag(op):
    return diag_(op, op.ctx.empty(op.nrow, dtype=op.eltype))


def solve_shifted_system_(x, B, b, sigma):
    """solve_shifted_system!(x, B, b, σ): solve (B + σI) x = b for a forward LBFGSOperator (src/utilities.jl:207-248)."""
    if sigma < 0:
        raise ValueError("σ must be nonnegative")                      # ArgumentError
    _lib.check(B.ctx.lib.b2o_lbfgs_solve_shifted(B.handle, _vp(x), x.shape[0], _vp(b), b.shape[0], float(sigma)))
    return x


def ldiv_(x, B, b):
    """ldiv!(x, B, b): solve B x = b (src/utilities.jl:281-289)"""
    return solve_shifted_system_(x, B, b, 0.0)
