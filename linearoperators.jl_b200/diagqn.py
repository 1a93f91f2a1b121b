"""Diagonal quasi-Newton operators and ShiftedOperator -- mirror of src/DiagonalHessianApproximation.jl and
src/shifted_operators.jl.  Their apply is the diagonal kernel; push! is one fused reduction pass + one
elementwise update (csrc/b2o_leaf.cu: b2o_diagqn_push)."""
import ctypes

from . import _lib
from ._lib import ErrorException, LinearOperatorException
from .abstract import (AbstractLinearOperator, Storage, adjoint, eltype, ishermitian, issymmetric, mul_, size, storage_type,
                       transpose)
from .context import default_context
from .special_operators import _ctx_of, _vp

F64 = _lib.B2O_F64


class AbstractDiagonalQuasiNewtonOperator(AbstractLinearOperator):
    """AbstractDiagonalQuasiNewtonOperator{T}: d (device vector, aliased), prod!/tprod!/ctprod! = mulSquareOpDiagonal!"""
    always_allocated5 = True
    _kind = None

    def __init__(self, d, ctx=None):
        self.ctx = ctx or default_context()
        self.d = d
        n = d.shape[0]
        self._init_common(n)
        lib, h = self.ctx.lib, self.ctx.handle

        def prod_(res, v, a, b):
            _lib.check(lib.b2o_diag_apply(h, F64, res.shape[0], v.shape[0], _vp(self.d, "d"), self.d.shape[0], _vp(res),
                                          res.shape[0], _vp(v), v.shape[0], float(a), float(b)))

        self.prod_ = self.tprod_ = self.ctprod_ = prod_
        self._expr = ("diag", d)

    def _init_common(self, n):
        import torch
        self.eltype = torch.float64
        self.nrow = self.ncol = int(n)
        self.symmetric = self.hermitian = True
        self.nprod = self.ntprod = self.nctprod = 0
        self.S = Storage("cuda", self.ctx.device)
        self.Mv = self.Mtu = None

    def _push(self, s, y):
        st = self.ctx.lib.b2o_diagqn_push(self.ctx.handle, self._kind, _vp(self.d, "d"), self.d.shape[0], _vp(s), _vp(y), s.shape[0])
        _lib.check(st)
        return self

    def _reset_state(self):
        self.d.fill_(1.0)            # op.d .= one(T)   (src/DiagonalHessianApproximation.jl:71-77)


class DiagonalPSB(AbstractDiagonalQuasiNewtonOperator):
    """DiagonalPSB(d) (src/DiagonalHessianApproximation.jl:1-64)"""
    _kind = 0


class DiagonalAndrei(AbstractDiagonalQuasiNewtonOperator):
    """DiagonalAndrei(d) (:79-141)"""
    _kind = 1


class DiagonalBFGS(AbstractDiagonalQuasiNewtonOperator):
    """DiagonalBFGS(d) (:198-248)"""
    _kind = 2


class SpectralGradient(AbstractDiagonalQuasiNewtonOperator):
    """SpectralGradient(σ, n): σI with σ > 0 (:143-196); `d` is the 1-element vector [σ] as in the reference"""
    _kind = 3

    def __init__(self, sigma, n, ctx=None):
        import torch
        assert sigma > 0
        self.ctx = ctx or default_context()
        self.d = torch.full((1,), float(sigma), dtype=torch.float64, device="cuda:%d" % self.ctx.device)
        self._init_common(n)
        lib, h = self.ctx.lib, self.ctx.handle

        def prod_(res, v, a, b):
            # mulSquareOpDiagonal!(res, [σ], v, α, β): `α .* d .* v` broadcasts the 1-element d -> (α*σ) .* v
            sig = float(self.d.item())
            _lib.check(lib.b2o_eye_apply(h, F64, res.shape[0], v.shape[0], _vp(res), res.shape[0], _vp(v), v.shape[0],
                                         float(a) * sig, float(b)))

        self.prod_ = self.tprod_ = self.ctprod_ = prod_


class ShiftedOperator(AbstractLinearOperator):
    """ShiftedOperator(H, σ=0): op = H + σI with a mutable σ (src/shifted_operators.jl:1-103):
    y = α·H·x + β·y, then axpy!(α·σ, x, y)."""
    always_allocated5 = True

    def __init__(self, H, sigma=0.0):
        if size(H, 1) != size(H, 2):
            raise ValueError("Operator H must be square.")            # DimensionMismatch
        self.base, self.sigma = H, float(sigma)      # data.H, data.σ (`.H` is the adjoint shortcut of the mirror)
        self.ctx = _ctx_of(H)
        self.eltype = eltype(H)
        self.nrow = self.ncol = size(H, 1)
        self.symmetric, self._herm = issymmetric(H), ishermitian(H)
        self.nprod = self.ntprod = self.nctprod = 0
        self.S = storage_type(H)
        self.Mv = self.Mtu = None
        lib, h = self.ctx.lib, self.ctx.handle

        def axpy(y, x, a):
            if self.sigma != 0 and a != 0:                                                      # :19-21
                _lib.check(lib.b2o_eye_apply(h, F64, y.shape[0], x.shape[0], _vp(y), y.shape[0], _vp(x), x.shape[0],
                                             float(a) * self.sigma, 1.0))

        def prod_(y, x, a, b):
            mul_(y, self.base, x, a, b)
            axpy(y, x, a)

        def tprod_(y, x, a, b):
            mul_(y, transpose(self.base), x, a, b)
            axpy(y, x, a)

        def ctprod_(y, x, a, b):
            mul_(y, adjoint(self.base), x, a, b)
            axpy(y, x, a)

        self.prod_, self.tprod_, self.ctprod_ = prod_, tprod_, ctprod_

    @property
    def hermitian(self):
        return self._herm

    @property
    def data(self):
        return self


def push_diag(op, s, y):
    return op._push(s, y)
