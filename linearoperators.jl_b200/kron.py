"""kron(A, B) -- mirror of src/kron.jl: `(A ⊗ B) * vec(X) = vec(B X Aᵀ)`.  The apply is a TMA-fed tcgen05 GEMM pair in one
clustered launch (csrc/b2o_kron.cu): bf16 operands, fp32 accumulation in TMEM, bf16 result."""
import ctypes

import numpy as np

from . import _lib
from ._lib import LinearOperatorException
from .abstract import LinearOperator, Storage
from .context import default_context


def _bf16_ptr(t, what):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.bfloat16 or not t.is_contiguous():
        raise _lib.B2OError("%s must be a contiguous torch CUDA bfloat16 tensor (no CPU fallback)" % what)
    return ctypes.c_void_p(t.data_ptr())


def _res_ptr(t):
    """result vector: bfloat16 (the reference's promoted eltype) or float32"""
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype not in (torch.bfloat16, torch.float32) or not t.is_contiguous():
        raise _lib.B2OError("res must be a contiguous torch CUDA bfloat16 or float32 tensor")
    return ctypes.c_void_p(t.data_ptr()), (_lib.B2O_BF16 if t.dtype == torch.bfloat16 else _lib.B2O_F32)


class KronOperator(LinearOperator):
    def apply_batch(self, X, alpha=1.0, beta=0.0, res=None, trans=False):
        """nb right-hand sides at once: X is (nb, ncol) row-major (each row one vec), result (nb, nrow)."""
        import torch
        nb = X.shape[0]
        ncol, nrow = (self.nrow, self.ncol) if trans else (self.ncol, self.nrow)
        if res is None:
            res = torch.empty((nb, nrow), dtype=torch.bfloat16, device=X.device)
        rp, rd = _res_ptr(res)
        _lib.check(self.ctx.lib.b2o_kron_apply(self._h, int(trans), rp, rd, nrow, _bf16_ptr(X, "x"), ncol, nb,
                                               float(alpha), float(beta)))
        return res

    def set_option(self, key, value):
        """tuning overrides of the clustered kernel: "cluster" (CTAs per 128-row unit), "tile_n" (columns per CTA tile); 0 = auto"""
        _lib.check(self.ctx.lib.b2o_kron_set_option(self._h, key.encode(), int(value)))
        return self

    def flops(self, nb=1):
        out = ctypes.c_double()
        _lib.check(self.ctx.lib.b2o_kron_flops(self._h, int(nb), ctypes.byref(out)))
        return out.value

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self.ctx.handle:
                self.ctx.lib.b2o_kron_destroy(self._h)
                self._h = None
        except Exception:
            pass


def _is_bf16_matrix(x):
    try:
        import torch
    except Exception:       # pragma: no cover
        return False
    return isinstance(x, torch.Tensor) and x.dim() == 2 and x.dtype == torch.bfloat16 and x.is_cuda


def _colmajor(x, r, c):
    """reshape(x, r, c) as Julia does it: an r x c VIEW of the vector x with unit-stride columns"""
    y = x.reshape(c, r)
    return y.T if isinstance(y, np.ndarray) else y.t()


def _rowmajor_copy(Mv):
    """contiguous row-major copy of a 2-D view (torch or numpy) -- i.e. the column-major storage of its transpose"""
    return Mv.contiguous() if hasattr(Mv, "contiguous") else np.ascontiguousarray(Mv)


def kron_operators(A, B):
    """kron(A::AbstractLinearOperator, B::AbstractLinearOperator) -- src/kron.jl:10-45, for ANY two operators (matrices are
    promoted with LinearOperator(M), :47-49): `(A ⊗ B) vec(X) = vec(B X Aᵀ)`.  B is applied to the columns of X and A to the rows
    of the result through the operators' own vector closures (matrix right-hand sides, src/operations.jl:34-36), so every
    multiply is one of the library's kernels; the two transposition copies and the final `α · + β res` are array plumbing.
    tprod! = vec(Bᵀ X A), ctprod! = vec(Bᴴ X conj(A)) are the same routine on the transposed / adjoint operators."""
    from .abstract import adjoint, ishermitian, issymmetric, mul_, promote_storage, size, storage_type, transpose
    from .abstract import _as_op
    A, B = _as_op(A), _as_op(B)
    m, n = size(A)
    p, q = size(B)
    S = promote_storage(storage_type(A), storage_type(B))

    def run(res, x, a, b, AA, BB):
        (mm, nn), (pp, qq) = size(AA), size(BB)
        X = _colmajor(x, qq, nn)                                   # reshape(x, q, n)                        kron.jl:16
        ybuf = S.alloc(pp * nn)
        Y = _colmajor(ybuf, pp, nn)
        mul_(Y, BB, X)                                             # B * X, column by column
        Yt = _rowmajor_copy(Y)                                     # (p x n) row-major: row l of Y is contiguous
        Yt = Yt.T if isinstance(Yt, np.ndarray) else Yt.t()       # n x p, unit-stride columns = rows of Y
        rbuf = S.alloc(mm * pp)
        Rt = _colmajor(rbuf, mm, pp)
        mul_(Rt, AA, Yt)                                           # column l: A * Y[l, :]  ==  row l of (B X) Aᵀ
        vec_r = _rowmajor_copy(Rt).reshape(-1)                     # vec(B X Aᵀ): entry l + i p = Rt[i, l]
        if b == 0:
            res[:] = a * vec_r                                     # res .= α .* Matrix(...)[:]                kron.jl:18
        else:
            res[:] = a * vec_r + b * res                           #                      ... .+ β .* res      :20

    def prod_(res, x, a, b):
        run(res, x, a, b, A, B)

    def tprod_(res, x, a, b):
        run(res, x, a, b, transpose(A), transpose(B))              # vec(Bᵀ X A)                                :22-30

    def ctprod_(res, x, a, b):
        run(res, x, a, b, adjoint(A), adjoint(B))                  # vec(Bᴴ X conj(A)) = vec(Bᴴ X (Aᴴ)ᵀ)        :31-39

    from .abstract import _promote_eltype, eltype
    T = _promote_eltype(eltype(A), eltype(B))                      # promote_type(eltype(A), eltype(B))         kron.jl:13
    return LinearOperator(T, m * p, n * q, issymmetric(A) and issymmetric(B), ishermitian(A) and ishermitian(B),
                          prod_, tprod_, ctprod_, S=S)


def kron(A, B, max_batch=1, ctx=None):
    """kron(A, B) -- src/kron.jl.  Two torch CUDA bfloat16 matrices A (m×n), B (p×q) take the tensor-core path (one TMA-fed
    tcgen05 GEMM pair, csrc/b2o_kron.cu); anything else -- operators, Float64 / Float32 / sparse matrices, or a mix (kron.jl:47-49)
    -- goes through `kron_operators`.  Result size (m*p) × (n*q).
    Vectors follow Julia's vec(): x[l + j*q] = X[l, j] (column-major reshape), exactly as src/kron.jl:16."""
    if not (_is_bf16_matrix(A) and _is_bf16_matrix(B)):
        return kron_operators(A, B)
    import torch
    ctx = ctx or default_context()
    if A.dim() != 2 or B.dim() != 2:
        raise LinearOperatorException("kron needs two matrices")
    m, n = A.shape
    p, q = B.shape
    A_cm, B_cm = A.t().contiguous(), B.t().contiguous()          # column-major images of A and B (Julia layout)
    h = ctypes.c_void_p()
    _lib.check(ctx.lib.b2o_kron_create(ctx.handle, _lib.B2O_BF16, _bf16_ptr(A_cm, "A"), m, n, _bf16_ptr(B_cm, "B"), p, q,
                                       int(max_batch), ctypes.byref(h)))
    lib = ctx.lib

    def prod_(res, x, a, b):
        rp, rd = _res_ptr(res)
        _lib.check(lib.b2o_kron_apply(h, 0, rp, rd, res.shape[0], _bf16_ptr(x, "x"), x.shape[0], 1, float(a), float(b)))

    def tprod_(res, x, a, b):
        rp, rd = _res_ptr(res)
        _lib.check(lib.b2o_kron_apply(h, 1, rp, rd, res.shape[0], _bf16_ptr(x, "x"), x.shape[0], 1, float(a), float(b)))

    op = KronOperator(torch.bfloat16, m * p, n * q, False, False, prod_, tprod_, tprod_,
                      S=Storage("cuda", ctx.device, dtype=torch.bfloat16))
    op.ctx, op._h, op._keep = ctx, h, (A_cm, B_cm)
    return op
