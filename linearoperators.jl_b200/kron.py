"""kron(A, B) -- mirror of src/kron.jl: `(A ⊗ B) * vec(X) = vec(B X Aᵀ)`.  The apply is a TMA-fed tcgen05 GEMM pair in one
cooperative launch (csrc/b2o_kron.cu): bf16 operands, fp32 accumulation in TMEM, bf16 result."""
import ctypes

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


def kron(A, B, max_batch=1, ctx=None):
    """kron(A, B) with A (m×n), B (p×q) torch CUDA bfloat16 matrices -> operator of size (m*p) × (n*q).
    Vectors follow Julia's vec(): x[l + j*q] = X[l, j] (column-major reshape), exactly as src/kron.jl:16."""
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
