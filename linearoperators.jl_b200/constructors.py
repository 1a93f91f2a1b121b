"""LinearOperator(M) for a dense matrix -- mirror of src/constructors.jl:15-29.

The three closures are `mul!(res, M, v, α, β)`, `mul!(res, transpose(M), u, α, β)`, `mul!(res, adjoint(M), w, α, β)`;
here each is one call of b2o_dense_apply (hand-written HBM-bound matrix-vector kernels, csrc/b2o_dense.cu).  This is the
leaf that `BlockDiagonalOperator(A, B, C)` of CUDA matrices runs (test/gpu/nvidia.jl:8-15), that `op * M`, `M + op`
(src/operations.jl:159-160,218-219,229-230) and `hcat(op, M)` / `vcat(M, op)` (src/cat.jl:3-5,61-63) promote to.

The matrix is ALIASED (the reference's closures capture M): later in-place changes of M change the operator.
Element types: float64 and float32 (real; adjoint ≡ transpose).  Layout: any 2-D torch CUDA tensor with one unit
stride -- a column-major (Julia-like) matrix is used as is, a row-major one (torch's default) is read as the
column-major storage of its transpose and the roles of the two kernels swap; nothing is copied."""
import ctypes

from . import _lib
from ._lib import LinearOperatorException
from .abstract import LinearOperator, Storage
from .context import default_context


def is_matrix(x):
    """AbstractMatrix in the reference's signatures: a 2-D torch tensor here"""
    return type(x).__module__.startswith("torch") and hasattr(x, "dim") and x.dim() == 2


def _dtype_code(t):
    import torch
    if t.dtype == torch.float64:
        return _lib.B2O_F64
    if t.dtype == torch.float32:
        return _lib.B2O_F32
    raise _lib.B2OError("LinearOperator(M): element type %s not supported (float64, float32)" % t.dtype)


def colmajor_view(nrow, ncol, s0, s1):
    """((rows, cols, lda), swap): the column-major matrix the C ABI sees for an nrow x ncol tensor with element strides
    (s0, s1).  swap=False: it is M itself; swap=True: it is the column-major storage of Mᵀ (a row-major M)."""
    if nrow == 0 or ncol == 0:
        return (nrow, ncol, max(1, nrow)), False
    if (s0 == 1 or nrow == 1) and (ncol == 1 or s1 >= nrow):
        return (nrow, ncol, max(1, nrow) if ncol == 1 else s1), False
    if (s1 == 1 or ncol == 1) and (nrow == 1 or s0 >= ncol):
        return (ncol, nrow, max(1, ncol) if nrow == 1 else s0), True
    raise _lib.B2OError("LinearOperator(M): M needs one unit stride (got strides %r); use M.contiguous()" % ((s0, s1),))


class DenseMatrixOperator(LinearOperator):
    """LinearOperator{T,S}(M; symmetric, hermitian) -- src/constructors.jl:19-29"""

    def __init__(self, M, symmetric=False, hermitian=False, ctx=None):
        import torch
        if not isinstance(M, torch.Tensor) or not M.is_cuda:
            raise _lib.B2OError("LinearOperator(M): M must be a torch CUDA tensor (no CPU fallback)")
        if M.dim() != 2:
            raise LinearOperatorException("LinearOperator(M) needs a matrix")
        self.ctx = ctx or default_context(M.device.index)
        if self.ctx.device != M.device.index:
            raise _lib.B2OError("LinearOperator(M): matrix lives on cuda:%d, context on cuda:%d" % (M.device.index, self.ctx.device))
        nrow, ncol = int(M.shape[0]), int(M.shape[1])
        s0, s1 = (int(x) for x in M.stride())
        cm, swap = colmajor_view(nrow, ncol, s0, s1)
        self.M = M                       # aliased, and kept alive for the handle
        self._swap = swap
        self._code = _dtype_code(M)
        self._h = ctypes.c_void_p()
        _lib.check(self.ctx.lib.b2o_dense_create(self.ctx.handle, self._code, ctypes.c_void_p(M.data_ptr()), cm[0], cm[1], cm[2],
                                                 ctypes.byref(self._h)))

        def prod_(res, v, a, b):         # mul!(res, M, v, α, β)                              constructors.jl:25
            self._run(1 if self._swap else 0, res, v, a, b)

        def tprod_(res, u, a, b):        # mul!(res, transpose(M), u, α, β) (≡ adjoint, real T) constructors.jl:26-27
            self._run(0 if self._swap else 1, res, u, a, b)

        super().__init__(M.dtype, nrow, ncol, symmetric, hermitian, prod_, tprod_, tprod_,
                         S=Storage("cuda", self.ctx.device, dtype=M.dtype))

    def _vec(self, t, what):
        import torch
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise _lib.B2OError("%s must be a torch CUDA tensor (no CPU fallback)" % what)
        if t.dtype != self.M.dtype:
            raise _lib.B2OError("%s must be %s like the matrix (got %s)" % (what, self.M.dtype, t.dtype))
        if t.dim() != 1 or (t.numel() > 1 and t.stride(0) != 1):
            raise _lib.B2OError("%s must be a unit-stride 1-D tensor" % what)
        return ctypes.c_void_p(t.data_ptr())

    def _run(self, trans, res, v, alpha, beta):
        _lib.check(self.ctx.lib.b2o_dense_apply(self._h, int(trans), self._vec(res, "res"), res.shape[0], self._vec(v, "v"),
                                                v.shape[0], float(alpha), float(beta)))

    def apply_bytes(self, trans=False, beta=0.0):
        """algorithmic DRAM bytes of one product (matrix once + vectors), for the roofline"""
        out = ctypes.c_double()
        t = (0 if trans else 1) if self._swap else (1 if trans else 0)
        _lib.check(self.ctx.lib.b2o_dense_apply_bytes(self._h, t, float(beta), ctypes.byref(out)))
        return out.value

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self.ctx.handle:
                self.ctx.lib.b2o_dense_destroy(self._h)
                self._h = None
        except Exception:
            pass


def _is_sparse(M):
    import torch
    return isinstance(M, torch.Tensor) and M.layout in (torch.sparse_csc, torch.sparse_csr)


class SparseMatrixOperator(LinearOperator):
    """LinearOperator(M; symmetric, hermitian) for a sparse matrix -- src/constructors.jl:15-29 with M::SparseMatrixCSC (the
    docstring :3-5 says "dense or sparse"; test/test_linop.jl:41-75 runs every predicate on both).  M: a torch CUDA tensor
    with layout sparse_csc (Julia's SparseMatrixCSC: colptr, rowval, nzval) or sparse_csr, float64 or float32.  The index
    arrays are copied to the host once (the C ABI takes the reference's 1-based arrays and transposes the structure there);
    the values are ALIASED -- after changing them in place call `refresh()` so the transposed copy follows."""

    def __init__(self, M, symmetric=False, hermitian=False, ctx=None):
        import torch
        if not _is_sparse(M):
            raise _lib.B2OError("LinearOperator(M): sparse matrices must use torch's sparse_csc or sparse_csr layout "
                                "(got %s); use M.to_sparse_csc()" % getattr(M, "layout", type(M)))
        if not M.is_cuda:
            raise _lib.B2OError("LinearOperator(M): M must be a torch CUDA tensor (no CPU fallback)")
        if M.dim() != 2:
            raise LinearOperatorException("LinearOperator(M) needs a matrix")
        self.ctx = ctx or default_context(M.device.index)
        if self.ctx.device != M.device.index:
            raise _lib.B2OError("LinearOperator(M): matrix lives on cuda:%d, context on cuda:%d" % (M.device.index, self.ctx.device))
        nrow, ncol = int(M.shape[0]), int(M.shape[1])
        if M.layout == torch.sparse_csc:
            fmt, ptr, idx = 0, M.ccol_indices(), M.row_indices()
        else:
            fmt, ptr, idx = 1, M.crow_indices(), M.col_indices()
        vals = M.values()
        if not vals.is_contiguous():
            raise _lib.B2OError("LinearOperator(M): the value array must be contiguous")
        self.M, self._vals = M, vals                     # aliased, and kept alive for the handle
        self._code = _dtype_code(vals)
        ptr1 = (ptr.to("cpu", torch.int64) + 1).contiguous()      # the reference's 1-based arrays
        idx1 = (idx.to("cpu", torch.int64) + 1).contiguous()
        nnz = int(vals.numel())
        self._h = ctypes.c_void_p()
        _lib.check(self.ctx.lib.b2o_sparse_create(self.ctx.handle, self._code, fmt, nrow, ncol, nnz, ctypes.c_void_p(ptr1.data_ptr()),
                                                  ctypes.c_void_p(idx1.data_ptr()), ctypes.c_void_p(vals.data_ptr() if nnz else 0),
                                                  ctypes.byref(self._h)))

        def prod_(res, v, a, b):         # mul!(res, M, v, α, β)                              constructors.jl:25
            self._run(0, res, v, a, b)

        def tprod_(res, u, a, b):        # mul!(res, transpose(M), u, α, β) (≡ adjoint, real T) constructors.jl:26-27
            self._run(1, res, u, a, b)

        super().__init__(vals.dtype, nrow, ncol, symmetric, hermitian, prod_, tprod_, tprod_,
                         S=Storage("cuda", self.ctx.device, dtype=vals.dtype))

    def _vec(self, t, what):
        import torch
        if not isinstance(t, torch.Tensor) or not t.is_cuda:
            raise _lib.B2OError("%s must be a torch CUDA tensor (no CPU fallback)" % what)
        if t.dtype != self._vals.dtype:
            raise _lib.B2OError("%s must be %s like the matrix (got %s)" % (what, self._vals.dtype, t.dtype))
        if t.dim() != 1 or (t.numel() > 1 and t.stride(0) != 1):
            raise _lib.B2OError("%s must be a unit-stride 1-D tensor" % what)
        return ctypes.c_void_p(t.data_ptr())

    def _run(self, trans, res, v, alpha, beta):
        _lib.check(self.ctx.lib.b2o_sparse_apply(self._h, int(trans), self._vec(res, "res"), res.shape[0], self._vec(v, "v"),
                                                 v.shape[0], float(alpha), float(beta)))

    def refresh(self):
        """the values were changed in place: re-gather the transposed copy"""
        _lib.check(self.ctx.lib.b2o_sparse_refresh(self._h))

    def apply_bytes(self, trans=False, beta=0.0):
        out = ctypes.c_double()
        _lib.check(self.ctx.lib.b2o_sparse_apply_bytes(self._h, int(bool(trans)), float(beta), ctypes.byref(out)))
        return out.value

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self.ctx.handle:
                self.ctx.lib.b2o_sparse_destroy(self._h)
                self._h = None
        except Exception:
            pass


def matrix_operator(M, **kw):
    """LinearOperator(M): dense (strided) or sparse (CSC / CSR) matrix leaf"""
    import torch
    if isinstance(M, torch.Tensor) and M.layout != torch.strided:
        return SparseMatrixOperator(M, **kw)
    return DenseMatrixOperator(M, **kw)


def as_operator(x):
    """the reference's `LinearOperator(M)` promotion of a matrix argument (operations.jl:159-160, cat.jl:3-5)"""
    return matrix_operator(x) if is_matrix(x) else x
