"""Host-side mirror of the reference's operator interface for the apply path.

Mirrors (behaviour, names, argument meaning, error behaviour -- not code):
  src/abstract.jl   : AbstractLinearOperator / LinearOperator fields, counters, size, Matrix(op)
  src/operations.jl : mul!(res, op, v, α, β), prod3!, unary/binary operator algebra
  src/adjtrans.jl   : Adjoint/Transpose/Conjugate wrappers and their mul! inference rules

Julia names ending in `!` end in `_` here (mul! -> mul_, push! -> push_, reset! -> reset_).
Vectors are torch CUDA tensors for the library's own operators; the dispatch layer itself is
array-agnostic (user closures over numpy arrays work too, which is how the CPU tests drive it)."""
import inspect

import numpy as np

from ._lib import LinearOperatorException


# ------------------------------------------------------------------ vector helpers (array-agnostic)
def _is_torch(v):
    return type(v).__module__.startswith("torch")


def similar(v, n=None, dtype=None):
    """similar(v, T, n): uninitialised vector of the same array family."""
    n = v.shape[0] if n is None else int(n)
    if _is_torch(v):
        import torch
        return torch.empty(n, dtype=dtype or v.dtype, device=v.device)
    return np.empty(n, dtype=dtype or v.dtype)


def _is_complex(v):
    if _is_torch(v):
        return v.is_complex()
    return np.iscomplexobj(v)


def _device_conj(dst, src):
    """conj on the device through the library (b2o_conj): complex128 unit-stride CUDA tensors only; False otherwise"""
    import torch
    for t in (dst, src):
        if not (t.is_cuda and t.dtype == torch.complex128 and t.dim() == 1 and (t.numel() <= 1 or t.stride(0) == 1) and not t.is_conj()):
            return False
    from . import _lib
    from .context import default_context
    import ctypes
    ctx = default_context(dst.device.index)
    _lib.check(ctx.lib.b2o_conj(ctx.handle, ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(src.data_ptr()), dst.shape[0]))
    return True


def _conj_inplace(res):
    """conj!(res)  (src/adjtrans.jl:128,136)"""
    if _is_complex(res):
        if _is_torch(res):
            if not _device_conj(res, res):
                res.copy_(res.conj().resolve_conj())
        else:
            np.conjugate(res, out=res)


def _conj_copy(v):
    """conj.(v)  (src/adjtrans.jl:129); real vectors are passed through without a copy"""
    if not _is_complex(v):
        return v
    if _is_torch(v):
        out = similar(v)
        if _device_conj(out, v):
            return out
        return v.conj().resolve_conj()
    return np.conjugate(v)


def _conj_scalar(x):
    return x.conjugate() if isinstance(x, complex) else x


def get_nargs(f):
    """number of positional parameters of a closure (has_args5 <=> 4: res, v, α, β)."""
    try:
        ps = [p for p in inspect.signature(f).parameters.values()
              if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        if any(p.kind == p.VAR_POSITIONAL for p in inspect.signature(f).parameters.values()):
            return 4
        return len(ps)
    except (TypeError, ValueError):
        return 4


class Storage:
    """storage_type(op) (src/abstract.jl:176-184): how temporaries (Mv, Mtu, vtmp) are allocated."""

    def __init__(self, kind="cuda", device=0, dtype=None):
        self.kind, self.device, self.dtype = kind, device, dtype

    def alloc(self, n, zero=False):
        if self.kind == "cuda":
            import torch
            dt = self.dtype or torch.float64
            f = torch.zeros if zero else torch.empty
            return f(int(n), dtype=dt, device="cuda:%d" % self.device)
        dt = self.dtype or np.float64
        return (np.zeros if zero else np.empty)(int(n), dtype=dt)

    def __eq__(self, o):
        return isinstance(o, Storage) and (self.kind, self.device) == (o.kind, o.device)

    def __repr__(self):
        return "Storage(%s)" % (self.kind if self.kind != "cuda" else "cuda:%d" % self.device)


def _dtype_rank(dt):
    """(bits, dtype) with None standing for the default Float64"""
    name = "float64" if dt is None else str(dt).replace("torch.", "")
    # complex types rank above every real type: promote_type(Vector{Float64}, Vector{ComplexF64}) == Vector{ComplexF64}
    return {"bfloat16": 16, "float16": 16, "float32": 32, "float64": 64, "complex64": 96, "complex128": 128}.get(name, 64), dt


def promote_storage(a, b):
    """promote_type(storage_type(op1), storage_type(op2)) must be concrete (src/operations.jl:138-147); the element type
    promotes to the wider one (promote_type(Vector{Float32}, Vector{Float64}) == Vector{Float64})."""
    if a == b:
        ra, rb = _dtype_rank(a.dtype), _dtype_rank(b.dtype)
        if ra[0] == rb[0]:
            return a
        return a if ra[0] > rb[0] else b
    raise LinearOperatorException(
        "storage types %r and %r cannot be promoted to a concrete type. "
        "Ensure both operators use compatible storage types (e.g., both GPU or both CPU)." % (a, b))


def _is_matrix(x):
    """AbstractMatrix arguments of the reference's methods (promoted with LinearOperator(M)): a 2-D torch tensor"""
    return _is_torch(x) and hasattr(x, "dim") and x.dim() == 2


def _as_op(x):
    if _is_matrix(x):
        from .constructors import matrix_operator
        return matrix_operator(x)
    return x


# ------------------------------------------------------------------ types
class AbstractLinearOperator:
    """AbstractLinearOperator{T} (src/abstract.jl:30).  Subclasses carry the fields
    nrow, ncol, symmetric, hermitian, prod_, tprod_, ctprod_, nprod, ntprod, nctprod."""

    # -- size / traits (src/abstract.jl:203-253)
    @property
    def shape(self):
        return size(self)

    def __matmul__(self, other):
        return self.__mul__(other)

    # -- algebra (src/operations.jl:100-234)
    def __neg__(self):
        return neg(self)

    def __pos__(self):
        return self

    def __mul__(self, other):
        if isinstance(other, AbstractLinearOperator):
            return op_times_op(self, other)
        if _is_matrix(other):
            return op_times_op(self, _as_op(other))              # op * M = op * LinearOperator(M)   operations.jl:160
        if isinstance(other, (int, float, complex, np.number)):
            return op_times_scalar(self, other)
        if hasattr(other, "shape") and len(other.shape) == 1:
            return apply(self, other)
        return NotImplemented

    def __rmul__(self, other):
        if isinstance(other, (int, float, complex, np.number)):
            return op_times_scalar(self, other)
        if _is_matrix(other):
            return op_times_op(_as_op(other), self)              # M * op = LinearOperator(M) * op   operations.jl:159
        return NotImplemented

    def __rmatmul__(self, other):
        return self.__rmul__(other)

    def __truediv__(self, x):
        return op_times_scalar(self, 1.0 / x)                 # op * (one(T) / x)      operations.jl:183

    def __add__(self, other):
        if isinstance(other, AbstractLinearOperator):
            return op_plus_op(self, other)
        if _is_matrix(other):
            return op_plus_op(self, _as_op(other))               # op + M                            operations.jl:219
        if isinstance(other, (int, float, complex, np.number)):
            from .special_operators import opOnes
            return op_plus_op(self, op_times_scalar(opOnes(self.nrow, self.ncol, like=self), other))  # :222
        return NotImplemented

    def __radd__(self, other):
        if _is_matrix(other):
            return op_plus_op(_as_op(other), self)               # M + op                            operations.jl:218
        if isinstance(other, (int, float, complex, np.number)):
            from .special_operators import opOnes
            return op_plus_op(op_times_scalar(opOnes(self.nrow, self.ncol, like=self), other), self)  # :223
        return NotImplemented

    def __sub__(self, other):
        if isinstance(other, AbstractLinearOperator):
            return op_plus_op(self, neg(other))                                                     # :226
        if _is_matrix(other):
            return op_plus_op(self, neg(_as_op(other)))                                             # :230
        if isinstance(other, (int, float, complex, np.number)):
            return self + (-other)                                                                   # :233
        return NotImplemented

    def __rsub__(self, other):
        if _is_matrix(other):
            return op_plus_op(_as_op(other), neg(self))                                             # :229
        if isinstance(other, (int, float, complex, np.number)):
            return other + neg(self)                                                                 # :234
        return NotImplemented

    def __getitem__(self, rc):
        from .special_operators import getindex
        return getindex(self, rc[0], rc[1])

    @property
    def T(self):
        return transpose(self)

    @property
    def H(self):
        return adjoint(self)

    def conj(self):
        return conj(self)

    def __repr__(self):
        nrow, ncol = size(self)
        return ("Linear operator\n  nrow: %s\n  ncol: %d\n  eltype: %s\n  symmetric: %s\n  hermitian: %s\n"
                "  nprod:   %d\n  ntprod:  %d\n  nctprod: %d\n\n" %
                (nrow, ncol, eltype(self), issymmetric(self), ishermitian(self), nprod(self), ntprod(self),
                 nctprod(self)))


class LinearOperator(AbstractLinearOperator):
    """LinearOperator{T,S}(nrow, ncol, symmetric, hermitian, prod!, tprod!, ctprod!) -- src/abstract.jl:46-143,
    src/constructors.jl:99-111.  Closures are `(res, v, α, β)`; `(res, v)` closures are emulated (prod3!).
    `LinearOperator(M; symmetric, hermitian)` with a matrix M (src/constructors.jl:15-29) builds the dense-matrix leaf
    (constructors.DenseMatrixOperator)."""

    def __new__(cls, *args, **kw):
        if cls is LinearOperator and args and _is_matrix(args[0]):
            from .constructors import DenseMatrixOperator, SparseMatrixOperator
            import torch
            return object.__new__(DenseMatrixOperator if args[0].layout == torch.strided else SparseMatrixOperator)
        return object.__new__(cls)

    def __init__(self, T, nrow, ncol, symmetric, hermitian, prod_, tprod_=None, ctprod_=None, S=None):
        self.eltype = T
        self.nrow, self.ncol = int(nrow), int(ncol)
        self.symmetric, self.hermitian = bool(symmetric), bool(hermitian)
        self.prod_, self.tprod_, self.ctprod_ = prod_, tprod_, ctprod_
        self.nprod = self.ntprod = self.nctprod = 0
        self.S = S if S is not None else Storage("numpy")
        self.Mv = None    # allocated lazily for 3-arg closures (allocate_vectors_args3!)
        self.Mtu = None
        self.args5 = get_nargs(prod_) == 4


def size(op, d=None):
    if isinstance(op, (AdjointLinearOperator, TransposeLinearOperator)):
        s = size(op.parent)[::-1]
    elif isinstance(op, ConjugateLinearOperator):
        s = size(op.parent)
    else:
        s = (op.nrow, op.ncol)
    if d is None:
        return s
    if d in (1, 2):
        return s[d - 1]
    raise LinearOperatorException("Linear operators only have 2 dimensions for now")


def eltype(op):
    return op.parent.eltype if isinstance(op, _Wrapper) else op.eltype


def issymmetric(op):
    return issymmetric(op.parent) if isinstance(op, _Wrapper) else op.symmetric


def ishermitian(op):
    return ishermitian(op.parent) if isinstance(op, _Wrapper) else op.hermitian


def storage_type(op):
    return storage_type(op.parent) if isinstance(op, _Wrapper) else op.S


def has_args5(op):
    return has_args5(op.parent) if isinstance(op, _Wrapper) else get_nargs(op.prod_) == 4


def isallocated5(op):
    if isinstance(op, _Wrapper):
        return isallocated5(op.parent)
    if getattr(op, "always_allocated5", False):
        return True
    return not (op.Mv is None or op.Mtu is None)


def allocate_vectors_args3_(op):
    """src/operations.jl:3-8"""
    if isinstance(op, _Wrapper):
        return allocate_vectors_args3_(op.parent)
    S = storage_type(op)
    op.Mv = S.alloc(op.nrow)
    op.Mtu = op.Mv if op.nrow == op.ncol else S.alloc(op.ncol)
    return op


# counters (src/abstract.jl:147-153, src/adjtrans.jl:46-58)
def nprod(op):
    if isinstance(op, AdjointLinearOperator):
        return nctprod(op.parent)
    if isinstance(op, TransposeLinearOperator):
        return ntprod(op.parent)
    if isinstance(op, ConjugateLinearOperator):
        return nprod(op.parent)
    return op.nprod


def ntprod(op):
    if isinstance(op, (AdjointLinearOperator, TransposeLinearOperator)):
        return nprod(op.parent)
    if isinstance(op, ConjugateLinearOperator):
        return ntprod(op.parent)
    return op.ntprod


def nctprod(op):
    if isinstance(op, (AdjointLinearOperator, TransposeLinearOperator)):
        return nprod(op.parent)
    if isinstance(op, ConjugateLinearOperator):
        return nctprod(op.parent)
    return op.nctprod


def reset_(op):
    """reset!(op): reset the product counters (src/abstract.jl:191-196); QN operators override."""
    if hasattr(op, "_reset_state"):
        op._reset_state()
    op.nprod = op.ntprod = op.nctprod = 0
    return op


# ------------------------------------------------------------------ wrappers (src/adjtrans.jl:7-44)
class _Wrapper(AbstractLinearOperator):
    def __init__(self, parent):
        self.parent = parent

    @property
    def nrow(self):
        return size(self)[0]

    @property
    def ncol(self):
        return size(self)[1]


class AdjointLinearOperator(_Wrapper):
    def __repr__(self):
        return "Adjoint of the following LinearOperator:\n" + repr(self.parent)


class TransposeLinearOperator(_Wrapper):
    def __repr__(self):
        return "Transpose of the following LinearOperator:\n" + repr(self.parent)


class ConjugateLinearOperator(_Wrapper):
    def __repr__(self):
        return "Conjugate of the following LinearOperator:\n" + repr(self.parent)


def adjoint(A):
    if isinstance(A, AdjointLinearOperator):
        return A.parent
    if isinstance(A, ConjugateLinearOperator):
        return transpose(A.parent)
    if isinstance(A, TransposeLinearOperator):
        return conj(A.parent)
    if getattr(A, "is_identity_singleton", False):
        return A
    return AdjointLinearOperator(A)


def transpose(A):
    if isinstance(A, TransposeLinearOperator):
        return A.parent
    if isinstance(A, AdjointLinearOperator):
        return conj(A.parent)
    if isinstance(A, ConjugateLinearOperator):
        return adjoint(A.parent)
    if getattr(A, "is_identity_singleton", False):
        return A
    return TransposeLinearOperator(A)


def conj(A):
    if isinstance(A, ConjugateLinearOperator):
        return A.parent
    if isinstance(A, AdjointLinearOperator):
        return transpose(A.parent)
    if isinstance(A, TransposeLinearOperator):
        return adjoint(A.parent)
    if getattr(A, "is_identity_singleton", False):
        return A
    return ConjugateLinearOperator(A)


# ------------------------------------------------------------------ apply dispatch
def prod3_(res, prod_, v, alpha, beta, Mv):
    """3-arg closure emulating the 5-arg one (src/operations.jl:10-20)."""
    if beta == 0:
        prod_(res, v)
        if alpha != 1:
            res *= alpha
    else:
        prod_(Mv, v)
        res[:] = alpha * Mv + beta * res


def _call(res, closure, v, alpha, beta, owner, use_Mtu):
    if get_nargs(closure) == 4:
        return closure(res, v, alpha, beta)
    buf = owner.Mtu if use_Mtu else owner.Mv
    if not (beta == 0 or buf is not None):
        allocate_vectors_args3_(owner)
        buf = owner.Mtu if use_Mtu else owner.Mv
    return prod3_(res, closure, v, alpha, beta, buf)


def mul_(res, op, v, alpha=None, beta=None):
    """mul!(res, op, v[, α, β]) -- src/operations.jl:22-40 and the wrapper methods of src/adjtrans.jl.
    3-arg form: α = one(eltype(v)), β = zero(eltype(v)) (Q6)."""
    if alpha is None:
        alpha, beta = 1.0, 0.0
    if getattr(v, "ndim", 1) == 2:
        return _mul_matrix(res, op, v, alpha, beta)
    if isinstance(op, AdjointLinearOperator):
        return _mul_adjoint(res, op, v, alpha, beta)
    if isinstance(op, TransposeLinearOperator):
        return _mul_transpose(res, op, v, alpha, beta)
    if isinstance(op, ConjugateLinearOperator):
        p = op.parent
        mul_(res, p, _conj_copy(v), alpha, beta)              # adjtrans.jl:226-249
        _conj_inplace(res)
        return res
    if not (v.shape[0] == size(op, 2) and res.shape[0] == size(op, 1)):
        raise LinearOperatorException("shape mismatch")
    op.nprod += 1
    _call(res, op.prod_, v, alpha, beta, op, False)
    return res


def _mul_matrix(res, op, m, alpha, beta):
    """matrix right-hand side: column by column through the vector path (src/operations.jl:34-36 hands the
    matrix to the closure; the library's closures are vector kernels, so the mirror loops)."""
    if not (m.shape[0] == size(op, 2) and res.shape[0] == size(op, 1) and m.shape[1] == res.shape[1]):
        raise LinearOperatorException("shape mismatch")
    block = getattr(op, "_mul_block", None)
    if block is not None and block(res, m, alpha, beta):
        return res
    # the reference's matrix mul! never touches the product counters (no increase_nprod! at src/operations.jl:34-36)
    saved = tuple(getattr(op, c, None) for c in ("nprod", "ntprod", "nctprod"))
    for j in range(m.shape[1]):
        mul_(res[:, j], op, m[:, j], alpha, beta)
    for c, v in zip(("nprod", "ntprod", "nctprod"), saved):
        if v is not None:
            try:
                setattr(op, c, v)
            except AttributeError:
                pass
    return res


def _mul_adjoint(res, op, v, alpha, beta):
    """src/adjtrans.jl:90-137"""
    p = op.parent
    if not (v.shape[0] == size(p, 1) and res.shape[0] == size(p, 2)):
        raise LinearOperatorException("shape mismatch")
    if ishermitian(p):
        return mul_(res, p, v, alpha, beta)
    ctprod_ = p.ctprod_
    if ctprod_ is not None:
        p.nctprod += 1
        _call(res, ctprod_, v, alpha, beta, p, True)
        return res
    tprod_ = p.tprod_
    increment_tprod = True
    if tprod_ is None:
        if issymmetric(p):
            increment_tprod = False
            tprod_ = p.prod_
        else:
            raise LinearOperatorException("unable to infer conjugate transpose operator")
    if increment_tprod:
        p.ntprod += 1
    else:
        p.nprod += 1
    _conj_inplace(res)
    _call(res, tprod_, _conj_copy(v), _conj_scalar(alpha), _conj_scalar(beta), p, True)
    _conj_inplace(res)
    return res


def _mul_transpose(res, op, v, alpha, beta):
    """src/adjtrans.jl:158-205"""
    p = op.parent
    if not (v.shape[0] == size(p, 1) and res.shape[0] == size(p, 2)):
        raise LinearOperatorException("shape mismatch")
    if issymmetric(p):
        return mul_(res, p, v, alpha, beta)
    tprod_ = p.tprod_
    if tprod_ is not None:
        p.ntprod += 1
        _call(res, tprod_, v, alpha, beta, p, True)
        return res
    increment_ctprod = True
    ctprod_ = p.ctprod_
    if ctprod_ is None:
        if ishermitian(p):
            increment_ctprod = False
            ctprod_ = p.prod_
        else:
            raise LinearOperatorException("unable to infer transpose operator")
    if increment_ctprod:
        p.nctprod += 1
    else:
        p.nprod += 1
    _conj_inplace(res)
    _call(res, ctprod_, _conj_copy(v), _conj_scalar(alpha), _conj_scalar(beta), p, True)
    _conj_inplace(res)
    return res


def apply(op, v):
    """op * v (src/operations.jl:43-48): the result vector is allocated uninitialised."""
    nrow, _ = size(op)
    res = similar(v, nrow)
    mul_(res, op, v)
    return res


def Matrix(op, like=None):
    """Matrix(op): materialise through ncol unit-vector applies (src/abstract.jl:282-292)."""
    m, n = size(op)
    S = storage_type(op)
    if like is not None:
        ei = similar(like, n)
        ei[:] = 0
    else:
        ei = S.alloc(n, zero=True)
    cols = []
    for i in range(n):
        ei[i] = 1
        cols.append(apply(op, ei))
        ei[i] = 0
    if _is_torch(ei):
        import torch
        return torch.stack(cols, dim=1) if cols else torch.empty((m, 0), dtype=ei.dtype, device=ei.device)
    return np.stack(cols, axis=1) if cols else np.empty((m, 0))


# ------------------------------------------------------------------ operator algebra (src/operations.jl:100-234)
def _promote_eltype(a, b):
    """promote_type(T1, T2) (src/operations.jl:139-147): e.g. an Int64 opRestriction times a Float64 operator is Float64."""
    if a == b or b is None:
        return a
    if a is None:
        return b
    try:
        import torch
        if isinstance(a, torch.dtype) and isinstance(b, torch.dtype):
            return torch.promote_types(a, b)
        import numpy as np
        return np.promote_types(a, b)
    except Exception:
        return a


def neg(op):
    """-op  (operations.jl:102-115); wrappers push the sign through (adjtrans.jl:263-265)."""
    if isinstance(op, AdjointLinearOperator):
        return adjoint(neg(op.parent))
    if isinstance(op, TransposeLinearOperator):
        return transpose(neg(op.parent))
    if isinstance(op, ConjugateLinearOperator):
        return conj(neg(op.parent))
    prod_ = lambda res, v, a, b: mul_(res, op, v, -a, b)
    tprod_ = lambda res, u, a, b: mul_(res, transpose(op), u, -a, b)
    ctprod_ = lambda res, w, a, b: mul_(res, adjoint(op), w, -a, b)
    out = LinearOperator(eltype(op), op.nrow, op.ncol, op.symmetric, op.hermitian, prod_, tprod_, ctprod_,
                         S=storage_type(op))
    out._expr = ("neg", op)
    return out


def prod_op_(res, op1, op2, vtmp, v, alpha, beta):
    """src/operations.jl:117-128"""
    mul_(vtmp, op2, v)
    mul_(res, op1, vtmp, alpha, beta)


def op_times_op(op1, op2):
    """op1 * op2 (operations.jl:131-156)"""
    if getattr(op1, "is_identity_singleton", False):
        return op2
    if getattr(op2, "is_identity_singleton", False):
        return op1
    m1, n1 = size(op1)
    m2, n2 = size(op2)
    if m2 != n1:
        raise LinearOperatorException("shape mismatch")
    S = promote_storage(storage_type(op1), storage_type(op2))
    vtmp = S.alloc(m2, zero=True)
    utmp = S.alloc(n1, zero=True)
    wtmp = S.alloc(n1, zero=True)
    prod_ = lambda res, v, a, b: prod_op_(res, op1, op2, vtmp, v, a, b)
    tprod_ = lambda res, u, a, b: prod_op_(res, transpose(op2), transpose(op1), utmp, u, a, b)
    ctprod_ = lambda res, w, a, b: prod_op_(res, adjoint(op2), adjoint(op1), wtmp, w, a, b)
    out = LinearOperator(_promote_eltype(eltype(op1), eltype(op2)), m1, n2, False, False, prod_, tprod_, ctprod_, S=S)
    out._expr = ("prod", op1, op2)
    return out


def op_times_scalar(op, x):
    """op * x, x * op (operations.jl:163-181); wrappers: adjtrans.jl:267-273"""
    if isinstance(op, AdjointLinearOperator):
        return adjoint(op_times_scalar(op.parent, _conj_scalar(x)))
    if isinstance(op, TransposeLinearOperator):
        return transpose(op_times_scalar(op.parent, x))
    if isinstance(op, ConjugateLinearOperator):
        return conj(op_times_scalar(op.parent, _conj_scalar(x)))
    prod_ = lambda res, v, a, b: mul_(res, op, v, x * a, b)
    tprod_ = lambda res, u, a, b: mul_(res, transpose(op), u, x * a, b)
    ctprod_ = lambda res, w, a, b: mul_(res, adjoint(op), w, _conj_scalar(x) * a, b)
    isreal = not isinstance(x, complex) or x.imag == 0
    out = LinearOperator(eltype(op), op.nrow, op.ncol, op.symmetric, op.hermitian and isreal, prod_, tprod_,
                         ctprod_, S=storage_type(op))
    out._expr = ("scale", op, x)
    return out


def sum_prod_(res, op1, op2, v, alpha, beta):
    """src/operations.jl:187-197"""
    mul_(res, op1, v, alpha, beta)
    mul_(res, op2, v, alpha, 1.0)


def op_plus_op(op1, op2):
    """op1 + op2 (operations.jl:199-215)"""
    m1, n1 = size(op1)
    m2, n2 = size(op2)
    if m1 != m2 or n1 != n2:
        raise LinearOperatorException("shape mismatch")
    prod_ = lambda res, v, a, b: sum_prod_(res, op1, op2, v, a, b)
    tprod_ = lambda res, u, a, b: sum_prod_(res, transpose(op1), transpose(op2), u, a, b)
    ctprod_ = lambda res, w, a, b: sum_prod_(res, adjoint(op1), adjoint(op2), w, a, b)
    symm = issymmetric(op1) and issymmetric(op2)
    herm = ishermitian(op1) and ishermitian(op2)
    S = promote_storage(storage_type(op1), storage_type(op2))
    out = LinearOperator(_promote_eltype(eltype(op1), eltype(op2)), m1, n1, symm, herm, prod_, tprod_, ctprod_, S=S)
    out._expr = ("sum", op1, op2)
    return out


def Hermitian(op):
    """src/abstract.jl:231-235"""
    if size(op, 1) != size(op, 2):
        raise LinearOperatorException("Operator is not square")
    if ishermitian(op):
        return op
    return (op + adjoint(op)) / 2


def Symmetric(op):
    """src/abstract.jl:249-253"""
    if size(op, 1) != size(op, 2):
        raise LinearOperatorException("Operator is not square")
    if issymmetric(op):
        return op
    return (op + transpose(op)) / 2
