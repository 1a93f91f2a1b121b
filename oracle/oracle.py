"""CPU ORACLE (test infrastructure, NOT the product) -- Python face of oracle/b2o_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
Heavy loops are in C (b2o_oracle.c, long-double reductions); operator composition (the closure tree of
src/operations.jl / src/adjtrans.jl / src/cat.jl) is restated here over numpy so that composed chains can
be checked as well.  See b2o_oracle.h for the parity-pinning statement."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libb2o_oracle.so")


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("b2o_oracle.c", "b2o_oracle.h", "Makefile")]
    if force or not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in src):
        subprocess.run(["make", "-C", HERE, "-s", "-B"], check=True)
    return SO


_lib = None
c_dp = ctypes.POINTER(ctypes.c_double)
FETCH_FN = ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int)


def usable_cpus():
    """CPUs this process may really use: affinity mask capped by the cgroup CPU quota (containers)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            n = max(1, min(n, int(-(-int(q) // int(per)))))
    except Exception:
        pass
    return n


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(SO)
        vp, d, i64, i32 = ctypes.c_void_p, ctypes.c_double, ctypes.c_int64, ctypes.c_int
        sig = {
            "orc_set_mode": (None, [i32, i32]), "orc_max_threads": (i32, []),
            "orc_fill_uniform": (None, [vp, i64, ctypes.c_uint64, d, d]),
            "orc_dot": (d, [vp, vp, i64]), "orc_sum": (d, [vp, i64]), "orc_nrm2": (d, [vp, i64]),
            "orc_eye": (None, [vp, i64, vp, i64, d, d, i64]), "orc_ones": (None, [vp, i64, vp, i64, d, d]),
            "orc_zeros": (None, [vp, i64, d, d]), "orc_diag_square": (None, [vp, vp, vp, i64, d, d]),
            "orc_diag_rect": (None, [vp, i64, vp, vp, d, d, i64]), "orc_householder": (None, [vp, vp, vp, i64, d, d]),
            "orc_cdiag": (None, [vp, vp, vp, i64, i32, vp, vp]), "orc_chouseholder": (None, [vp, vp, vp, i64, vp, vp]),
            "orc_conj": (None, [vp, vp, i64]),
            "orc_restrict": (None, [vp, vp, i64, vp]), "orc_extend": (None, [vp, i64, vp, i64, vp]),
            "orc_lbfgs_create": (vp, [i64, i32, i32, i32, d, d, i32]), "orc_lbfgs_destroy": (None, [vp]),
            "orc_lbfgs_apply": (None, [vp, vp, vp, d, d]), "orc_lbfgs_push": (i32, [vp, vp, vp]),
            "orc_lbfgs_push_damped_fwd": (i32, [vp, vp, vp, vp]),
            "orc_lbfgs_push_damped_inv": (i32, [vp, vp, vp, d, vp, vp]), "orc_lbfgs_diag": (i32, [vp, vp]),
            "orc_lbfgs_reset": (None, [vp]), "orc_lbfgs_col": (c_dp, [vp, i32, i32]), "orc_lbfgs_ys": (c_dp, [vp]),
            "orc_lbfgs_gamma": (d, [vp]), "orc_lbfgs_set_gamma": (None, [vp, d]), "orc_lbfgs_insert": (i32, [vp]),
            "orc_lbfgs_set_insert": (None, [vp, i32]), "orc_lbfgs_opnorm_upper_bound": (d, [vp]),
            "orc_lbfgs_solve_shifted": (i32, [vp, vp, vp, d]),
            "orc_lbfgs_create_external": (vp, [i64, i32, i32, i32, FETCH_FN, vp]),
            "orc_lbfgs_last_dots": (c_dp, [vp, ctypes.POINTER(i32)]),
            "orc_lsr1_create": (vp, [i64, i32, i32]), "orc_lsr1_destroy": (None, [vp]),
            "orc_lsr1_apply": (None, [vp, vp, vp, d, d]), "orc_lsr1_push": (i32, [vp, vp, vp]),
            "orc_lsr1_diag": (None, [vp, vp]), "orc_lsr1_reset": (None, [vp]), "orc_lsr1_col": (c_dp, [vp, i32, i32]),
            "orc_lsr1_ys": (c_dp, [vp]), "orc_lsr1_as": (c_dp, [vp]), "orc_lsr1_gamma": (d, [vp]),
            "orc_lsr1_set_gamma": (None, [vp, d]), "orc_lsr1_insert": (i32, [vp]),
            "orc_lsr1_set_insert": (None, [vp, i32]), "orc_lsr1_opnorm_upper_bound": (d, [vp]),
            "orc_diagqn_push": (i32, [i32, vp, vp, vp, i64]),
            "orc_kron": (None, [vp, vp, i64, i64, vp, i64, i64, vp, d, d, i32]),
            "orc_gemv": (None, [vp, vp, i64, i64, i64, vp, d, d, i32]),
            "orc_gemv_f32": (None, [vp, vp, i64, i64, i64, vp, ctypes.c_float, ctypes.c_float, i32]),
            "orc_spmv_csc": (None, [vp, i64, i64, vp, vp, vp, vp, d, d, i32]),
            "orc_spmv_csc_f32": (None, [vp, i64, i64, vp, vp, vp, vp, ctypes.c_float, ctypes.c_float, i32]),
            "orc_f32_to_bf16": (ctypes.c_uint16, [ctypes.c_float]), "orc_bf16_to_f32": (ctypes.c_float, [ctypes.c_uint16]),
        }
        for name, (res, args) in sig.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _f64(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def set_mode(accurate=True, threads=1):
    lib().orc_set_mode(int(accurate), int(threads))


def max_threads():
    return max(1, min(lib().orc_max_threads(), usable_cpus()))


def uniform(n, seed, lo=0.0, hi=1.0):
    x = np.empty(int(n), dtype=np.float64)
    lib().orc_fill_uniform(_p(x), int(n), int(seed), float(lo), float(hi))
    return x


def dot(a, b):
    return lib().orc_dot(_p(a), _p(b), a.shape[0])


# ---------------------------------------------------------------- leaf closures (in place on res)
def eye_(res, v, alpha, beta, n_min):
    lib().orc_eye(_p(res), res.shape[0], _p(v), v.shape[0], alpha, beta, int(n_min))


def ones_(res, v, alpha, beta):
    lib().orc_ones(_p(res), res.shape[0], _p(v), v.shape[0], alpha, beta)


def zeros_(res, v, alpha, beta):
    lib().orc_zeros(_p(res), res.shape[0], alpha, beta)


def diag_(res, d, v, alpha, beta, n_min=None):
    if n_min is None:
        lib().orc_diag_square(_p(res), _p(d), _p(v), res.shape[0], alpha, beta)
    else:
        lib().orc_diag_rect(_p(res), res.shape[0], _p(d), _p(v), alpha, beta, int(n_min))


def householder_(res, h, v, alpha, beta):
    lib().orc_householder(_p(res), _p(h), _p(v), res.shape[0], alpha, beta)


def _c2(z):
    return np.array([complex(z).real, complex(z).imag], dtype=np.float64)


def cdiag_(res, d, v, alpha=1.0, beta=0.0, conj_d=False):
    """complex128 mulSquareOpDiagonal! (conj_d: the ctprod! closure); res updated in place"""
    a, b = _c2(alpha), _c2(beta)
    lib().orc_cdiag(_p(res), _p(np.ascontiguousarray(d, dtype=np.complex128)), _p(np.ascontiguousarray(v, dtype=np.complex128)),
                    res.shape[0], int(conj_d), _p(a), _p(b))


def chouseholder_(res, h, v, alpha=1.0, beta=0.0):
    a, b = _c2(alpha), _c2(beta)
    lib().orc_chouseholder(_p(res), _p(np.ascontiguousarray(h, dtype=np.complex128)), _p(np.ascontiguousarray(v, dtype=np.complex128)),
                           res.shape[0], _p(a), _p(b))


def restrict_(res, idx1, v):
    idx1 = np.ascontiguousarray(idx1, dtype=np.int64)
    lib().orc_restrict(_p(res), _p(idx1), idx1.shape[0], _p(v))


def extend_(res, idx1, u):
    idx1 = np.ascontiguousarray(idx1, dtype=np.int64)
    lib().orc_extend(_p(res), res.shape[0], _p(idx1), idx1.shape[0], _p(u))


def diagqn_push(kind, d, s, y):
    """push! of DiagonalPSB (0) / DiagonalAndrei (1) / DiagonalBFGS (2) / SpectralGradient (3); d is updated in place"""
    s, y = _f64(s), _f64(y)
    if lib().orc_diagqn_push(int(kind), _p(d), _p(s), _p(y), s.shape[0]) != 0:
        raise ZeroDivisionError("Cannot update DiagonalQN operator with s=0")
    return d


def kron_(res, A, B, x, alpha=1.0, beta=0.0, trans=0):
    A = np.asfortranarray(A, dtype=np.float64)
    B = np.asfortranarray(B, dtype=np.float64)
    lib().orc_kron(_p(res), _p(A), A.shape[0], A.shape[1], _p(B), B.shape[0], B.shape[1], _p(x), alpha, beta, trans)


def gemv_(res, M, v, alpha=1.0, beta=0.0, trans=0):
    """LinearOperator(M) closures (src/constructors.jl:25-27): res = α M v + β res (trans=0) or α Mᵀ v + β res (trans=1).
    M: 2-D float64 or float32 array (any layout; passed column-major); res, v of the same dtype, res updated in place."""
    dt = np.float32 if M.dtype == np.float32 else np.float64
    Mf = np.asfortranarray(M, dtype=dt)
    v = np.ascontiguousarray(v, dtype=dt)
    assert res.dtype == dt and res.flags.c_contiguous
    f = lib().orc_gemv_f32 if dt == np.float32 else lib().orc_gemv
    f(_p(res), _p(Mf), Mf.shape[0], Mf.shape[1], max(1, Mf.shape[0]), _p(v), alpha, beta, int(trans))
    return res


def spmv_csc_(res, m, n, colptr1, rowval1, nzval, v, alpha=1.0, beta=0.0, trans=0):
    """LinearOperator(M::SparseMatrixCSC) closures: res = α M v + β res (trans=0) / α Mᵀ v + β res (trans=1).
    colptr1, rowval1: Julia's 1-based int64 arrays; nzval float64 or float32; res updated in place."""
    dt = np.float32 if nzval.dtype == np.float32 else np.float64
    colptr1 = np.ascontiguousarray(colptr1, dtype=np.int64)
    rowval1 = np.ascontiguousarray(rowval1, dtype=np.int64)
    nzval = np.ascontiguousarray(nzval, dtype=dt)
    v = np.ascontiguousarray(v, dtype=dt)
    assert res.dtype == dt and res.flags.c_contiguous and colptr1.shape[0] == n + 1
    f = lib().orc_spmv_csc_f32 if dt == np.float32 else lib().orc_spmv_csc
    f(_p(res), int(m), int(n), _p(colptr1), _p(rowval1), _p(nzval), _p(v), alpha, beta, int(trans))
    return res


def bf16_round(a):
    """round float array to bf16 (RNE) and back to float64"""
    a32 = np.asarray(a, dtype=np.float32)
    u = a32.view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    out = ((u + r) & 0xFFFF0000).astype(np.uint32).view(np.float32)
    return out.astype(np.float64)


# ---------------------------------------------------------------- QN operators
class LBFGS:
    def __init__(self, n, mem=5, scaling=True, damped=False, sigma2=0.99, sigma3=10.0, inverse=False):
        self.n, self.mem, self.inverse = int(n), max(int(mem), 1), bool(inverse)
        self.h = lib().orc_lbfgs_create(self.n, int(mem), int(scaling), int(damped), sigma2, sigma3, int(inverse))
        if not self.h or not lib().orc_lbfgs_col(self.h, 0, 0) or not (inverse or lib().orc_lbfgs_col(self.h, 2, 0)):
            raise MemoryError("oracle L-BFGS state of %d x %d doubles does not fit in host memory" % (self.n, self.mem))

    def __del__(self):
        try:
            lib().orc_lbfgs_destroy(self.h)
        except Exception:
            pass

    def apply(self, x, alpha=1.0, beta=0.0, res=None):
        x = _f64(x)
        res = np.empty(self.n) if res is None else res
        lib().orc_lbfgs_apply(self.h, _p(res), _p(x), alpha, beta)
        return res

    def push(self, s, y):
        s, y = _f64(s), _f64(y)
        r = lib().orc_lbfgs_push(self.h, _p(s), _p(y))
        if r < 0:
            raise RuntimeError("wrong push! variant")
        return bool(r)

    def push_damped_fwd(self, s, y, Bs=None):
        s, y = _f64(s), _f64(y)
        Bs = np.empty(self.n) if Bs is None else Bs
        r = lib().orc_lbfgs_push_damped_fwd(self.h, _p(s), _p(y), _p(Bs))
        if r < 0:
            raise RuntimeError("wrong push! variant")
        return bool(r)

    def push_damped_inv(self, s, y, alpha, g, Bs=None):
        """y is modified in place, as in the reference"""
        s, g = _f64(s), _f64(g)
        Bs = np.empty(self.n) if Bs is None else Bs
        r = lib().orc_lbfgs_push_damped_inv(self.h, _p(s), _p(y), alpha, _p(g), _p(Bs))
        if r < 0:
            raise RuntimeError("wrong push! variant")
        return bool(r)

    def diag(self):
        d = np.empty(self.n)
        if lib().orc_lbfgs_diag(self.h, _p(d)) < 0:
            raise RuntimeError("only the diagonal of a forward L-BFGS approximation is available")
        return d

    def solve_shifted(self, b, sigma=0.0, x=None):
        """solve_shifted_system!(x, B, b, σ)  (src/utilities.jl:207-248)"""
        b = _f64(b)
        x = np.empty(self.n) if x is None else x
        if lib().orc_lbfgs_solve_shifted(self.h, _p(x), _p(b), float(sigma)) != 0:
            raise ValueError("σ must be nonnegative")
        return x

    def reset(self):
        lib().orc_lbfgs_reset(self.h)

    def col(self, which, k0):
        p = lib().orc_lbfgs_col(self.h, "syab".index(which), int(k0))
        return np.ctypeslib.as_array(p, shape=(self.n,))

    @property
    def ys(self):
        return np.ctypeslib.as_array(lib().orc_lbfgs_ys(self.h), shape=(self.mem,))

    insert = property(lambda self: lib().orc_lbfgs_insert(self.h))
    scaling_factor = property(lambda self: lib().orc_lbfgs_gamma(self.h))
    opnorm_upper_bound = property(lambda self: lib().orc_lbfgs_opnorm_upper_bound(self.h))

    def set_state(self, insert, gamma):
        lib().orc_lbfgs_set_insert(self.h, int(insert))
        lib().orc_lbfgs_set_gamma(self.h, float(gamma))

    def matrix(self):
        return np.stack([self.apply(e) for e in np.eye(self.n)], axis=1)


class ExternalLBFGS(LBFGS):
    """Apply-only L-BFGS oracle whose columns stay with the caller (full-size parity runs: the state lives on the GPU and is
    streamed to the host one column at a time).  fetch(which, k0, slot) -> address of a host buffer holding column k0 of
    'syab'[which]; slot in {0, 1} says which of the caller's two buffers to use.  The apply code is the resident one's."""

    def __init__(self, n, mem, inverse, fetch, scaling=True):
        self.n, self.mem, self.inverse = int(n), max(int(mem), 1), bool(inverse)
        self._cb = FETCH_FN(lambda user, which, k0, slot: fetch(which, k0, slot))
        self.h = lib().orc_lbfgs_create_external(self.n, int(mem), int(scaling), int(inverse), self._cb, None)

    def last_dots(self):
        cnt = ctypes.c_int()
        p = lib().orc_lbfgs_last_dots(self.h, ctypes.byref(cnt))
        return np.array([p[i] for i in range(cnt.value)])


class LSR1:
    def __init__(self, n, mem=5, scaling=True):
        self.n, self.mem = int(n), max(int(mem), 1)
        self.h = lib().orc_lsr1_create(self.n, int(mem), int(scaling))

    def __del__(self):
        try:
            lib().orc_lsr1_destroy(self.h)
        except Exception:
            pass

    def apply(self, x, alpha=1.0, beta=0.0, res=None):
        x = _f64(x)
        res = np.empty(self.n) if res is None else res
        lib().orc_lsr1_apply(self.h, _p(res), _p(x), alpha, beta)
        return res

    def push(self, s, y):
        s, y = _f64(s), _f64(y)
        return bool(lib().orc_lsr1_push(self.h, _p(s), _p(y)))

    def diag(self):
        d = np.empty(self.n)
        lib().orc_lsr1_diag(self.h, _p(d))
        return d

    def reset(self):
        lib().orc_lsr1_reset(self.h)

    def col(self, which, k0):
        p = lib().orc_lsr1_col(self.h, "sya".index(which), int(k0))
        return np.ctypeslib.as_array(p, shape=(self.n,))

    @property
    def ys(self):
        return np.ctypeslib.as_array(lib().orc_lsr1_ys(self.h), shape=(self.mem,))

    @property
    def as_(self):
        return np.ctypeslib.as_array(lib().orc_lsr1_as(self.h), shape=(self.mem,))

    insert = property(lambda self: lib().orc_lsr1_insert(self.h))
    scaling_factor = property(lambda self: lib().orc_lsr1_gamma(self.h))
    opnorm_upper_bound = property(lambda self: lib().orc_lsr1_opnorm_upper_bound(self.h))

    def set_state(self, insert, gamma):
        lib().orc_lsr1_set_insert(self.h, int(insert))
        lib().orc_lsr1_set_gamma(self.h, float(gamma))

    def matrix(self):
        return np.stack([self.apply(e) for e in np.eye(self.n)], axis=1)


# ---------------------------------------------------------------- composition: closure tree over numpy
class Op:
    """minimal restatement of LinearOperator + mul! (src/operations.jl:22-32) for the oracle side: closures
    `(res, v, α, β)`; symmetric/hermitian shortcuts of src/adjtrans.jl for real element types."""

    def __init__(self, nrow, ncol, symmetric, hermitian, prod, tprod=None, ctprod=None):
        self.nrow, self.ncol, self.symmetric, self.hermitian = nrow, ncol, symmetric, hermitian
        self.prod, self.tprod, self.ctprod = prod, tprod, ctprod

    def mul(self, res, v, alpha=1.0, beta=0.0):
        assert v.shape[0] == self.ncol and res.shape[0] == self.nrow, "shape mismatch"
        self.prod(res, v, alpha, beta)
        return res

    def tmul(self, res, v, alpha=1.0, beta=0.0):
        """mul!(res, transpose(op), v, α, β) for real T (adjoint == transpose)"""
        assert v.shape[0] == self.nrow and res.shape[0] == self.ncol, "shape mismatch"
        if self.symmetric or self.hermitian:
            return self.mul(res, v, alpha, beta)
        f = self.tprod or self.ctprod
        assert f is not None, "unable to infer transpose operator"
        f(res, v, alpha, beta)
        return res

    def __call__(self, v):
        return self.mul(np.empty(self.nrow), v)

    @property
    def T(self):
        return Op(self.ncol, self.nrow, self.symmetric, self.hermitian, lambda r, v, a, b: self.tmul(r, v, a, b),
                  lambda r, v, a, b: self.mul(r, v, a, b), lambda r, v, a, b: self.mul(r, v, a, b))

    def __mul__(self, o):
        if isinstance(o, Op):                                   # prod_op!  src/operations.jl:117-156
            vtmp, utmp = np.zeros(o.nrow), np.zeros(self.ncol)

            def prod(res, v, a, b):
                o.mul(vtmp, v)
                self.mul(res, vtmp, a, b)

            def tprod(res, u, a, b):
                self.tmul(utmp, u)
                o.tmul(res, utmp, a, b)

            return Op(self.nrow, o.ncol, False, False, prod, tprod, tprod)
        x = float(o)                                            # op * x  :163-177
        return Op(self.nrow, self.ncol, self.symmetric, self.hermitian, lambda r, v, a, b: self.mul(r, v, x * a, b),
                  lambda r, v, a, b: self.tmul(r, v, x * a, b), lambda r, v, a, b: self.tmul(r, v, x * a, b))

    __rmul__ = __mul__

    def __neg__(self):                                          # :102-115
        return Op(self.nrow, self.ncol, self.symmetric, self.hermitian, lambda r, v, a, b: self.mul(r, v, -a, b),
                  lambda r, v, a, b: self.tmul(r, v, -a, b), lambda r, v, a, b: self.tmul(r, v, -a, b))

    def __add__(self, o):                                       # sum_prod!  :187-215
        def prod(res, v, a, b):
            self.mul(res, v, a, b)
            o.mul(res, v, a, 1.0)

        def tprod(res, v, a, b):
            self.tmul(res, v, a, b)
            o.tmul(res, v, a, 1.0)

        return Op(self.nrow, self.ncol, self.symmetric and o.symmetric, self.hermitian and o.hermitian, prod, tprod, tprod)

    def __sub__(self, o):
        return self + (-o)

    def matrix(self):
        return np.stack([self(e) for e in np.eye(self.ncol)], axis=1)


def opEye(nrow, ncol=None):
    ncol = nrow if ncol is None else ncol
    f = lambda r, v, a, b: eye_(r, v, a, b, min(nrow, ncol))
    return Op(nrow, ncol, nrow == ncol, nrow == ncol, f, f, f)


def opOnes(nrow, ncol):
    return Op(nrow, ncol, nrow == ncol, nrow == ncol, ones_, ones_, ones_)


def opZeros(nrow, ncol):
    return Op(nrow, ncol, nrow == ncol, nrow == ncol, zeros_, zeros_, zeros_)


def opDiagonal(d, nrow=None, ncol=None):
    d = _f64(d)
    if nrow is None:
        n = d.shape[0]
        f = lambda r, v, a, b: diag_(r, d, v, a, b)
        return Op(n, n, True, True, f, f, f)
    nm = min(nrow, ncol)
    f = lambda r, v, a, b: diag_(r, d, v, a, b, nm)
    return Op(nrow, ncol, False, False, f, f, f)


def opHouseholder(h):
    h = _f64(h)
    f = lambda r, v, a, b: householder_(r, h, v, a, b)
    return Op(h.shape[0], h.shape[0], True, True, f, None, f)


def opRestriction(idx1, ncol):
    idx1 = np.asarray(idx1, dtype=np.int64)
    return Op(idx1.shape[0], ncol, False, False, lambda r, v, a, b: restrict_(r, idx1, v),
              lambda r, u, a, b: extend_(r, idx1, u), lambda r, u, a, b: extend_(r, idx1, u))


def opExtension(idx1, ncol):
    return opRestriction(idx1, ncol).T


def wrap_qn(q):
    f = lambda r, v, a, b: q.apply(v, a, b, res=r)
    return Op(q.n, q.n, True, True, f, f, f)


def hcat(A, B):                                                 # src/cat.jl:7-51
    def prod(res, v, a, b):
        A.mul(res, v[:A.ncol], a, b)
        B.mul(res, v[A.ncol:], a, 1.0)

    def tprod(res, u, a, b):
        A.tmul(res[:A.ncol], u, a, b)
        B.tmul(res[A.ncol:], u, a, b)

    return Op(A.nrow, A.ncol + B.ncol, False, False, prod, tprod, tprod)


def vcat(A, B):                                                 # src/cat.jl:65-109
    def prod(res, u, a, b):
        A.mul(res[:A.nrow], u, a, b)
        B.mul(res[A.nrow:], u, a, b)

    def tprod(res, v, a, b):
        A.tmul(res, v[:A.nrow], a, b)
        B.tmul(res, v[A.nrow:], a, 1.0)

    return Op(A.nrow + B.nrow, A.ncol, False, False, prod, tprod, tprod)


def block_diagonal(*ops):                                       # src/special-operators.jl:249-294
    nrow, ncol = sum(o.nrow for o in ops), sum(o.ncol for o in ops)

    def prod(y, x, a, b):
        k = j = 0
        for o in ops:
            o.mul(y[k:k + o.nrow], x[j:j + o.ncol], a, b)
            k += o.nrow
            j += o.ncol

    def tprod(y, x, a, b):
        k = j = 0
        for o in ops:
            o.tmul(y[k:k + o.ncol], x[j:j + o.nrow], a, b)
            k += o.ncol
            j += o.nrow

    return Op(nrow, ncol, all(o.symmetric for o in ops), all(o.hermitian for o in ops), prod, tprod, tprod)
