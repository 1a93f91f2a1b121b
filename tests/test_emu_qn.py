"""CPU check of the streaming quasi-Newton kernels (csrc/b2o_stream.cuh, b2o_qn_kernels.cuh, b2o_qn_multi.cuh) under the host SIMT
emulator (tests/emu/): the product's own kernel bodies run as one block of 288 OS threads with the mbarrier / bulk-copy stand-ins of
simt_emu.h.  Checked: the TMA ring protocol with MORE and with FEWER slots than a tile has items (a hand-back bug deadlocks here, on a
CPU box, instead of hanging a GPU), the sweep bookkeeping of the two-loop recursion and of its block variant, ragged tiles, views that
are only 8-byte aligned, and that the block recursion reproduces the vector kernel bit for bit.  References: numpy restatements of
src/lbfgs.jl:117-154, :173-202 and src/lsr1.jl:89-107 (same statement order; inner products via np.dot)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMU = os.path.join(HERE, "emu")
SO = os.path.join(EMU, "_build", "libqn_emu.so")
CSRC = os.path.join(HERE, "..", "linearoperators.jl_b200", "csrc")
R = 1024
pytestmark = pytest.mark.timeout(600)


@pytest.fixture(scope="module")
def emu():
    deps = [os.path.join(EMU, f) for f in ("qn_emu.cpp", "simt_emu.h")] + [os.path.join(CSRC, f) for f in
                                                                            ("b2o_stream.cuh", "b2o_qn_kernels.cuh", "b2o_qn_multi.cuh", "b2o_shared_defs.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fvisibility=hidden",
                        "-Wl,-Bsymbolic", "-o", SO, deps[0]], check=True)
    L = ctypes.CDLL(SO)
    i32, i64, d, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
    L.emu_qn_compact.restype = i32
    L.emu_qn_compact.argtypes = [i32, i64, i64, i32, vp, vp, vp, vp, d, d, d, i32, i32]
    L.emu_qn_twoloop.restype = i32
    L.emu_qn_twoloop.argtypes = [i64, i64, i32, vp, vp, vp, vp, i64, vp, i64, i32, i32, vp, d, d, d, i32, i32]
    return L


def aligned(shape, align=16, offset=0):
    """float64 array whose data pointer is `align`-byte aligned plus `offset` bytes"""
    n = int(np.prod(shape))
    raw = np.zeros(n + 8, dtype=np.float64)
    start = (-(raw.ctypes.data // 8) % (align // 8) + offset // 8) % 8
    return raw[start:start + n].reshape(shape)


def columns(ncols, n, rng, scale=1.0):
    pitch = (n + R - 1) // R * R
    c = aligned((ncols, pitch))
    c[:, :n] = scale * rng.random((ncols, n))
    return c, pitch


def rel(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


@pytest.mark.parametrize("stages", [2, 3, 7])
@pytest.mark.parametrize("n,npairs,off", [(2500, 3, 0), (1024, 1, 8), (3073, 5, 8), (700, 0, 0)])
def test_forward_apply_ring_protocol(emu, stages, n, npairs, off):
    """qn_compact_kernel<1024, LBFGS_FWD>: src/lbfgs.jl:183-201"""
    rng = np.random.default_rng(n + stages)
    cols, pitch = columns(2 * npairs, n, rng, 0.1)
    x, res = aligned(n, offset=off), aligned(n, offset=off)
    x[:] = rng.random(n)
    r0 = rng.random(n)
    gamma = 0.7
    for alpha, beta in ((1.0, 0.0), (-0.75, 0.5)):
        res[:] = r0 if beta != 0 else np.nan
        assert emu.emu_qn_compact(0, n, pitch, 2 * npairs, cols.ctypes.data, None, x.ctypes.data, res.ctypes.data, alpha, beta, gamma, 1, stages) == 0
        q = x / gamma
        for k in range(npairs):
            a, b = cols[2 * k, :n], cols[2 * k + 1, :n]
            q = q + ((b @ x) * b - (a @ x) * a)
        ref = alpha * q + beta * r0 if beta != 0 else alpha * q
        assert rel(res, ref) <= 1e-14, (alpha, beta, rel(res, ref))


@pytest.mark.parametrize("stages", [2, 5])
def test_lsr1_apply(emu, stages):
    """qn_compact_kernel<1024, LSR1>: src/lsr1.jl:92-105"""
    rng = np.random.default_rng(stages)
    n, m = 2100, 4
    cols, pitch = columns(m, n, rng, 0.2)
    asv = aligned(m)
    asv[:] = rng.random(m) + 0.5
    x, res = aligned(n), aligned(n)
    x[:] = rng.random(n)
    r0 = rng.random(n)
    res[:] = r0
    emu.emu_qn_compact(1, n, pitch, m, cols.ctypes.data, asv.ctypes.data, x.ctypes.data, res.ctypes.data, 1.5, -0.5, 0.9, 1, stages)
    q = (1.5 * x) / 0.9 + (-0.5) * r0
    for k in range(m):
        q = q + ((1.5 * (cols[k, :n] @ x)) / asv[k]) * cols[k, :n]
    assert rel(res, q) <= 1e-14


def numpy_twoloop(S, Y, ys, x, alpha, beta, gamma, r0):
    """src/lbfgs.jl:127-153 with the pairs given newest -> oldest"""
    A = len(ys)
    q = x.copy()
    al = np.zeros(A)
    for i in range(A):
        al[i] = (S[i] @ q) / ys[i]
        q = q - al[i] * Y[i]
    q = q * gamma
    for i in range(A - 1, -1, -1):
        q = q + (al[i] - (Y[i] @ q) / ys[i]) * S[i]
    return alpha * q + beta * r0 if beta != 0 else alpha * q


def twoloop_state(n, A, rng):
    S, pitch = columns(A, n, rng)
    Y, _ = columns(A, n, rng)
    Y[:, :n] = S[:, :n] + 0.1 * Y[:, :n]
    ys = aligned(A)
    ys[:] = [S[i, :n] @ Y[i, :n] for i in range(A)]
    return S, Y, ys, pitch


@pytest.mark.parametrize("stages", [3, 7])
@pytest.mark.parametrize("n,A", [(2500, 3), (1025, 1)])
def test_twoloop_vector_kernel(emu, stages, n, A):
    rng = np.random.default_rng(10 * n + A)
    S, Y, ys, pitch = twoloop_state(n, A, rng)
    x, res, q = aligned(n), aligned(n), aligned(pitch)
    x[:] = rng.random(n)
    r0 = rng.random(n)
    for alpha, beta in ((1.0, 0.0), (2.0, -0.25)):
        res[:] = r0
        q[:] = 0
        emu.emu_qn_twoloop(n, pitch, A, S.ctypes.data, Y.ctypes.data, ys.ctypes.data, x.ctypes.data, n, res.ctypes.data, n, 0, 0, q.ctypes.data,
                           alpha, beta, 0.8, 1, stages)
        assert rel(res, numpy_twoloop(S[:, :n], Y[:, :n], ys, x, alpha, beta, 0.8, r0)) <= 1e-14


@pytest.mark.parametrize("NR,nrhs,stages", [(4, 3, 7), (4, 4, 3), (8, 8, 7), (8, 5, 3), (8, 8, 12), (4, 2, 2)])
def test_block_twoloop_equals_vector_kernel_bit_for_bit(emu, NR, nrhs, stages):
    """qn_twoloop_multi_kernel: rings with fewer slots than the nrhs + 2 items of a tile (8 right-hand sides on 7 or 3 slots) must not
    deadlock, and every column must carry the bits of the vector kernel"""
    rng = np.random.default_rng(100 * NR + nrhs + stages)
    n, A = 2300, 3
    S, Y, ys, pitch = twoloop_state(n, A, rng)
    ld = n + 1                                                # odd leading dimension: columns are only 8-byte aligned
    X, Res = aligned((nrhs, ld)), aligned((nrhs, ld))
    X[:, :n] = rng.random((nrhs, n))
    R0 = rng.random((nrhs, n))
    for alpha, beta in ((1.0, 0.0), (-0.5, 0.75)):
        Res[:, :n] = R0
        qm = aligned((NR, pitch))
        emu.emu_qn_twoloop(n, pitch, A, S.ctypes.data, Y.ctypes.data, ys.ctypes.data, X.ctypes.data, ld, Res.ctypes.data, ld, nrhs, NR, qm.ctypes.data,
                           alpha, beta, 0.8, 1, stages)
        for j in range(nrhs):
            xv, rv, q = aligned(n), aligned(n), aligned(pitch)
            xv[:] = X[j, :n]
            rv[:] = R0[j]
            emu.emu_qn_twoloop(n, pitch, A, S.ctypes.data, Y.ctypes.data, ys.ctypes.data, xv.ctypes.data, n, rv.ctypes.data, n, 0, 0, q.ctypes.data,
                               alpha, beta, 0.8, 1, 7)
            assert np.array_equal(Res[j, :n], rv), (j, rel(Res[j, :n], rv))
            assert rel(rv, numpy_twoloop(S[:, :n], Y[:, :n], ys, xv, alpha, beta, 0.8, R0[j])) <= 1e-14


def aligned32(shape, offset=0):
    n = int(np.prod(shape))
    raw = np.zeros(n + 16, dtype=np.float32)
    start = (-(raw.ctypes.data // 4) % 4 + offset // 4) % 16
    return raw[start:start + n].reshape(shape)


def setup_f32(emu):
    i32, i64, d, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
    emu.emu_qn_f32.restype = i32
    emu.emu_qn_f32.argtypes = [i32, i64, i64, i32, vp, vp, vp, vp, vp, vp, d, d, d, i32, i32]
    emu.emu_qn_multi.restype = i32
    emu.emu_qn_multi.argtypes = [i32, i32, i64, i64, i32, vp, vp, vp, i64, vp, i64, i32, d, d, d, i32, i32]


@pytest.mark.parametrize("n,off", [(2500, 0), (1027, 4), (4096, 12)])
def test_float32_instantiations(emu, n, off):
    """qn_compact_kernel<R, LBFGS_FWD, float> and qn_twoloop_kernel<R, float>: 16-byte vectors of FOUR rows, ragged tails, 4-byte
    aligned views; Float32 statements, inner products in double rounded to Float32 (oracle/oracle_f32.py convention)"""
    setup_f32(emu)
    F = np.float32
    rng = np.random.default_rng(n)
    pitch = (n + R - 1) // R * R
    npairs, gamma = 3, F(0.7)
    cols = aligned32((2 * npairs, pitch))
    cols[:, :n] = (0.1 * rng.random((2 * npairs, n))).astype(F)
    x, res = aligned32(n, off), aligned32(n, off)
    x[:] = rng.random(n).astype(F)
    r0 = rng.random(n).astype(F)
    dot = lambda a, b: F(np.dot(a.astype(np.float64), b.astype(np.float64)))
    for alpha, beta in ((1.0, 0.0), (-0.75, 0.5)):
        res[:] = r0
        emu.emu_qn_f32(0, n, pitch, 2 * npairs, cols.ctypes.data, None, None, x.ctypes.data, res.ctypes.data, None, alpha, beta, float(gamma), 1, 5)
        q = x / gamma
        for k in range(npairs):
            a, b = cols[2 * k, :n], cols[2 * k + 1, :n]
            q = q + (dot(b, x) * b - dot(a, x) * a)
        ref = F(alpha) * q + F(beta) * r0 if beta != 0 else F(alpha) * q
        assert res.dtype == F and rel(res.astype(np.float64), ref.astype(np.float64)) <= 1e-6
    A = 3
    S, Y = aligned32((A, pitch)), aligned32((A, pitch))
    S[:, :n] = rng.random((A, n)).astype(F)
    Y[:, :n] = (S[:, :n] + F(0.1) * rng.random((A, n)).astype(F)).astype(F)
    ys = aligned(A)
    ys[:] = [float(dot(S[i, :n], Y[i, :n])) for i in range(A)]
    q = aligned32(pitch)
    res[:] = 0
    emu.emu_qn_f32(1, n, pitch, A, S.ctypes.data, Y.ctypes.data, ys.ctypes.data, x.ctypes.data, res.ctypes.data, q.ctypes.data, 1.0, 0.0, 0.8, 1, 4)
    qq, al = x.copy(), np.zeros(A, F)
    for i in range(A):
        al[i] = F(dot(S[i, :n], qq) / F(ys[i]))
        qq = qq - al[i] * Y[i, :n]
    qq = qq * F(0.8)
    for i in range(A - 1, -1, -1):
        qq = qq + F(al[i] - F(dot(Y[i, :n], qq) / F(ys[i]))) * S[i, :n]
    assert rel(res.astype(np.float64), qq.astype(np.float64)) <= 1e-6


@pytest.mark.parametrize("op,NR,nrhs,stages", [(0, 2, 2, 3), (0, 4, 3, 4), (0, 8, 8, 5), (0, 8, 5, 14), (1, 4, 4, 3), (1, 8, 7, 6)])
def test_block_apply_kernel(emu, op, NR, nrhs, stages):
    """qn_multi_kernel<NR, OP>: mul!(Res, op, X) for the forward operator and L-SR1, every column against the statement-level numpy
    restatement (the block kernel is contracted and uses one reciprocal for the base term: a few ulp)"""
    setup_f32(emu)
    rng = np.random.default_rng(10 * NR + nrhs)
    n = 4100
    RR = 1024 if NR == 8 else 2048
    ncols = 6 if op == 0 else 5
    pitch = (n + 4095) // 4096 * 4096
    cols = aligned((ncols, pitch))
    cols[:, :n] = 0.1 * rng.random((ncols, n))
    cdiv = aligned(ncols)
    cdiv[:] = rng.random(ncols) + 0.5
    ld = n + 2
    X, Res = aligned((nrhs, ld)), aligned((nrhs, ld))
    X[:, :n] = rng.random((nrhs, n))
    R0 = rng.random((nrhs, n))
    gamma = 0.7
    for alpha, beta in ((1.0, 0.0), (-0.75, 0.5)):
        Res[:, :n] = R0
        assert emu.emu_qn_multi(op, NR, n, pitch, ncols, cols.ctypes.data, cdiv.ctypes.data, X.ctypes.data, ld, Res.ctypes.data, ld, nrhs,
                                alpha, beta, gamma, 1, stages) == 0
        for j in range(nrhs):
            x = X[j, :n]
            if op == 0:
                q = x / gamma
                for k in range(ncols // 2):
                    a, b = cols[2 * k, :n], cols[2 * k + 1, :n]
                    q = q + ((b @ x) * b - (a @ x) * a)
                ref = alpha * q + beta * R0[j] if beta != 0 else alpha * q
            else:
                ref = (alpha * x) / gamma + (beta * R0[j] if beta != 0 else 0.0)
                for k in range(ncols):
                    ref = ref + ((alpha * (cols[k, :n] @ x)) / cdiv[k]) * cols[k, :n]
            assert rel(Res[j, :n], ref) <= 1e-13, (j, rel(Res[j, :n], ref))


@pytest.mark.parametrize("stages", [2, 6])
def test_push_rebuild_steps(emu, stages):
    """OP_PUSH_A (src/lbfgs.jl:239-248: a_k = s_k/γ, then `.+=` / `.-=` per older pair, dot(s_k, a_k) on the way out) and OP_PUSH_L
    (src/lsr1.jl:169-179: a_k = y_k - s_k/γ, `.-=` per older column, a_k·s_k and |a_k|² on the way out)"""
    i32, i64, d, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
    emu.emu_qn_push_step.restype = i32
    emu.emu_qn_push_step.argtypes = [i32, i64, i64, i32, vp, vp, vp, vp, vp, d, i32, vp]
    rng = np.random.default_rng(stages)
    n, gamma = 2700, 0.6
    pitch = (n + R - 1) // R * R
    for nprev in (0, 1, 3):
        cols = aligned((max(2 * nprev, 1), pitch))
        cols[:2 * nprev, :n] = 0.1 * rng.random((2 * nprev, n))
        sk, yk, res, out2 = aligned(pitch), aligned(pitch), aligned(n), aligned(2)
        sk[:n], yk[:n] = rng.random(n), rng.random(n)
        emu.emu_qn_push_step(0, n, pitch, 2 * nprev, cols.ctypes.data, None, sk.ctypes.data, yk.ctypes.data, res.ctypes.data, gamma, stages, out2.ctypes.data)
        a = sk[:n] / gamma
        for l in range(nprev):
            al, bl = cols[2 * l, :n], cols[2 * l + 1, :n]
            a = a + (bl @ sk[:n]) * bl
            a = a - (al @ sk[:n]) * al
        assert rel(res, a) <= 1e-14 and abs(out2[0] - sk[:n] @ a) <= 1e-13 * abs(sk[:n] @ a)
        cdiv = aligned(max(nprev, 1))
        cdiv[:] = rng.random(max(nprev, 1)) + 0.5
        emu.emu_qn_push_step(1, n, pitch, nprev, cols.ctypes.data, cdiv.ctypes.data, sk.ctypes.data, yk.ctypes.data, res.ctypes.data, gamma, stages, out2.ctypes.data)
        a = yk[:n] - sk[:n] / gamma
        for l in range(nprev):
            a = a - ((cols[l, :n] @ sk[:n]) / cdiv[l]) * cols[l, :n]
        assert rel(res, a) <= 1e-14
        assert abs(out2[0] - a @ sk[:n]) <= 1e-12 * max(1.0, abs(a @ sk[:n])) and abs(out2[1] - a @ a) <= 1e-13 * (a @ a)


@pytest.mark.parametrize("NR,nrhs,ld_extra", [(2, 2, 0), (4, 3, 1), (8, 8, 4), (8, 6, 2)])
def test_block_apply_kernel_float32(emu, NR, nrhs, ld_extra):
    """qn_multi_kernel<NR, LBFGS_FWD, float>: the Float32 operators' matrix right-hand sides (16-byte vectors of four rows; leading
    dimensions that keep / break the 16-byte alignment of the columns)"""
    i32, i64, d, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_double, ctypes.c_void_p
    emu.emu_qn_multi_f32.restype = i32
    emu.emu_qn_multi_f32.argtypes = [i32, i64, i64, i32, vp, vp, i64, vp, i64, i32, d, d, d, i32]
    F = np.float32
    rng = np.random.default_rng(NR + nrhs)
    n, ncols, gamma = 4104, 6, F(0.7)
    pitch = (n + 4095) // 4096 * 4096
    cols = aligned32((ncols, pitch))
    cols[:, :n] = (0.1 * rng.random((ncols, n))).astype(F)
    ld = n + ld_extra
    X, Res = aligned32((nrhs, ld)), aligned32((nrhs, ld))
    X[:, :n] = rng.random((nrhs, n)).astype(F)
    R0 = rng.random((nrhs, n)).astype(F)
    dot = lambda a, b: F(np.dot(a.astype(np.float64), b.astype(np.float64)))
    for alpha, beta in ((1.0, 0.0), (-0.75, 0.5)):
        Res[:, :n] = R0
        emu.emu_qn_multi_f32(NR, n, pitch, ncols, cols.ctypes.data, X.ctypes.data, ld, Res.ctypes.data, ld, nrhs, alpha, beta, float(gamma), 5)
        for j in range(nrhs):
            x = X[j, :n]
            q = x / gamma
            for k in range(ncols // 2):
                a, b = cols[2 * k, :n], cols[2 * k + 1, :n]
                q = q + (dot(b, x) * b - dot(a, x) * a)
            ref = F(alpha) * q + F(beta) * R0[j] if beta != 0 else F(alpha) * q
            assert rel(Res[j, :n].astype(np.float64), ref.astype(np.float64)) <= 2e-6, (j, rel(Res[j, :n].astype(np.float64), ref.astype(np.float64)))


@pytest.mark.parametrize("base_div,stages", [(0, 3), (1, 7)])
def test_compact_representation_kernel(emu, base_div, stages):
    """qn_compact_kernel<R, INV_COMPACT>: H x = γ x + [S γY] W [S'x; γY'x] (compact inverse) / B x = x/γ + [S Y] W' [S'x; Y'x] (compact
    forward form): all 2m inner products, the 2m x 2m middle product repeated by the CTA, the combine"""
    emu.emu_qn_set_compact.restype = None
    emu.emu_qn_set_compact.argtypes = [ctypes.c_void_p, ctypes.c_int]
    rng = np.random.default_rng(stages)
    n, m, gamma = 2600, 3, 0.8
    cols, pitch = columns(2 * m, n, rng, 0.2)
    W = aligned((2 * m, 2 * m))
    W[:] = rng.random((2 * m, 2 * m)) - 0.5
    x, res = aligned(n), aligned(n)
    x[:] = rng.random(n)
    r0 = rng.random(n)
    emu.emu_qn_set_compact(W.ctypes.data, base_div)
    try:
        for alpha, beta in ((1.0, 0.0), (1.5, -0.25)):
            res[:] = r0
            emu.emu_qn_compact(2, n, pitch, 2 * m, cols.ctypes.data, None, x.ctypes.data, res.ctypes.data, alpha, beta, gamma, 1, stages)
            coef = W @ (cols[:, :n] @ x)
            q = (x / gamma if base_div else x * gamma) + coef @ cols[:, :n]
            ref = alpha * q + beta * r0 if beta != 0 else alpha * q
            assert rel(res, ref) <= 1e-13
    finally:
        emu.emu_qn_set_compact(None, 0)
