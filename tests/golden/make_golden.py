"""Generates tests/golden/golden_v1.json from the CPU oracle (run once, after test_oracle_pinning passed):
    python tests/golden/make_golden.py
Seeded inputs come from the shared counter-based generator (orc_fill_uniform), so the GPU tests regenerate the
same inputs on the device and compare against these frozen outputs."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def lbfgs_state(orc, n, mem, inverse, npush, scaling=True):
    op = orc.LBFGS(n, mem=mem, scaling=scaling, inverse=inverse)
    for i in range(npush):
        s = orc.uniform(n, 100 + i)
        y = s + 0.1 * orc.uniform(n, 200 + i)
        op.push(s, y)
    return op


def compute_cases(orc):
    out = {}
    n = 1000
    x = orc.uniform(n, 7)
    r0 = orc.uniform(n, 8)
    d = orc.uniform(n, 1)
    v = orc.uniform(n, 2)
    res = r0.copy()
    orc.diag_(res, d, v, 2.0, 2.0)
    out["diag_a2_b2"] = res[:16]
    h = orc.uniform(n, 3)
    h /= np.linalg.norm(h)
    res = np.empty(n)
    orc.householder_(res, h, v, 1.0, 0.0)
    out["householder"] = res[:16]
    # cfg3 chain (H*D + 0.1*I) v
    dd = orc.uniform(n, 4, 0.5, 1.5)
    vv = orc.uniform(n, 5)
    chain = orc.opHouseholder(h) * orc.opDiagonal(dd) + 0.1 * orc.opEye(n)
    out["cfg3_chain"] = chain(vv)[:16]
    for inverse in (False, True):
        op = lbfgs_state(orc, n, 5, inverse, 7)
        out["lbfgs_inv%d_apply" % inverse] = op.apply(x)[:16]
        res = r0.copy()
        op.apply(x, 1.5, -0.25, res=res)
        out["lbfgs_inv%d_apply_ab" % inverse] = res[:16]
    op = lbfgs_state(orc, n, 5, False, 7)
    out["lbfgs_diag"] = op.diag()[:16]
    sr = orc.LSR1(n, mem=5)
    for i in range(7):
        s = orc.uniform(n, 300 + i, -1.0, 1.0)
        y = 2.0 * s + 0.3 * orc.uniform(n, 400 + i, -1.0, 1.0)
        sr.push(s, y)
    out["lsr1_apply"] = sr.apply(x)[:16]
    out["lsr1_diag"] = sr.diag()[:16]
    return {k: np.asarray(v, dtype=np.float64).tolist() for k, v in out.items()}


def sparse_pattern(orc, m, n, density, seed):
    """deterministic m x n sparse matrix from the shared counter-based generator: entry (i, j) is stored when
    u[i*n + j] < density, with value 2*w[i*n + j] - 1; returned as Julia-style CSC (1-based colptr, rowval) + nzval (float64)"""
    u = orc.uniform(m * n, seed).reshape(m, n)
    w = orc.uniform(m * n, seed + 1).reshape(m, n)
    mask = u < density
    colptr1 = np.concatenate([[1], 1 + np.cumsum(mask.sum(axis=0))]).astype(np.int64)
    rows, cols = np.nonzero(mask.T)[1], np.nonzero(mask.T)[0]           # column-major order of the stored entries
    rowval1 = (rows + 1).astype(np.int64)
    nzval = (2.0 * w[rows, cols] - 1.0).astype(np.float64)
    return colptr1, rowval1, nzval


def compute_cases_v2(orc):
    """matrix leaves (LinearOperator(M), dense and sparse: src/constructors.jl:15-29), index operators and the α/β quirks
    Q1-Q4 of SURVEY §8c -- frozen after the oracle passed test_oracle_pinning."""
    out = {}
    m, n = 37, 23
    A = np.asfortranarray((2.0 * orc.uniform(m * n, 11) - 1.0).reshape(n, m).T)          # column-major m x n
    v, u, r0, t0 = orc.uniform(n, 12), orc.uniform(m, 13), orc.uniform(m, 14), orc.uniform(n, 15)
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        Ad, vd, ud = np.asfortranarray(A.astype(dt)), v.astype(dt), u.astype(dt)
        res = np.empty(m, dtype=dt)
        orc.gemv_(res, Ad, vd)
        out["dense_%s_N" % tag] = res
        res = np.empty(n, dtype=dt)
        orc.gemv_(res, Ad, ud, trans=1)
        out["dense_%s_T" % tag] = res
        res = r0.astype(dt)
        orc.gemv_(res, Ad, vd, 1.5, -0.25)
        out["dense_%s_N_ab" % tag] = res
    sm, sn = 60, 45
    colptr1, rowval1, nzval = sparse_pattern(orc, sm, sn, 0.2, 21)
    sv, su, sr0 = orc.uniform(sn, 23), orc.uniform(sm, 24), orc.uniform(sm, 25)
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        res = np.empty(sm, dtype=dt)
        orc.spmv_csc_(res, sm, sn, colptr1, rowval1, nzval.astype(dt), sv.astype(dt))
        out["sparse_%s_N" % tag] = res
        res = np.empty(sn, dtype=dt)
        orc.spmv_csc_(res, sm, sn, colptr1, rowval1, nzval.astype(dt), su.astype(dt), trans=1)
        out["sparse_%s_T" % tag] = res
        res = sr0.astype(dt)
        orc.spmv_csc_(res, sm, sn, colptr1, rowval1, nzval.astype(dt), sv.astype(dt), 1.5, -0.25)
        out["sparse_%s_N_ab" % tag] = res
    # index operators (exact): restriction ignores α, β (Q1); extension zero-fills and the last duplicate wins (Q4)
    x10 = orc.uniform(10, 31)
    idx = np.array([1, 2, 4, 7], dtype=np.int64)
    res = np.full(4, 9.0)
    orc.restrict_(res, idx, x10)
    out["restrict_1247"] = res
    dup = np.array([3, 7, 3, 10], dtype=np.int64)
    res = np.full(10, 9.0)
    orc.extend_(res, dup, x10[:4])
    out["extend_dup_last_wins"] = res
    # rectangular eye: tail = β (not β*res) when β != 0 (Q2); rectangular diagonal: tail zeroed whatever β (Q3)
    v6, r9 = orc.uniform(6, 32), orc.uniform(9, 33)
    res = r9.copy()
    orc.eye_(res, v6, 2.0, 0.5, 6)
    out["eye_9x6_a2_b05"] = res
    d6 = orc.uniform(6, 34)
    res = r9.copy()
    orc.diag_(res, d6, v6, 2.0, 0.5, 6)
    out["diag_9x6_a2_b05"] = res
    return {k: np.asarray(val, dtype=np.float64).tolist() for k, val in out.items()}


def compute_cases_v3(orc):
    """Float32 quasi-Newton operators (oracle/oracle_f32.py): push! + apply of LBFGSOperator / InverseLBFGSOperator / LSR1Operator
    with T = Float32 on seeded inputs (the shared generator, rounded to Float32 exactly as b2o_fill_uniform does)."""
    import oracle_f32 as o32
    f = lambda n, seed, lo=0.0, hi=1.0: orc.uniform(n, seed, lo, hi).astype(np.float32)
    out = {}
    n, mem, npush = 257, 3, 5
    x, r0 = f(n, 7), f(n, 8)
    for tag, op in (("lbfgs", o32.LBFGS32(n, mem)), ("inverse", o32.LBFGS32(n, mem, inverse=True)), ("lsr1", o32.LSR1_32(n, mem))):
        for i in range(npush):
            s = f(n, 100 + i)
            y = (s + np.float32(0.1) * f(n, 200 + i)) if tag != "lsr1" else f(n, 200 + i, -0.5, 1.0)
            op.push(s, y)
        out["f32_%s_apply" % tag] = op.apply(x)
        out["f32_%s_apply_ab" % tag] = op.apply(x, -0.75, 0.5, res=r0)
        out["f32_%s_scalars" % tag] = np.array([op.insert, op.gamma, op.opnorm_upper_bound], dtype=np.float64)
        if tag != "inverse":
            out["f32_%s_a_last" % tag] = op.a[(op.insert - 2) % mem]
            out["f32_%s_diag" % tag] = op.diag()
    # Powell-damped operators: pairs with under- / over-estimated curvature take both damping branches (src/lbfgs.jl:308-314)
    B, H = o32.LBFGS32(n, mem, damped=True), o32.LBFGS32(n, mem, inverse=True, damped=True)
    for i in range(npush):
        s, y, g = f(n, 100 + i), f(n, 200 + i, 0.0, 3.0 if i % 2 else 0.05), f(n, 300 + i)
        B.push(s, y)
        H.push_damped(s, y, 0.7, g)
    out["f32_damped_lbfgs_apply"] = B.apply(x)
    out["f32_damped_inverse_apply"] = H.apply(x)
    return {k: np.asarray(val, dtype=np.float64).tolist() for k, val in out.items()}


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
    import oracle
    oracle.build()
    oracle.set_mode(True, 1)
    cases = compute_cases(oracle)
    json.dump({"generator": "tests/golden/make_golden.py", "oracle": "oracle/b2o_oracle.c (long-double reductions)",
               "cases": cases}, open(os.path.join(HERE, "golden_v1.json"), "w"), indent=1)
    print("wrote", len(cases), "cases")
    cases2 = compute_cases_v2(oracle)
    json.dump({"generator": "tests/golden/make_golden.py (compute_cases_v2)", "oracle": "oracle/b2o_oracle.c (long-double reductions)",
               "cases": cases2}, open(os.path.join(HERE, "golden_v2.json"), "w"), indent=1)
    print("wrote", len(cases2), "cases (v2)")
    cases3 = compute_cases_v3(oracle)
    json.dump({"generator": "tests/golden/make_golden.py (compute_cases_v3)", "oracle": "oracle/oracle_f32.py (numpy Float32 statements, Float64 reductions)",
               "cases": cases3}, open(os.path.join(HERE, "golden_v3.json"), "w"), indent=1)
    print("wrote", len(cases3), "cases (v3)")
