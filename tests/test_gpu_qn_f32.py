"""GPU parity for the Float32 quasi-Newton operators (LBFGSOperator(Float32, n), InverseLBFGSOperator(Float32, n),
LSR1Operator(Float32, n); test/test_lbfgs.jl:162-178, test/test_lsr1.jl:74-86): the Float32 instantiations of the persistent
TMA-ring kernels against the numpy Float32 restatement (oracle/oracle_f32.py).  Elementwise statements are the same Float32
operations on both sides; inner products are accumulated in double on both sides in different orders and rounded to Float32,
so parity is to Float32 rounding: norm-wise relative <= 1e-5 (L-BFGS), <= 1e-4 (L-SR1 recurrences)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (d if d > 0 else 1.0)


def host(t):
    return t.detach().cpu().numpy()


def f32(ctx, n, seed, lo=0.0, hi=1.0):
    import torch
    return ctx.fill_uniform(ctx.empty(n, dtype=torch.float32), seed, lo, hi)


def test_reference_precision_testset_float32(lo, ctx):
    """test/test_lbfgs.jl:162-178 and test/test_lsr1.jl:74-86 for T = Float32, statement by statement"""
    import torch
    n, mem = 10, 5
    dev = "cuda:%d" % ctx.device
    B = lo.LBFGSOperator(n, mem=mem, T=torch.float32, ctx=ctx)
    H = lo.InverseLBFGSOperator(n, mem=mem, T=torch.float32, ctx=ctx)
    L = lo.LSR1Operator(n, mem=mem, T=torch.float32, ctx=ctx)
    v = torch.tensor([-(-1.0) ** i for i in range(1, n + 1)], dtype=torch.float32, device=dev)
    for op in (B, H, L):
        assert np.array_equal(host(op * v), host(v))                       # identity before the first push
        s, y = torch.ones(n, dtype=torch.float32, device=dev), torch.ones(n, dtype=torch.float32, device=dev)
        lo.push_(op, s, y)
        assert lo.eltype(op) == torch.float32
        r = op * v
        assert r.dtype == torch.float32 and r.shape == (n,)
    with pytest.raises(lo.B2OError):
        lo.push_(B, torch.ones(n, dtype=torch.float64, device=dev), torch.ones(n, dtype=torch.float64, device=dev))   # wrong element type
    with pytest.raises(lo.B2OError):
        lo.solve_shifted_system_(torch.empty(n, dtype=torch.float32, device=dev), B, v, 0.5)      # Float64 only


@pytest.mark.parametrize("kind", ["lbfgs", "inverse", "lsr1"])
@pytest.mark.parametrize("n,mem,npush", [(10, 3, 2), (1000, 5, 7), (4097, 5, 5), (100003, 10, 13), (75776, 3, 3)])
def test_f32_push_and_apply_vs_numpy_oracle(lo, ctx, kind, n, mem, npush):
    import torch
    import oracle_f32 as o32
    if kind == "lbfgs":
        op, o = lo.LBFGSOperator(n, mem=mem, T=torch.float32, ctx=ctx), o32.LBFGS32(n, mem)
    elif kind == "inverse":
        op, o = lo.InverseLBFGSOperator(n, mem=mem, T=torch.float32, ctx=ctx), o32.LBFGS32(n, mem, inverse=True)
    else:
        op, o = lo.LSR1Operator(n, mem=mem, T=torch.float32, ctx=ctx), o32.LSR1_32(n, mem)
    tol = 1e-5 if kind != "lsr1" else 1e-4
    for i in range(npush):
        s = f32(ctx, n, 100 + i)
        y = (s + 0.1 * f32(ctx, n, 200 + i)) if kind != "lsr1" else f32(ctx, n, 200 + i, -0.5, 1.0)
        lo.push_(op, s, y)
        acc = o.push(host(s), host(y))
        assert bool(op.last_push_accepted) == bool(acc)
    ins, gamma, ub, ys, aux = op.data._scalars()
    assert ins == o.insert
    assert abs(gamma - float(o.gamma)) <= 1e-6 * abs(float(o.gamma))
    assert np.allclose(ys, o.ys, rtol=1e-6, atol=0)
    assert abs(ub - float(o.opnorm_upper_bound)) <= 1e-4 * abs(float(o.opnorm_upper_bound))
    for k in range(mem):
        assert np.array_equal(host(op.data.col("s", k)), o.s[k]) and np.array_equal(host(op.data.col("y", k)), o.y[k])
        if kind != "inverse":
            assert rel(host(op.data.col("a", k)), o.a[k]) <= tol, (k, rel(host(op.data.col("a", k)), o.a[k]))
        if kind == "lbfgs":
            assert rel(host(op.data.col("b", k)), o.b[k]) <= 1e-6
    x, r0 = f32(ctx, n, 7), f32(ctx, n, 8)
    l0 = ctx.launch_count()
    res = op * x
    assert ctx.launch_count() - l0 == 1                                     # ONE persistent launch per apply
    assert res.dtype == torch.float32
    assert rel(host(res), o.apply(host(x))) <= tol, rel(host(res), o.apply(host(x)))
    out = r0.clone()
    lo.mul_(out, op, x, -0.75, 0.5)                                          # 5-arg form
    assert rel(host(out), o.apply(host(x), -0.75, 0.5, res=host(r0))) <= tol
    # unaligned views (4-byte aligned only): same values
    big = torch.empty(n + 3, dtype=torch.float32, device=x.device)
    big[1:n + 1] = x
    out2 = torch.empty(n + 3, dtype=torch.float32, device=x.device)
    lo.mul_(out2[3:], op, big[1:n + 1])
    assert np.array_equal(host(out2[3:]), host(res))
    assert np.array_equal(host(op * x), host(res))                           # run-to-run bit determinism
    # matrix right-hand sides (src/operations.jl:34-36: column j of Res = the vector apply of column j of X) and host buffers
    for k in (3, 8, 11):                                                     # NR = 4, 8, 8 + 4 instantiations of the Float32 block kernel
        Xb = torch.stack([f32(ctx, n, 400 + j) for j in range(k)]).contiguous()
        Rb = torch.empty_like(Xb)
        l0 = ctx.launch_count()
        lo.mul_(Rb.T, op, Xb.T)
        if kind != "inverse":
            assert ctx.launch_count() - l0 == (1 if k <= 8 else 2)           # forward / L-SR1: one launch per 8 right-hand sides
        for j in (0, k - 1):
            assert rel(host(Rb[j]), host(op * Xb[j])) <= 2e-6                # contracted block kernel vs the vector kernel
            assert rel(host(Rb[j]), o.apply(host(Xb[j]))) <= tol
    xh, rh = x.cpu().pin_memory(), torch.empty(n, dtype=torch.float32).pin_memory()
    op.apply_host(rh, xh)
    assert np.array_equal(rh.numpy(), host(res))
    # a pair with non-positive curvature is rejected and leaves the state alone (src/lbfgs.jl:281)
    if kind != "lsr1":
        lo.push_(op, x, -x)
        assert not op.last_push_accepted
        assert np.array_equal(host(op * x), host(res))
    lo.reset_(op)                                                            # reset!: identity again
    assert np.array_equal(host(op * x), host(x))


def test_f32_secant_equation_and_inverse_pair_on_gpu(lo, ctx):
    """the reference's predicates (test/test_lbfgs.jl:45-52) on the Float32 CUDA path: B s = y, H y = s, H (B x) = x"""
    import torch
    n, mem = 50000, 6
    B = lo.LBFGSOperator(n, mem=mem, T=torch.float32, ctx=ctx)
    H = lo.InverseLBFGSOperator(n, mem=mem, T=torch.float32, ctx=ctx)
    for i in range(9):
        s = f32(ctx, n, 100 + i)
        y = s + 0.1 * f32(ctx, n, 200 + i)
        lo.push_(B, s, y)
        lo.push_(H, s, y)
        assert rel(host(B * s), host(y)) <= 2e-5
        assert rel(host(H * y), host(s)) <= 2e-5
    x = f32(ctx, n, 9)
    assert rel(host(H * (B * x)), host(x)) <= 1e-4
    assert B.apply_bytes() == (4 * mem + 3) * 4.0 * n                         # half the bytes of the Float64 operator


def test_f32_golden_vectors_on_gpu(lo, ctx, orc):
    """tests/golden/golden_v3.json: frozen outputs of the numpy Float32 restatement (push! x 5 + apply, n = 257, mem = 3) against the
    CUDA path on the same seeded inputs (b2o_fill_uniform rounds the shared generator to Float32 exactly as the fixture script does)"""
    import json
    import os
    import torch
    G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "golden_v3.json")))["cases"]
    n, mem, npush = 257, 3, 5
    x, r0 = f32(ctx, n, 7), f32(ctx, n, 8)
    assert np.array_equal(host(x), orc.uniform(n, 7).astype(np.float32))
    for tag, op in (("lbfgs", lo.LBFGSOperator(torch.float32, n, mem=mem, ctx=ctx)),
                    ("inverse", lo.InverseLBFGSOperator(torch.float32, n, mem=mem, ctx=ctx)),
                    ("lsr1", lo.LSR1Operator(torch.float32, n, mem=mem, ctx=ctx))):
        for i in range(npush):
            s = f32(ctx, n, 100 + i)
            y = (s + 0.1 * f32(ctx, n, 200 + i)) if tag != "lsr1" else f32(ctx, n, 200 + i, -0.5, 1.0)
            lo.push_(op, s, y)
        tol = 1e-5 if tag != "lsr1" else 1e-4
        assert rel(host(op * x), G["f32_%s_apply" % tag]) <= tol, (tag, rel(host(op * x), G["f32_%s_apply" % tag]))
        out = r0.clone()
        lo.mul_(out, op, x, -0.75, 0.5)
        assert rel(host(out), G["f32_%s_apply_ab" % tag]) <= tol
        ins, gamma, ub, _, _ = op.data._scalars()
        gi, gg, gu = G["f32_%s_scalars" % tag]
        assert ins == int(gi) and abs(gamma - gg) <= 1e-6 * abs(gg) and abs(ub - gu) <= 1e-4 * abs(gu)
        if tag != "inverse":
            assert rel(host(op.data.col("a", (ins - 2) % mem)), G["f32_%s_a_last" % tag]) <= tol
            assert rel(host(lo.diag(op)), G["f32_%s_diag" % tag]) <= tol
    B = lo.LBFGSOperator(torch.float32, n, mem=mem, damped=True, ctx=ctx)
    H = lo.InverseLBFGSOperator(torch.float32, n, mem=mem, damped=True, ctx=ctx)
    for i in range(npush):
        s, y, g = f32(ctx, n, 100 + i), f32(ctx, n, 200 + i, 0.0, 3.0 if i % 2 else 0.05), f32(ctx, n, 300 + i)
        lo.push_(B, s, y)
        lo.push_(H, s, y.clone(), 0.7, g)
    assert rel(host(B * x), G["f32_damped_lbfgs_apply"]) <= 2e-5
    assert rel(host(H * x), G["f32_damped_inverse_apply"]) <= 2e-5


@pytest.mark.parametrize("n,mem,npush", [(1000, 4, 7), (100003, 5, 9)])
def test_f32_damped_push_and_diag_vs_numpy_oracle(lo, ctx, n, mem, npush):
    """Powell-damped push! in all calling sequences (src/lbfgs.jl:269-367) and diag! (:379-395, src/lsr1.jl:196-211) for T = Float32;
    pairs alternate between strongly under- and over-estimated curvature so that both damping branches are taken"""
    import torch
    import oracle_f32 as o32
    B, oB = lo.LBFGSOperator(torch.float32, n, mem=mem, damped=True, ctx=ctx), o32.LBFGS32(n, mem, damped=True)
    H, oH = lo.InverseLBFGSOperator(torch.float32, n, mem=mem, damped=True, ctx=ctx), o32.LBFGS32(n, mem, inverse=True, damped=True)
    L, oL = lo.LSR1Operator(torch.float32, n, mem=mem, ctx=ctx), o32.LSR1_32(n, mem)
    for i in range(npush):
        s = f32(ctx, n, 100 + i)
        y = f32(ctx, n, 200 + i, 0.0, 3.0 if i % 2 else 0.05)
        g = f32(ctx, n, 300 + i)
        if i % 3 == 0:
            lo.push_(B, s, y)                                                # push!(op, s, y) on a damped operator  :274-276
        else:
            lo.push_(B, s, y, torch.empty_like(s))                           # push!(op, s, y, Bs)
        oB.push(host(s), host(y))
        yd = y.clone()
        lo.push_(H, s, yd, 0.7, g)                                           # push!(op, s, y, α, g): y is overwritten with the damped y
        yo = oH.push_damped(host(s), host(y), 0.7, host(g))
        assert rel(host(yd), yo) <= 1e-6
        yl = f32(ctx, n, 200 + i, -0.5, 1.0)
        lo.push_(L, s, yl)
        oL.push(host(s), host(yl))
    with pytest.raises(lo.ErrorException):
        lo.push_(H, s, y)                                                    # damped inverse operators need α and g  :296-298
    x = f32(ctx, n, 7)
    assert rel(host(B * x), oB.apply(host(x))) <= 2e-5
    assert rel(host(H * x), oH.apply(host(x))) <= 2e-5
    assert rel(host(lo.diag(B)), oB.diag()) <= 1e-5
    assert rel(host(lo.diag(L)), oL.diag()) <= 1e-4
    assert lo.diag(B).dtype == torch.float32
    with pytest.raises(lo.LinearOperatorException):
        lo.diag(H)                                                           # only forward approximations  :380-382
