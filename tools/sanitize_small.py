#!/usr/bin/env python
"""Small invocation of every kernel family, meant to run under compute-sanitizer (memcheck / synccheck / initcheck)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import linearoperators_jl_b200 as lo  # noqa: E402


def main():
    ctx = lo.default_context(0)
    n = 2 * 4096 + 333
    x, r = ctx.uniform(n, 7), ctx.uniform(n, 8)
    d = ctx.uniform(n, 1)
    h = ctx.uniform(n, 3)
    h /= float(np.sqrt(ctx.dot(h, h)))
    for op in (lo.opDiagonal(d), lo.opEye(n), lo.opZeros(n, n), lo.opOnes(n, n), lo.opHouseholder(h), lo.opEye(n + 5, n),
               lo.opDiagonal(n + 3, n, d)):
        res = ctx.uniform(lo.size(op, 1), 9)
        lo.mul_(res, op, x, 1.5, 0.5)
        lo.mul_(res, op, x)
    P = lo.opRestriction(np.random.default_rng(0).integers(1, n + 1, size=1000), n)
    lo.transpose(P) * (P * x)
    for kind in ("fwd", "inv", "lsr1", "invc"):
        if kind == "lsr1":
            g = lo.LSR1Operator(n, mem=3, ctx=ctx)
        else:
            g = lo.LBFGSOperator(n, mem=3, inverse=kind in ("inv", "invc"), ctx=ctx)
        if kind == "invc":
            g.set_option("inverse_mode", 1)
        for i in range(4):
            s = ctx.uniform(n, 100 + i, -1.0, 1.0) if kind == "lsr1" else ctx.uniform(n, 100 + i)
            y = (2.0 * s + 0.3 * ctx.uniform(n, 200 + i, -1.0, 1.0)) if kind == "lsr1" else s + 0.1 * ctx.uniform(n, 200 + i)
            lo.push_(g, s, y)
        lo.mul_(r, g, x, 1.5, -0.5)
        lo.mul_(r, g, x[:n].clone())
        if kind in ("fwd", "lsr1"):
            lo.diag(g)
        if True:
            # block apply (NR = 2, 4, 8 kernels; two-loop inverse: the block recursion with 4 / 8 columns) and, for the forward
            # form, the streamed a_k rebuild ran in push_ above
            for k in (2, 3, 8):
                Xb = torch.empty((k, n + 2), dtype=torch.float64, device="cuda")
                for j in range(k):
                    Xb[j, :n] = ctx.uniform(n, 300 + j)
                Rb = torch.zeros((k, n + 2), dtype=torch.float64, device="cuda")
                lo.mul_(Rb[:, :n].T, g, Xb[:, :n].T, 1.5, 0.5)
        if kind == "fwd":
            xs = ctx.zeros(n)
            lo.solve_shifted_system_(xs, g, x, 0.5)
            gc = lo.LBFGSOperator(n, mem=3, compact=True, ctx=ctx)
            for i in range(4):
                s = ctx.uniform(n, 100 + i)
                lo.push_(gc, s, s + 0.1 * ctx.uniform(n, 200 + i))
            lo.mul_(r, gc, x, 1.5, -0.5)
            lo.solve_shifted_system_(xs, gc, x, 0.5)
        xh, rh = x.cpu().pin_memory(), torch.empty(n, dtype=torch.float64).pin_memory()
        g.apply_host(rh, xh)
    h2 = ctx.uniform(n, 13)
    h2 /= float(np.sqrt(ctx.dot(h2, h2)))
    for tree in (lo.opHouseholder(h) * lo.opDiagonal(d) + 0.1 * lo.opEye(n),            # in the ahead-of-time table
                 lo.opHouseholder(h) * lo.opDiagonal(d) * lo.opHouseholder(h2) + lo.opDiagonal(d)):   # not in it: NVRTC / interpreter
        for jit, interp in ((1, 0), (2, 0), (0, 0), (0, 2)):     # default; ahead-of-time only; interpreter small / general machine
            ctx.set_option("graph_jit", jit)
            ctx.set_option("graph_interp", interp)
            f = lo.fuse(tree)
            lo.mul_(r, f, x, 2.0, 0.5)
            lo.mul_(r, lo.transpose(f), x)
    ctx.set_option("graph_jit", 1)
    ctx.set_option("graph_interp", 0)
    # ComplexF64 leaves and the conj-sandwich
    cd = torch.complex(ctx.uniform(n, 21, -1.0, 1.0), ctx.uniform(n, 22, -1.0, 1.0))
    cv = torch.complex(ctx.uniform(n, 23, -1.0, 1.0), ctx.uniform(n, 24, -1.0, 1.0))
    ch = torch.complex(ctx.uniform(n, 25, -1.0, 1.0), ctx.uniform(n, 26, -1.0, 1.0))
    ch = ch / torch.linalg.vector_norm(ch)
    cop = lo.opHouseholder(ch) * lo.opDiagonal(cd) + 0.5j * lo.opEye(n)
    for w in (cop, lo.adjoint(cop), lo.transpose(cop), lo.conj(cop)):
        cr = cv.clone()
        lo.mul_(cr, w, cv, 1.5 - 0.5j, 0.25j)
        lo.mul_(cr, w, cv)
    A = torch.randn(64, 72, device="cuda").to(torch.bfloat16)
    B = torch.randn(136, 40, device="cuda").to(torch.bfloat16)
    K = lo.kron(A, B, max_batch=2, ctx=ctx)
    xk = torch.randn(72 * 40, device="cuda").to(torch.bfloat16)
    K * xk
    lo.transpose(K) * (K * xk)
    K.apply_batch(torch.stack([xk, xk]).contiguous())
    for bm, bn, cl in ((64, 32, 4), (128, 64, 2), (128, 128, 1), (64, 64, 8)):       # every tile shape of the clustered kernel
        K.set_option("tile_m", bm)
        K.set_option("tile_n", bn)
        K.set_option("cluster", cl)
        r32 = torch.empty(64 * 136, dtype=torch.float32, device="cuda")
        lo.mul_(r32, K, xk)
        lo.mul_(r32, K, xk, 2.0, -0.5)                                               # beta != 0: direct-store epilogue
        K.apply_batch(torch.stack([xk, xk]).contiguous())
    # the cta_group::2 pair kernel (ragged rows / columns / K, both store paths, beta != 0)
    A2 = torch.randn(264, 200, device="cuda").to(torch.bfloat16)
    B2 = torch.randn(392, 328, device="cuda").to(torch.bfloat16)
    K2 = lo.kron(A2, B2, max_batch=2, ctx=ctx)
    K2.set_option("tile_m", 256)
    X2 = torch.randn(2, 200 * 328, device="cuda").to(torch.bfloat16)
    for tma in (1, 0):
        K2.set_option("pair_tma_stores", tma)
        R2 = K2.apply_batch(X2)
        K2.apply_batch(X2, alpha=2.0, beta=-0.5, res=R2)
        K2.apply_batch(R2, trans=True, res=torch.empty((2, 200 * 328), dtype=torch.float32, device="cuda"))
    # Float32 quasi-Newton operators (Float32 instantiations of the streaming kernels, 4096-row tiles)
    n32 = 3 * 4096 + 77
    f32 = lambda seed, lo_=0.0, hi_=1.0: ctx.fill_uniform(ctx.empty(n32, dtype=torch.float32), seed, lo_, hi_)
    for kind in ("fwd", "inv", "lsr1"):
        g = lo.LSR1Operator(n32, mem=3, T=torch.float32, ctx=ctx) if kind == "lsr1" else lo.LBFGSOperator(n32, mem=3, inverse=kind == "inv", T=torch.float32, ctx=ctx)
        for i in range(4):
            s = f32(100 + i)
            lo.push_(g, s, s + 0.1 * f32(200 + i) if kind != "lsr1" else f32(200 + i, -0.5, 1.0))
        r32v = f32(8)
        lo.mul_(r32v, g, f32(7), 1.5, -0.5)
        big = torch.zeros(n32 + 3, dtype=torch.float32, device="cuda")
        lo.mul_(big[3:], g, f32(7))                                                      # 4-byte aligned views
        if kind != "inv":
            lo.diag(g)
    for inv in (False, True):                                                            # Powell-damped Float32 push!
        g = lo.LBFGSOperator(torch.float32, n32, mem=3, damped=True, inverse=inv, ctx=ctx)
        for i in range(4):
            s, y = f32(100 + i), f32(200 + i, 0.0, 3.0 if i % 2 else 0.05)
            if inv:
                lo.push_(g, s, y, 0.7, f32(300 + i))
            else:
                lo.push_(g, s, y)
        lo.mul_(r32v, g, f32(7))
    torch.cuda.synchronize()
    print("SANITIZE_OK launches=%d" % ctx.launch_count())


if __name__ == "__main__":
    main()
