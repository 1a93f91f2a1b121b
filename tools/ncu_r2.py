#!/usr/bin/env python
"""Round-2 ncu targets, one profiled launch each inside a cudaProfilerStart/Stop window (run under
`ncu --profile-from-start off ...`): forward L-BFGS apply (cfg2), inverse two-loop and compact inverse (cfg5 slab), the clustered
kron kernel (cfg4), random gather / gather-form extension at the three L2 fetch granularities."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import linearoperators_jl_b200 as lo  # noqa: E402


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "qn"
    ctx = lo.default_context(0)
    n = 10**8
    targets, pre = [], []
    if what == "qn":
        x, res = ctx.uniform(n, 7), ctx.empty(n)
        B = lo.LBFGSOperator(n, mem=10, ctx=ctx)
        for i in range(10):
            s = ctx.uniform(n, 100 + i)
            lo.push_(B, s, s + 0.1 * ctx.uniform(n, 200 + i))
        H = lo.InverseLBFGSOperator(n, mem=20, ctx=ctx)
        for i in range(20):
            s = ctx.uniform(n, 100 + i)
            lo.push_(H, s, s + 0.1 * ctx.uniform(n, 200 + i))
        del s
        Hc = lo.InverseLBFGSOperator(n, mem=20, compact=True, ctx=ctx) if False else None
        targets = [lambda: lo.mul_(res, B, x), lambda: lo.mul_(res, H, x)]
        pre = targets
    elif what == "kron":
        m = 512
        A = torch.randn(m, m, device="cuda").to(torch.bfloat16)
        Bm = torch.randn(m, m, device="cuda").to(torch.bfloat16)
        xx = torch.randn(m * m, device="cuda").to(torch.bfloat16)
        X64 = torch.randn(64, m * m, device="cuda").to(torch.bfloat16)
        K = lo.kron(A, Bm, max_batch=64, ctx=ctx)
        r = torch.empty(m * m, dtype=torch.bfloat16, device="cuda")
        R64 = torch.empty((64, m * m), dtype=torch.bfloat16, device="cuda")
        targets = [lambda: lo.mul_(r, K, xx), lambda: K.apply_batch(X64, res=R64)]
        pre = targets * 5
    elif what == "new":
        # second half of round 2: Float32 forward apply, block two-loop (4 RHS), block apply 8 RHS on the SIMT and on the DMMA kernel
        f32 = lambda seed: ctx.fill_uniform(ctx.empty(n, dtype=torch.float32), seed)
        B32 = lo.LBFGSOperator(n, mem=10, T=torch.float32, ctx=ctx)
        for i in range(10):
            s = f32(100 + i)
            lo.push_(B32, s, s + 0.1 * f32(200 + i))
        x32, r32 = f32(7), ctx.empty(n, dtype=torch.float32)
        B = lo.LBFGSOperator(n, mem=10, ctx=ctx)
        for i in range(10):
            s = ctx.uniform(n, 100 + i)
            lo.push_(B, s, s + 0.1 * ctx.uniform(n, 200 + i))
        H = lo.InverseLBFGSOperator(n, mem=10, ctx=ctx)
        for i in range(10):
            s = ctx.uniform(n, 100 + i)
            lo.push_(H, s, s + 0.1 * ctx.uniform(n, 200 + i))
        del s
        X8 = torch.empty((8, n), dtype=torch.float64, device="cuda")
        for j in range(8):
            X8[j] = ctx.uniform(n, 300 + j)
        R8 = torch.empty((8, n), dtype=torch.float64, device="cuda")

        def mma(flag, f):
            def run():
                ctx.set_option("multi_mma", flag)
                f()
                ctx.set_option("multi_mma", 0)
            return run
        targets = [lambda: lo.mul_(r32, B32, x32), lambda: lo.mul_(R8[:4].T, H, X8[:4].T), mma(0, lambda: lo.mul_(R8.T, B, X8.T)),
                   mma(1, lambda: lo.mul_(R8.T, B, X8.T))]
        pre = targets
    elif what == "index":
        k = n // 4
        v, uk = ctx.uniform(n, 7), ctx.uniform(k, 8)
        rk, res = ctx.empty(k), ctx.empty(n)
        P = lo.opRestriction(np.random.default_rng(0).integers(1, n + 1, size=k), n)
        Z = lo.transpose(P)

        def with_gran(g, f):
            def run():
                ctx.set_option("l2_fetch_granularity", g)
                f()
            return run
        for g in (128, 64, 32):
            targets += [with_gran(g, lambda: lo.mul_(rk, P, v)), with_gran(g, lambda: lo.mul_(res, Z, uk))]
        ctx.set_option("extend_form", 1)
        pre = [lambda: lo.mul_(rk, P, v), lambda: lo.mul_(res, Z, uk)]
        ctx.set_option("extend_form", 0)
    for f in pre:
        f()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for f in targets:
        f()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("NCU_R2_DONE", what)


if __name__ == "__main__":
    main()
