#!/usr/bin/env python
"""One profiled launch each of the block-apply kernels (NR = 8, NR = 4) and of the streamed a_k rebuild (PUSH_A) at n = 5e7
(ncu --profile-from-start off; the cudaProfilerStart/Stop window below)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import linearoperators_jl_b200 as lo  # noqa: E402


def main():
    ctx = lo.default_context(0)
    n, m = 5 * 10**7, 10
    B = lo.LBFGSOperator(n, mem=m, ctx=ctx)
    for i in range(m):
        s = ctx.uniform(n, 100 + i)
        lo.push_(B, s, s + 0.1 * ctx.uniform(n, 200 + i))
    mats = {}
    for k in (8, 4):
        Xb = torch.empty((k, n), dtype=torch.float64, device="cuda")
        for j in range(k):
            Xb[j] = ctx.uniform(n, 300 + j)
        mats[k] = (Xb.T, torch.empty((k, n), dtype=torch.float64, device="cuda").T)
        lo.mul_(mats[k][1], B, mats[k][0])
    s = ctx.uniform(n, 900)
    y = s + 0.1 * ctx.uniform(n, 901)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for k in (8, 4):
        lo.mul_(mats[k][1], B, mats[k][0])
    lo.push_(B, s, y)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("NCU_MULTI_DONE")


if __name__ == "__main__":
    main()
