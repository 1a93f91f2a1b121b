#!/usr/bin/env python
"""debug: determinism + per-column agreement of the block apply at scale"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import linearoperators_jl_b200 as lo  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10**8
m = 10
ctx = lo.default_context(0)
B = lo.LBFGSOperator(n, mem=m, ctx=ctx)
for i in range(m):
    s = ctx.uniform(n, 100 + i)
    lo.push_(B, s, s + 0.1 * ctx.uniform(n, 200 + i))
for k in (8, 4):
    Xb = torch.empty((k, n), dtype=torch.float64, device="cuda")
    for j in range(k):
        Xb[j] = ctx.uniform(n, 300 + j)
    X = Xb.T
    ref = torch.empty((k, n), dtype=torch.float64, device="cuda")
    for j in range(k):
        lo.mul_(ref[j], B, Xb[j])
    prev = None
    for it in range(int(os.environ.get('ITS', '24'))):
        Res = torch.full((k, n), float("nan"), dtype=torch.float64, device="cuda").T
        lo.mul_(Res, B, X)
        d = [float(torch.linalg.norm(Res[:, j] - ref[j]) / torch.linalg.norm(ref[j])) for j in range(k)]
        same = None if prev is None else bool(torch.equal(prev, Res))
        bad = [(j, int((Res[:, j] - ref[j]).abs().gt(1e-9 * ref[j].abs().max()).sum())) for j in range(k)]
        if max(d) > 1e-13 or same is False:
            print("k=%d it=%d rel=%s bitwise_same_as_prev=%s n_bad_rows=%s" % (k, it, ["%.1e" % v for v in d], same, bad), flush=True)
        prev = Res.clone()
    print("k=%d done" % k, flush=True)
    del Xb, X, ref, Res, prev
    torch.cuda.empty_cache()
