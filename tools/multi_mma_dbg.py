"""one small block apply on the DMMA kernel against the column loop (debug aid)"""
import sys

sys.path.insert(0, ".")
import torch

import linearoperators_jl_b200 as lo

n, mem, k = int(sys.argv[1]) if len(sys.argv) > 1 else 4099, 5, 8
ctx = lo.default_context(0)
B = lo.LBFGSOperator(n, mem=mem, ctx=ctx)
for i in range(7):
    s = ctx.uniform(n, 100 + i)
    lo.push_(B, s, s + 0.1 * ctx.uniform(n, 200 + i))
Xb = torch.empty((k, n), dtype=torch.float64, device="cuda")
for j in range(k):
    Xb[j] = ctx.uniform(n, 300 + j)
R = torch.zeros((k, n), dtype=torch.float64, device="cuda")
print("launching", flush=True)
lo.mul_(R.T, B, Xb.T)
torch.cuda.synchronize()
r = ctx.empty(n)
err = 0.0
for j in range(k):
    lo.mul_(r, B, Xb[j])
    err = max(err, float(torch.linalg.norm(R[j] - r) / torch.linalg.norm(r)))
print("MULTI_MMA_OK max rel diff vs vector apply", err, flush=True)
