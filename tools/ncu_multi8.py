"""one launch of the SIMT block apply with 8 right-hand sides inside a profiler window (ncu --import-source on ...)"""
import sys

sys.path.insert(0, ".")
import torch

import linearoperators_jl_b200 as lo

ctx = lo.default_context(0)
n = 10**8
B = lo.LBFGSOperator(n, mem=10, ctx=ctx)
for i in range(10):
    s = ctx.uniform(n, 100 + i)
    lo.push_(B, s, s + 0.1 * ctx.uniform(n, 200 + i))
del s
X8 = torch.empty((8, n), dtype=torch.float64, device="cuda")
for j in range(8):
    X8[j] = ctx.uniform(n, 300 + j)
R8 = torch.empty((8, n), dtype=torch.float64, device="cuda")
lo.mul_(R8.T, B, X8.T)
torch.cuda.synchronize()
torch.cuda.profiler.start()
lo.mul_(R8.T, B, X8.T)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("NCU_MULTI8_DONE")
