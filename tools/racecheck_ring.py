"""Does compute-sanitizer's racecheck understand the ring's mbarrier ordering (consumer arrive on `empty` -> producer wait -> TMA
refill of the slot)?  tools/sanitize_small.py gives every CTA one tile per phase, so no slot is ever refilled between two CTA
barriers there.  Here a grid of 2 CTAs walks 8 tiles each: every kernel refills slots under mbarrier ordering alone.
  compute-sanitizer --tool racecheck python tools/racecheck_ring.py [vector|block]"""
import sys

sys.path.insert(0, ".")
import torch

import linearoperators_jl_b200 as lo

what = sys.argv[1] if len(sys.argv) > 1 else "vector"
ctx = lo.default_context(0)
ctx.set_option("grid", 2)
n = 2048 * 16 + 5
x, r = ctx.uniform(n, 7), ctx.empty(n)
if what == "vector":
    B = lo.LBFGSOperator(n, mem=3, ctx=ctx)
    for i in range(3):
        s = ctx.uniform(n, 100 + i)
        lo.push_(B, s, s + 0.1 * ctx.uniform(n, 200 + i))
    lo.mul_(r, B, x)
else:
    H = lo.InverseLBFGSOperator(n, mem=3, ctx=ctx)
    for i in range(3):
        s = ctx.uniform(n, 100 + i)
        lo.push_(H, s, s + 0.1 * ctx.uniform(n, 200 + i))
    k = 8 if what == "block8" else 3
    Xb = torch.empty((k, n), dtype=torch.float64, device="cuda")
    for j in range(k):
        Xb[j] = ctx.uniform(n, 300 + j)
    Rb = torch.zeros((k, n), dtype=torch.float64, device="cuda")
    lo.mul_(Rb.T, H, Xb.T)
torch.cuda.synchronize()
print("RACECHECK_RING_DONE", what)
