import sys, ctypes, struct
sys.path.insert(0, "."); sys.path.insert(0, "oracle")
import torch, numpy as np
import linearoperators_jl_b200 as lo
from linearoperators_jl_b200 import _lib
ctx = lo.default_context(0)
m = 512
A = torch.randn(m, m, device="cuda").to(torch.bfloat16); B = torch.randn(m, m, device="cuda").to(torch.bfloat16)
x = torch.randn(m * m, device="cuda").to(torch.bfloat16); res = torch.empty(m * m, dtype=torch.bfloat16, device="cuda")
K = lo.kron(A, B, ctx=ctx)
for _ in range(20): lo.mul_(res, K, x)
ctx.set_option("kron_debug", 1)
for rep in range(3):
    lo.mul_(res, K, x); torch.cuda.synchronize()
    buf = (ctypes.c_double * 16)()
    _lib.check(ctx.lib.b2o_ctx_debug_read(ctx.handle, 448, 16, buf))
    t = struct.unpack("16Q", bytes(buf))
    t0 = t[0]
    names = {1: "setup done", 2: "ph0 first stage landed", 3: "ph0 accumulator complete", 4: "ph0 epilogue done", 5: "grid barrier passed", 6: "ph1 first stage landed", 7: "ph1 accumulator complete", 8: "ph1 epilogue done", 10: "exit", 11: "ph1 epi: first tcgen05.ld done", 13: "ph1 epi: second tcgen05.ld done (first 32 cols stored)", 15: "ph1 epi: all stored"}
    print("rep", rep, {names[i]: (t[i] - t0) / 1000.0 for i in names})
