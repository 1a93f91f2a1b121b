"""%globaltimer timeline of CTA 0 of the kron kernel (ctx option "kron_debug"), incl. the arrival of every ring stage of the
first tile of each phase and the SM clock seen by the kernel (clock64 / globaltimer)."""
import ctypes
import json
import struct
import sys

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import torch

import linearoperators_jl_b200 as lo
from linearoperators_jl_b200 import _lib

ctx = lo.default_context(0)
m = 512
A = torch.randn(m, m, device="cuda").to(torch.bfloat16)
B = torch.randn(m, m, device="cuda").to(torch.bfloat16)
x = torch.randn(m * m, device="cuda").to(torch.bfloat16)
res = torch.empty(m * m, dtype=torch.bfloat16, device="cuda")
K = lo.kron(A, B, ctx=ctx)
cfgs = [(0, 0, 0), (64, 64, 8), (128, 64, 8), (128, 128, 4), (128, 32, 16)]
for bm, bn, cl in cfgs:
    K.set_option("tile_m", bm)
    K.set_option("tile_n", bn)
    K.set_option("cluster", cl)
    for _ in range(20):
        lo.mul_(res, K, x)
    ctx.set_option("kron_debug", 1)
    best = None
    for rep in range(6):
        lo.mul_(res, K, x)
        torch.cuda.synchronize()
        buf = (ctypes.c_double * 40)()
        _lib.check(ctx.lib.b2o_ctx_debug_read(ctx.handle, 448, 40, buf))
        t = struct.unpack("40Q", bytes(buf))
        us = lambda i: round((t[i] - t[0]) / 1000.0, 2)
        cur = {"cfg": [bm, bn, cl], "setup": us(1), "ph0_stage_arrivals": [us(16 + i) for i in range(8) if t[16 + i] >= t[0]],
               "ph0_acc_done": us(3), "Y_published": us(4), "Y_visible": us(5), "ph1_stage_arrivals": [us(24 + i) for i in range(8) if t[24 + i] >= t[0]],
               "ph1_acc_done": us(7), "ph1_store_issued": us(8), "exit": us(10),
               "sm_clock_mhz": round((t[33] - t[32]) / max(1, t[10] - t[0]) * 1000.0, 1)}
        if best is None or cur["exit"] < best["exit"]:
            best = cur
    ctx.set_option("kron_debug", 0)
    print(json.dumps(best), flush=True)
