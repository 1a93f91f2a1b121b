"""%globaltimer timeline of CTA 0 of the cta_group::2 pair kernel (ctx option "kron_debug").
  python tools/kron_pair_dbg.py [nb]
kron_debug = 1: per accumulator tile, when the MMA warp got the accumulator / issued its last commit, and when the epilogue got the
full accumulator / finished draining it.  kron_debug = 2 + k: inside tile k, per 32-column chunk of the epilogue: start, TMEM load
returned, staging buffer free (+ barrier), staging written, store issued."""
import ctypes
import json
import struct
import sys

sys.path.insert(0, ".")
import torch

import linearoperators_jl_b200 as lo
from linearoperators_jl_b200 import _lib

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 148
m = 512
ctx = lo.default_context(0)
A = torch.randn(m, m, device="cuda").to(torch.bfloat16)
B = torch.randn(m, m, device="cuda").to(torch.bfloat16)
X = torch.randn(nb, m * m, device="cuda").to(torch.bfloat16)
R = torch.empty((nb, m * m), dtype=torch.bfloat16, device="cuda")
K = lo.kron(A, B, max_batch=nb, ctx=ctx)
K.set_option("tile_m", 256)
for _ in range(3):
    K.apply_batch(X, res=R)


def read():
    torch.cuda.synchronize()
    buf = (ctypes.c_double * 64)()
    _lib.check(ctx.lib.b2o_ctx_debug_read(ctx.handle, 448, 64, buf))
    return struct.unpack("64Q", bytes(buf))


ctx.set_option("kron_debug", 1)
best = None
for _ in range(5):
    K.apply_batch(X, res=R)
    t = read()
    us = lambda i: round((t[i] - t[0]) / 1000.0, 2) if t[i] else None
    cur = {"exit": us(10), "tiles": [{"tile": k, "mma_got_accumulator": us(16 + 2 * k), "mma_last_commit_issued": us(17 + 2 * k),
                                      "epi_accumulator_full": us(48 + 2 * k), "epi_drained": us(49 + 2 * k)} for k in range(8)]}
    if best is None or cur["exit"] < best["exit"]:
        best = cur
print(json.dumps({"case": "pair kernel timeline, CTA 0 (leader of cluster 0), us from kernel entry; tiles alternate GEMM 1 (Y, 2 per unit) / GEMM 2 in the "
                          "order P0(u0) x2, P0(u1) x2, P1(u0) x2, ...", "nb": nb, **best}))
for tile in (2, 4):
    ctx.set_option("kron_debug", 2 + tile)
    K.apply_batch(X, res=R)
    K.apply_batch(X, res=R)
    t = read()
    base = t[16]
    rows = [[round((t[16 + 5 * c + k] - base) / 1000.0, 3) for k in range(5)] for c in range(8)]
    print(json.dumps({"case": "epilogue chunks of tile %d (%s), us from the tile's first chunk: [start, tcgen05.ld returned, staging buffer free + barrier, "
                              "staging written, store issued]" % (tile, "GEMM 1: Y" if tile < 4 else "GEMM 2: result"), "chunks": rows}))
ctx.set_option("kron_debug", 0)
