"""kron(A,B)*vec on the clustered tcgen05 kernel: correctness against the oracle + event-timed sweep over (cluster, tile_n).

  python tools/kron_bench.py [--m 512] [--quick]
Prints one JSON line per configuration.  Timing: 200 back-to-back launches between ONE pair of CUDA events on the context
stream (the launches are identical, so event time / 200 = steady-state kernel period incl. launch gap), and the per-launch
event pair of ctx option "time_kernels" (one kernel, cold-ish: includes the event overhead)."""
import ctypes
import json
import struct
import sys
import time

sys.path.insert(0, ".")
sys.path.insert(0, "oracle")
import numpy as np
import torch

import linearoperators_jl_b200 as lo
from linearoperators_jl_b200 import _lib
import oracle as orc

m = 512
if "--m" in sys.argv:
    m = int(sys.argv[sys.argv.index("--m") + 1])
quick = "--quick" in sys.argv
PEAK = 1644.5
try:
    PEAK = float(json.load(open("MEASURED_PEAKS.json"))["bf16_tflops"])
except Exception:
    pass
ctx = lo.default_context(0)
orc.set_mode(True, 1)
mk_np = lambda seed, shape: orc.bf16_round(orc.uniform(int(np.prod(shape)), seed, -1.0, 1.0)).reshape(shape)
to_dev = lambda a: torch.as_tensor(a).cuda().to(torch.bfloat16).contiguous()
An, Bn, xn = mk_np(11, (m, m)), mk_np(12, (m, m)), mk_np(13, (m * m,))
A, B, x = to_dev(An), to_dev(Bn), to_dev(xn)
ref = np.empty(m * m)
orc.kron_(ref, An, Bn, xn)
K = lo.kron(A, B, max_batch=296, ctx=ctx)
res = torch.empty(m * m, dtype=torch.bfloat16, device="cuda")
r32 = torch.empty(m * m, dtype=torch.float32, device="cuda")
fl = K.flops()


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def timed(fn, reps, warmup=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def per_launch(fn, reps=100):
    ctx.set_option("time_kernels", 1)
    ctx.kernel_time(reset=True)
    for _ in range(reps):
        fn()
    kms, kn = ctx.kernel_time(reset=True)
    ctx.set_option("time_kernels", 0)
    return kms / max(kn, 1)


def timeline():
    ctx.set_option("kron_debug", 1)
    best = None
    for _ in range(5):
        lo.mul_(res, K, x)
        torch.cuda.synchronize()
        buf = (ctypes.c_double * 16)()
        _lib.check(ctx.lib.b2o_ctx_debug_read(ctx.handle, 448, 16, buf))
        t = struct.unpack("16Q", bytes(buf))
        names = {1: "setup+cluster sync", 2: "ph0 first stage landed", 3: "ph0 accumulator complete", 4: "Y stored+published",
                 5: "Y of whole cluster visible", 6: "ph1 first stage landed", 7: "ph1 accumulator complete", 8: "ph1 tile staged+TMA store issued",
                 10: "exit"}
        cur = {names[i]: round((t[i] - t[0]) / 1000.0, 2) for i in names}
        if best is None or cur["exit"] < best["exit"]:
            best = cur
    ctx.set_option("kron_debug", 0)
    return best


def floor(grid, cluster, smem):
    fn = lambda: _lib.check(ctx.lib.b2o_kron_launch_floor(ctx.handle, grid, cluster, smem))
    return {"grid": grid, "cluster": cluster, "smem": smem, "us_event_pair": round(per_launch(fn) * 1e3, 2), "us_back_to_back": round(timed(fn, 200) * 1e3, 2)}


print(json.dumps({"case": "launch floor: EMPTY kernel, same launch shape", "floors": [floor(128, 16, 150 * 1024), floor(64, 8, 200 * 1024),
                  floor(32, 8, 220 * 1024), floor(128, 1, 150 * 1024), floor(128, 1, 0)]}), flush=True)
configs = [(0, 0, 0)] + ([] if quick else [(64, 32, 16), (64, 32, 8), (64, 64, 8), (64, 64, 4), (128, 32, 16), (128, 64, 8), (64, 32, 4), (64, 32, 1)])
for bm, bn, cluster in configs:
    try:
        K.set_option("cluster", cluster)
        K.set_option("tile_n", bn)
        K.set_option("tile_m", bm)
        lo.mul_(r32, K, x)
        e32 = rel(r32.double().cpu().numpy(), ref)
        lo.mul_(res, K, x)
        e16 = rel(res.double().cpu().numpy(), orc.bf16_round(ref))
        t32 = torch.empty(m * m, dtype=torch.float32, device="cuda")
        lo.mul_(t32, lo.transpose(K), x)
        reft = np.empty(m * m)
        orc.kron_(reft, An, Bn, xn, trans=1)
        et = rel(t32.double().cpu().numpy(), reft)
        ms_b2b = timed(lambda: lo.mul_(res, K, x), 200)
        ms_one = per_launch(lambda: lo.mul_(res, K, x))
        print(json.dumps({"case": "cfg4 kron %dx%d bf16" % (m, m), "tile_m": bm or "auto", "tile_n": bn or "auto", "cluster": cluster or "auto",
                          "rel_err_f32_out": e32, "rel_err_bf16_out_vs_rounded_oracle": e16, "rel_err_transpose_f32": et,
                          "us_kernel_event_pair_per_launch": round(ms_one * 1e3, 2), "us_back_to_back_period": round(ms_b2b * 1e3, 2),
                          "TFLOPs_event_pair": round(fl / (ms_one * 1e-3) / 1e12, 1), "frac_of_measured_bf16_burst": round(fl / (ms_one * 1e-3) / 1e12 / PEAK, 4),
                          "timeline_us": timeline()}), flush=True)
    except Exception as e:
        print(json.dumps({"tile_m": bm, "tile_n": bn, "cluster": cluster, "error": str(e)[:300]}), flush=True)
K.set_option("tile_m", 0)
K.set_option("cluster", 0)
K.set_option("tile_n", 0)

# ---- batched right-hand sides (EXTRA: not a reference feature)
Xn = mk_np(14, (64, m * m))
X = to_dev(Xn)
R = torch.empty((64, m * m), dtype=torch.bfloat16, device="cuda")
refs = {}
for b in (0, 31, 63):
    rb = np.empty(m * m)
    orc.kron_(rb, An, Bn, Xn[b])
    refs[b] = orc.bf16_round(rb)
for bm, bn, cluster in [(0, 0, 0), (256, 0, 0), (128, 128, 2)] + ([] if quick else [(128, 128, 4), (128, 128, 1), (128, 64, 4), (64, 128, 4), (128, 64, 8)]):
    try:
        K.set_option("cluster", cluster)
        K.set_option("tile_n", bn)
        K.set_option("tile_m", bm)
        K.apply_batch(X, res=R)
        errs = [rel(R[b].double().cpu().numpy(), refs[b]) for b in refs]
        ms = timed(lambda: K.apply_batch(X, res=R), 50, warmup=5)
        f = K.flops(64)
        print(json.dumps({"case": "EXTRA: 64 right-hand sides in one launch", "tile_m": bm or "auto", "tile_n": bn or "auto", "cluster": cluster or "auto",
                          "ms": round(ms, 5), "max_rel_err": max(errs), "TFLOPs_algorithmic": round(f / (ms * 1e-3) / 1e12, 1),
                          "TFLOPs_issued_(phase 1 runs hi+lo)": round(1.5 * f / (ms * 1e-3) / 1e12, 1),
                          "frac_of_measured_bf16_burst": round(f / (ms * 1e-3) / 1e12 / PEAK, 4)}), flush=True)
    except Exception as e:
        print(json.dumps({"batch": 64, "tile_m": bm, "tile_n": bn, "cluster": cluster, "error": str(e)[:300]}), flush=True)
K.set_option("tile_m", 0)
K.set_option("cluster", 0)
K.set_option("tile_n", 0)
# ---- how the two kernels scale with the number of right-hand sides (pair kernel: 256-row units on 74 CTA pairs)
for nbv in (16, 32, 64, 74, 148, 296):
    Xv = to_dev(mk_np(15, (nbv, m * m)))
    Rv = torch.empty((nbv, m * m), dtype=torch.bfloat16, device="cuda")
    row = {"case": "EXTRA: right-hand-side sweep", "nb": nbv}
    for name, bm in (("pair_256", 256), ("pair_256_tma_stores", 256), ("single_cta_128", 128)):
        try:
            K.set_option("pair_tma_stores", 1 if name.endswith("tma_stores") else 0)
            K.set_option("tile_m", bm)
            K.set_option("tile_n", 128 if bm == 128 else 0)
            K.set_option("cluster", 2 if bm == 128 else 0)
            ms = timed(lambda: K.apply_batch(Xv, res=Rv), 30, warmup=5)
            row[name] = {"us": round(ms * 1e3, 2), "TFLOPs_algorithmic": round(K.flops(nbv) / (ms * 1e-3) / 1e12, 1)}
        except Exception as e:
            row[name] = {"error": str(e)[:200]}
    print(json.dumps(row), flush=True)
K.set_option("tile_m", 0)
K.set_option("cluster", 0)
K.set_option("tile_n", 0)
Af, Bf, Xf = A.float(), B.float(), x.float().reshape(m, m).t()   # X = reshape(x, q, n) column-major
ms = timed(lambda: (Bf @ Xf) @ Af.t(), 200)
print(json.dumps({"case": "for scale: torch fp32 (B@X)@A.T via cuBLAS, 2 launches, back-to-back period", "us": round(ms * 1e3, 2)}), flush=True)
Ab, Bb, Xb = A, B, x.reshape(m, m).t().contiguous()
ms = timed(lambda: (Bb @ Xb) @ Ab.t(), 200)
print(json.dumps({"case": "for scale: torch bf16 (B@X)@A.T via cuBLAS, 2 launches, back-to-back period", "us": round(ms * 1e3, 2)}), flush=True)
