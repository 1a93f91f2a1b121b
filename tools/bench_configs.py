#!/usr/bin/env python
"""Measures the BASELINE.json configs other than the bench.py headline, plus the remaining kernels, on ONE B200:
one JSON line per case with the algorithmic-bytes roofline.  Output is copied to profiles/."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import linearoperators_jl_b200 as lo  # noqa: E402

PEAK = 6548.2   # fallback only: MEASURED_PEAKS.json (hbm_gbs) is read below when present
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, steps, warmup=3, flush=None):
    for _ in range(warmup):
        if flush is not None:
            flush.zero_()
        fn()
    torch.cuda.synchronize()
    if flush is None:
        # working set >> L2: time `steps` back-to-back launches between ONE pair of events.  Timing each launch on its own
        # charges the host's call overhead (~20-30 us through the Python mirror) to kernels that finish in ~100 us, because
        # the GPU sits idle between the first event and the launch.
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1) / steps
    tot = 0.0
    for _ in range(steps):
        if flush is not None:
            flush.zero_()          # > L2: evicts the working set between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / steps


def line(name, ms, alg_bytes, **kw):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    d = {"case": name, "ms": round(ms, 5), "alg_bytes": alg_bytes, "GBps": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 3),
         "frac_of_nominal_8TBs": round(gbs / 8000.0, 3)}
    d.update(kw)
    print(json.dumps(d), flush=True)


def main():
    ctx = lo.default_context(0)
    only = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--only=")]
    big = "--small" not in sys.argv
    n = 10**8 if big else 10**6
    flush = torch.empty(256 * 2**20 // 8, dtype=torch.float64, device="cuda")
    if only == ["dense"]:
        # LinearOperator(M) (src/constructors.jl:15-29): matrix-vector kernels, matrix >> L2 so no flush is needed for the
        # large cases; the small ones (test/gpu/nvidia.jl block sizes) are launch-latency bound and flushed
        for dtype in (torch.float64, torch.float32):
            E = 8 if dtype == torch.float64 else 4
            for (m, k) in ((32768, 32768), (8192, 131072), (1 << 22, 32), (32, 1 << 22), (20, 20)):
                A = (torch.rand((k, m), dtype=dtype, device="cuda") * 2 - 1).t()           # column-major m x k
                op = lo.LinearOperator(A)
                v = torch.rand(k, dtype=dtype, device="cuda")
                u = torch.rand(m, dtype=dtype, device="cuda")
                r, rt = torch.empty(m, dtype=dtype, device="cuda"), torch.empty(k, dtype=dtype, device="cuda")
                fl = flush if m * k * E < 2**28 else None
                l0 = ctx.launch_count()
                lo.mul_(r, op, v)
                nl = ctx.launch_count() - l0
                ms = timeit(lambda: lo.mul_(r, op, v), 20, flush=fl)
                line("dense N %dx%d %s" % (m, k, str(dtype).replace("torch.", "")), ms, op.apply_bytes(False), launches=nl)
                l0 = ctx.launch_count()
                lo.mul_(rt, lo.transpose(op), u)
                nl = ctx.launch_count() - l0
                ms = timeit(lambda: lo.mul_(rt, lo.transpose(op), u), 20, flush=fl)
                line("dense T %dx%d %s" % (m, k, str(dtype).replace("torch.", "")), ms, op.apply_bytes(True), launches=nl)
                if m * k * E >= 2**30:                                                   # library GEMV beside it (cuBLAS via torch.mv)
                    ms = timeit(lambda: torch.mv(A, v, out=r), 20)
                    line("cublas gemv N %dx%d %s (torch.mv, for comparison)" % (m, k, str(dtype).replace("torch.", "")), ms, op.apply_bytes(False))
                    ms = timeit(lambda: torch.mv(A.t(), u, out=rt), 20)
                    line("cublas gemv T %dx%d %s (torch.mv, for comparison)" % (m, k, str(dtype).replace("torch.", "")), ms, op.apply_bytes(True))
                del A, op, v, u, r, rt
                torch.cuda.empty_cache()
        return
    if only == ["sparse"]:
        # LinearOperator(M::SparseMatrixCSC) (src/constructors.jl:15-29): compressed-row kernels.  Two patterns far above L2:
        # (a) 2^21 rows x 24 entries (12 banded + 12 uniformly random columns: the gathers from x are sector-granular),
        # (b) the 5-point Laplacian of a 2048 x 2048 grid (local gathers).  cuSPARSE (torch.sparse.mm) beside it.
        dev = "cuda"
        for dtype in (torch.float64, torch.float32):
            tn = str(dtype).replace("torch.", "")
            for pat in ("band12+rand12", "laplace2d"):
                if pat == "laplace2d":
                    g = 2048
                    nr = g * g
                    ii = torch.arange(nr, device=dev, dtype=torch.int64)
                    gx, gy = ii % g, ii // g
                    cand = torch.stack([ii - g, ii - 1, ii, ii + 1, ii + g], dim=1)
                    ok = torch.stack([gy > 0, gx > 0, gx >= 0, gx < g - 1, gy < g - 1], dim=1)
                    vals_full = torch.tensor([-1.0, -1.0, 4.0, -1.0, -1.0], device=dev, dtype=dtype).repeat(nr, 1)
                    crow = torch.zeros(nr + 1, device=dev, dtype=torch.int64)
                    crow[1:] = torch.cumsum(ok.sum(dim=1), 0)
                    cols, vals = cand[ok], vals_full[ok]
                    del ii, gx, gy, cand, ok, vals_full
                else:
                    nr, per_row = 1 << 21, 24
                    gen = torch.Generator(device=dev).manual_seed(11)
                    rows = torch.arange(nr, device=dev, dtype=torch.int64)
                    band = (rows[:, None] + torch.arange(-6, 6, device=dev)[None, :]) % nr
                    rnd = torch.randint(0, nr, (nr, per_row - 12), generator=gen, device=dev, dtype=torch.int64)
                    cols = torch.sort(torch.cat([band, rnd], dim=1), dim=1).values.reshape(-1)
                    vals = (torch.rand(nr * per_row, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(dtype)
                    crow = torch.arange(0, nr * per_row + 1, per_row, device=dev, dtype=torch.int64)
                    del rows, band, rnd
                M = torch.sparse_csr_tensor(crow, cols, vals, size=(nr, nr), device=dev)
                t0 = time.perf_counter()
                op = lo.LinearOperator(M)
                torch.cuda.synchronize()
                t_create = time.perf_counter() - t0
                v = torch.rand(nr, dtype=dtype, device=dev)
                r = torch.empty(nr, dtype=dtype, device=dev)
                nnz = int(vals.numel())
                knames = {1: "row kernel", 2: "tile kernel (TMA-staged)", 3: "pipelined row kernel"}
                auto = max(0, min(5, int(np.ceil(np.log2(max(1.0, np.ceil(np.ceil(nnz / nr) / 4.0)))))))
                for kern in (1, 3, 2):
                    ctx.set_option("sparse_kernel", kern)
                    for trans, tag in ((False, "N"), (True, "T")):
                        o = lo.transpose(op) if trans else op
                        l0 = ctx.launch_count()
                        lo.mul_(r, o, v)
                        nl = ctx.launch_count() - l0
                        ms = timeit(lambda: lo.mul_(r, o, v), 50)
                        line("sparse %s %s rows=%d nnz=%d %s, %s" % (tag, pat, nr, nnz, tn, knames[kern]), ms, op.apply_bytes(trans),
                             launches=nl, create_s=round(t_create, 2), nnz_per_s=round(nnz / (ms * 1e-3), 0), lanes_log2=auto)
                    # lane-group width sweep around the heuristic's choice (2^k lanes per row)
                    for k in sorted({max(0, auto - 2), max(0, auto - 1), min(5, auto + 1), min(5, auto + 2)} - {auto}):
                        ctx.set_option("sparse_lanes", k)
                        ms = timeit(lambda: lo.mul_(r, op, v), 50)
                        line("sparse N %s %s, %s, forced 2^%d lanes per row" % (pat, tn, knames[kern], k), ms, op.apply_bytes(False),
                             lanes_log2=k)
                    ctx.set_option("sparse_lanes", -1)
                ctx.set_option("sparse_kernel", 0)
                v2 = v[:, None].contiguous()
                ms = timeit(lambda: torch.sparse.mm(M, v2), 50)
                line("cusparse N %s %s (torch.sparse.mm, for comparison)" % (pat, tn), ms, op.apply_bytes(False))
                del M, op, v, r, v2, cols, vals, crow
                torch.cuda.empty_cache()
        return
    if only == ["index"]:
        # opRestriction / opExtension at n = 1e8 with k = n/4 random indices (duplicates allowed): gather, and the extension in
        # both forms (one gather-form pass through the inverse map vs memset + scatter)
        k = n // 4
        rng = np.random.default_rng(0)
        idx = rng.integers(1, n + 1, size=k)
        t0 = time.perf_counter()
        P = lo.opRestriction(idx, n)
        t_create = time.perf_counter() - t0
        v, uk = ctx.uniform(n, 7), ctx.uniform(k, 8)
        rk, res = ctx.empty(k), ctx.empty(n)
        line("opRestriction random k=n/4 (gather)", timeit(lambda: lo.mul_(rk, P, v), 20), 24.0 * k, create_s=round(t_create, 2))
        Z = lo.transpose(P)
        for form, name in ((0, "gather form (inverse map, one pass)"), (1, "memset + scatter")):
            ctx.set_option("extend_form", form)
            line("opExtension random k=n/4, %s" % name, timeit(lambda: lo.mul_(res, Z, uk), 20), 8.0 * n + 24.0 * k)
        ctx.set_option("extend_form", 0)
        # the L2 fetch-granularity hint: random 8-byte hits pull in whole lines by default
        m = 10
        B = lo.LBFGSOperator(n, mem=m, ctx=ctx)
        for i in range(m):
            s = ctx.uniform(n, 100 + i)
            lo.push_(B, s, s + 0.1 * ctx.uniform(n, 200 + i))
        del s
        for g in (32, 64, 128):
            ctx.set_option("l2_fetch_granularity", g)
            line("opRestriction random k=n/4 (gather), l2_fetch_granularity=%d" % g, timeit(lambda: lo.mul_(rk, P, v), 20), 24.0 * k)
            line("opExtension random k=n/4 gather form, l2_fetch_granularity=%d" % g, timeit(lambda: lo.mul_(res, Z, uk), 20), 8.0 * n + 24.0 * k)
            line("LBFGS(mem=10) forward apply, l2_fetch_granularity=%d" % g, timeit(lambda: lo.mul_(res, B, v), 20), (4 * m + 3) * 8.0 * n)
        return
    if only == ["fwdc"]:
        v, res = ctx.uniform(n, 7), ctx.empty(n)
        m = 10
        for compact in (False, True):
            B = lo.LBFGSOperator(n, mem=m, compact=compact, ctx=ctx)
            for i in range(m):                      # fill the memory
                s = ctx.uniform(n, 100 + i)
                y = s + 0.1 * ctx.uniform(n, 200 + i)
                lo.push_(B, s, y)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(5):                      # steady state: memory full, every push evicts
                s = ctx.uniform(n, 500 + i)
                y = s + 0.1 * ctx.uniform(n, 600 + i)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                lo.push_(B, s, y)
                torch.cuda.synchronize()
                t0 += time.perf_counter() - t1 - (time.perf_counter() - time.perf_counter())
                if i == 0:
                    tp = 0.0
                tp += time.perf_counter() - t1
            # three more pushes (same pairs for both representations); the reference form takes them through the generic
            # multi-dot + linear-combination passes (push_mode 0) for comparison with the streaming rebuild above
            if not compact:
                B.set_option("push_mode", 0)
            t_g = 0.0
            for i in range(3):
                s = ctx.uniform(n, 700 + i)
                y = s + 0.1 * ctx.uniform(n, 800 + i)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                lo.push_(B, s, y)
                torch.cuda.synchronize()
                t_g += time.perf_counter() - t1
            tp_generic = None if compact else round(t_g / 3 * 1e3, 2)
            if not compact:
                B.set_option("push_mode", 1)
            ms = timeit(lambda: lo.mul_(res, B, v), 20)
            xs = ctx.zeros(n)
            ms_solve = timeit(lambda: lo.solve_shifted_system_(xs, B, v, 0.5), 3, warmup=1)
            if not compact:
                ref = res.clone()
            line("LBFGSOperator(mem=10) %s" % ("compact form (extension)" if compact else "a_k/b_k form (reference algorithm)"), ms,
                 (4 * m + 3) * 8.0 * n, push_ms_steady_state=round(tp / 5 * 1e3, 2), push_ms_generic_passes=tp_generic, solve_shifted_system_ms=round(ms_solve, 2),
                 rel_diff_vs_reference_form=(float(torch.linalg.norm(res - ref) / torch.linalg.norm(ref)) if compact else 0.0))
            del B
            torch.cuda.empty_cache()
        return
    if only == ["lsr1push"]:
        m = 10
        for mode in (1, 0):
            L = lo.LSR1Operator(n, mem=m, ctx=ctx)
            L.set_option("push_mode", mode)
            for i in range(m):
                s = ctx.uniform(n, 300 + i, -1.0, 1.0)
                lo.push_(L, s, 2.0 * s + 0.3 * ctx.uniform(n, 400 + i, -1.0, 1.0))
            tp = 0.0
            for i in range(4):
                s = ctx.uniform(n, 500 + i, -1.0, 1.0)
                y = 2.0 * s + 0.3 * ctx.uniform(n, 600 + i, -1.0, 1.0)
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                lo.push_(L, s, y)
                torch.cuda.synchronize()
                tp += time.perf_counter() - t1
            print(json.dumps({"case": "LSR1Operator(mem=10) push!, steady state, %s" % ("streaming rebuild" if mode else "generic passes"),
                              "ms": round(tp / 4 * 1e3, 2), "accepted": bool(L.last_push_accepted)}), flush=True)
            del L
            torch.cuda.empty_cache()
        return
    if only == ["multi"]:
        # §8f rank 4: mul!(Res, B, X) with nrhs right-hand sides; the state columns are streamed once per 8 of them
        m = 10
        B = lo.LBFGSOperator(n, mem=m, ctx=ctx)
        for i in range(m):
            s = ctx.uniform(n, 100 + i)
            lo.push_(B, s, s + 0.1 * ctx.uniform(n, 200 + i))
        del s
        for k in (1, 2, 4, 8, 16):
            Xb = torch.empty((k, n), dtype=torch.float64, device="cuda")
            for j in range(k):
                Xb[j] = ctx.uniform(n, 300 + j)
            X, Res = Xb.T, torch.empty((k, n), dtype=torch.float64, device="cuda").T
            ms = timeit(lambda: lo.mul_(Res, B, X), 10)
            v, r = ctx.empty(n), ctx.empty(n)
            def loop():
                for j in range(k):
                    lo.mul_(r, B, Xb[j])
            ms_loop = timeit(loop, 5) if k > 1 else ms
            lo.mul_(r, B, Xb[k - 1])
            diff = float(torch.linalg.norm(Res[:, k - 1] - r) / torch.linalg.norm(r))
            passes = (k + 7) // 8
            line("LBFGSOperator(mem=10) block apply, %d right-hand sides" % k, ms, (4 * m * passes + 3 * k) * 8.0 * n,
                 ms_column_by_column=round(ms_loop, 3), speedup=round(ms_loop / ms, 2), vector_equivalent_GBps=round(k * (4 * m + 3) * 8.0 * n / ms / 1e6, 1),
                 rel_diff_vs_vector_apply=diff)
            del Xb, X, Res
            torch.cuda.empty_cache()
        del B
        torch.cuda.empty_cache()
        # block two-loop recursion: matrix right-hand sides of the two-loop InverseLBFGSOperator (src/operations.jl:34-36)
        m = 20
        H = lo.InverseLBFGSOperator(n, mem=m, ctx=ctx)
        for i in range(m):
            s = ctx.uniform(n, 100 + i)
            lo.push_(H, s, s + 0.1 * ctx.uniform(n, 200 + i))
        del s
        for k in (2, 4, 8):
            Xb = torch.empty((k, n), dtype=torch.float64, device="cuda")
            for j in range(k):
                Xb[j] = ctx.uniform(n, 300 + j)
            X, Res = Xb.T, torch.empty((k, n), dtype=torch.float64, device="cuda").T
            ms = timeit(lambda: lo.mul_(Res, H, X), 5)
            ctx.set_option("twoloop_block", 0)
            ms_loop = timeit(lambda: lo.mul_(Res, H, X), 3)
            ctx.set_option("twoloop_block", 1)
            r = ctx.empty(n)
            lo.mul_(Res, H, X)
            lo.mul_(r, H, Xb[k - 1])
            same = bool(torch.equal(Res[:, k - 1], r))
            line("InverseLBFGS(mem=20) two-loop, block recursion, %d right-hand sides" % k, ms, ((4 * k + 4) * m + k) * 8.0 * n,
                 ms_column_by_column=round(ms_loop, 3), speedup=round(ms_loop / ms, 2), vector_equivalent_GBps=round(k * (8 * m + 2) * 8.0 * n / ms / 1e6, 1),
                 bit_identical_to_vector_apply=same)
            del Xb, X, Res
            torch.cuda.empty_cache()
        return
    if only == ["multi8"]:
        # 8 and 16 right-hand sides: the FP64 tensor-core block kernel (DMMA) against the SIMT block kernel
        m = 10
        B = lo.LBFGSOperator(n, mem=m, ctx=ctx)
        for i in range(m):
            s = ctx.uniform(n, 100 + i)
            lo.push_(B, s, s + 0.1 * ctx.uniform(n, 200 + i))
        del s
        for k in (8, 16, 6):
            Xb = torch.empty((k, n), dtype=torch.float64, device="cuda")
            for j in range(k):
                Xb[j] = ctx.uniform(n, 300 + j)
            X, Res = Xb.T, torch.empty((k, n), dtype=torch.float64, device="cuda").T
            r = ctx.empty(n)
            lo.mul_(r, B, Xb[k - 1])
            for mma in (1, 0):
                ctx.set_option("multi_mma", mma)
                ms = timeit(lambda: lo.mul_(Res, B, X), 10)
                diff = float(torch.linalg.norm(Res[:, k - 1] - r) / torch.linalg.norm(r))
                passes = (k + 7) // 8
                line("LBFGSOperator(mem=10) block apply, %d right-hand sides, %s" % (k, "FP64 tensor-core kernel (DMMA)" if mma else "SIMT kernel"), ms,
                     (4 * m * passes + 3 * k) * 8.0 * n, vector_equivalent_GBps=round(k * (4 * m + 3) * 8.0 * n / ms / 1e6, 1), rel_diff_vs_vector_apply=diff)
            ctx.set_option("multi_mma", 1)
            del Xb, X, Res
            torch.cuda.empty_cache()
        return
    if only == ["f32"]:
        # Float32 quasi-Newton operators: the Float32 instantiations of the persistent kernels (half the bytes per row)
        f32 = lambda nn, seed: ctx.fill_uniform(ctx.empty(nn, dtype=torch.float32), seed)
        for nn in (n, 2 * n):
            x, res = f32(nn, 7), ctx.empty(nn, dtype=torch.float32)
            for name, mk, m, per_row in (("LBFGSOperator(Float32, n=%d, mem=10)" % nn, lambda: lo.LBFGSOperator(nn, mem=10, T=torch.float32, ctx=ctx), 10, 4 * 10 + 3),
                                         ("InverseLBFGSOperator(Float32, n=%d, mem=20)" % nn, lambda: lo.InverseLBFGSOperator(nn, mem=20, T=torch.float32, ctx=ctx), 20, 8 * 20 + 2),
                                         ("LSR1Operator(Float32, n=%d, mem=10)" % nn, lambda: lo.LSR1Operator(nn, mem=10, T=torch.float32, ctx=ctx), 10, 2 * 10 + 3)):
                op = mk()
                tpush = 0.0
                for i in range(m):
                    s = f32(nn, 100 + i)
                    y = s + 0.1 * f32(nn, 200 + i) if "LSR1" not in name else ctx.fill_uniform(ctx.empty(nn, dtype=torch.float32), 200 + i, -0.5, 1.0)
                    torch.cuda.synchronize()
                    t1 = time.perf_counter()
                    lo.push_(op, s, y)
                    torch.cuda.synchronize()
                    tpush = time.perf_counter() - t1
                del s, y
                ms = timeit(lambda: lo.mul_(res, op, x), 20)
                line(name + " apply", ms, per_row * 4.0 * nn, applies_per_s=round(1e3 / ms, 1), last_push_ms=round(tpush * 1e3, 2),
                     float64_equivalent_GBps=round(per_row * 8.0 * nn / ms / 1e6, 1))
                del op
                torch.cuda.empty_cache()
            del x, res
            torch.cuda.empty_cache()
        return
    if only == ["invc"]:
        v, res = ctx.uniform(n, 7), ctx.empty(n)
        for m in (10, 20):
            H = lo.InverseLBFGSOperator(n, mem=m, ctx=ctx)
            t0 = time.perf_counter()
            for i in range(m):
                s = ctx.uniform(n, 100 + i)
                y = s + 0.1 * ctx.uniform(n, 200 + i)
                lo.push_(H, s, y)
            del s, y
            line("InverseLBFGS(mem=%d) two-loop (reference algorithm)" % m, timeit(lambda: lo.mul_(res, H, v), 10), (8 * m + 2) * 8.0 * n)
            two = res.clone()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            H.set_option("inverse_mode", 1)
            torch.cuda.synchronize()
            t_gram = time.perf_counter() - t0
            ms = timeit(lambda: lo.mul_(res, H, v), 10)
            diff = float(torch.linalg.norm(res - two) / torch.linalg.norm(two))
            line("InverseLBFGS(mem=%d) compact representation (extension)" % m, ms, (4 * m + 3) * 8.0 * n,
                 two_loop_equivalent_GBps=round((8 * m + 2) * 8.0 * n / ms / 1e6, 1), rel_diff_vs_two_loop=diff, gram_rebuild_s=round(t_gram, 3))
            del H, two
            torch.cuda.empty_cache()
        return
    if only == ["kron"]:
        import oracle as orc
        orc.set_mode(True, 1)
        m = 512
        mk = lambda seed, shape: torch.as_tensor(orc.bf16_round(orc.uniform(int(np.prod(shape)), seed, -1.0, 1.0)).reshape(shape)).cuda().to(torch.bfloat16).contiguous()
        A, B, x = mk(11, (m, m)), mk(12, (m, m)), mk(13, (m * m,))
        K = lo.kron(A, B, max_batch=64, ctx=ctx)
        res = torch.empty(m * m, dtype=torch.bfloat16, device="cuda")
        fl = K.flops()
        for name, fn, f in (("cfg4 kron(A,B)*x 512x512 bf16, one launch (tcgen05 GEMM pair)", lambda: lo.mul_(res, K, x), fl),
                            ("cfg4 transpose(kron)*x", lambda: lo.mul_(res, lo.transpose(K), x), fl)):
            ms = timeit(fn, 200, warmup=10)
            ctx.set_option("time_kernels", 1)
            ctx.kernel_time(reset=True)
            for _ in range(100):
                fn()
            kms, kn = ctx.kernel_time(reset=True)
            ctx.set_option("time_kernels", 0)
            kms /= max(kn, 1)
            print(json.dumps({"case": name, "ms_call_to_completion": round(ms, 5), "ms_kernel_events": round(kms, 5), "flops": f,
                              "TFLOPs_kernel": round(f / (kms * 1e-3) / 1e12, 2),
                              "frac_of_measured_bf16_peak": round(f / (kms * 1e-3) / 1e12 / 1686.8, 4)}), flush=True)
        X = mk(14, (64, m * m))
        R = torch.empty((64, m * m), dtype=torch.bfloat16, device="cuda")
        ms = timeit(lambda: K.apply_batch(X, res=R), 100, warmup=5)
        f = K.flops(64)
        print(json.dumps({"case": "EXTRA (not a reference feature): 64 right-hand sides in one launch", "ms": round(ms, 5), "flops": f,
                          "TFLOPs": round(f / (ms * 1e-3) / 1e12, 2), "frac_of_measured_bf16_peak": round(f / (ms * 1e-3) / 1e12 / 1686.8, 4)}), flush=True)
        Af, Bf, Xf = A.float(), B.float(), x.float().reshape(m, m).t()   # X = reshape(x, q, n) column-major
        ms = timeit(lambda: (Bf @ Xf) @ Af.t(), 100, warmup=5)
        print(json.dumps({"case": "for scale: torch fp32 (B@X)@A.T via cuBLAS, 2 launches", "ms": round(ms, 5)}), flush=True)
        return
    if only == ["cfg3"]:
        v, res = ctx.uniform(n, 2), ctx.empty(n)
        h = ctx.uniform(n, 3)
        h /= float(np.sqrt(ctx.dot(h, h)))
        dd = ctx.uniform(n, 4, 0.5, 1.5)
        tree = lo.opHouseholder(h) * lo.opDiagonal(dd) + 0.1 * lo.opEye(n)
        fused = lo.fuse(tree)
        line("cfg3 closure tree", timeit(lambda: lo.mul_(res, tree, v), 30), 56.0 * n)
        line("cfg3 fused, ONE launch, NVRTC-specialised", timeit(lambda: lo.mul_(res, fused, v), 50), 56.0 * n, jit=fused.info()["jit"])
        line("cfg3 fused 5-arg beta=0.5, NVRTC-specialised", timeit(lambda: lo.mul_(res, fused, v, 2.0, 0.5), 50), 64.0 * n)
        # transpose(H*D + 0.1 I) = D*H + 0.1 I: the dot reads h, v only -> 6n*8 bytes (b2o_graph_info), not 7n*8
        tb = fused.info(transposed=True)["alg_bytes"]
        line("cfg3 fused transpose, NVRTC-specialised", timeit(lambda: lo.mul_(res, lo.transpose(fused), v), 50), tb)
        ctx.set_option("graph_jit", 2)       # ahead-of-time table only: what a deployment without libnvrtc runs
        fused2 = lo.fuse(tree)
        line("cfg3 fused, NO NVRTC: ahead-of-time instantiation compiled into libb2o", timeit(lambda: lo.mul_(res, fused2, v), 50), 56.0 * n,
             executor=fused2.info()["executor"])
        line("cfg3 fused 5-arg beta=0.5, NO NVRTC: ahead-of-time instantiation", timeit(lambda: lo.mul_(res, fused2, v, 2.0, 0.5), 50), 64.0 * n,
             executor=fused2.info(beta=0.5)["executor"])
        line("cfg3 fused transpose, NO NVRTC: ahead-of-time instantiation", timeit(lambda: lo.mul_(res, lo.transpose(fused2), v), 50), tb,
             executor=fused2.info(transposed=True)["executor"])
        ctx.set_option("graph_jit", 0)
        for interp, name in ((0, "interpreter, 4 rows per dispatch (small-program machine)"), (2, "interpreter, general machine (round-1 form)")):
            ctx.set_option("graph_interp", interp)
            for gb in (2, 3):
                ctx.set_option("graph_blocks", gb)
                line("cfg3 fused, NO NVRTC: %s, graph_blocks=%d" % (name, gb), timeit(lambda: lo.mul_(res, fused, v), 30), 56.0 * n)
                line("cfg3 fused 5-arg beta=0.5, NO NVRTC: %s, graph_blocks=%d" % (name, gb), timeit(lambda: lo.mul_(res, fused, v, 2.0, 0.5), 30), 64.0 * n)
            line("cfg3 fused transpose, NO NVRTC: %s" % name, timeit(lambda: lo.mul_(res, lo.transpose(fused), v), 30), tb)
        ctx.set_option("graph_interp", 0)
        ctx.set_option("graph_blocks", 3)
        ctx.set_option("graph_jit", 1)
        return
    # cfg1: opDiagonal(n=1e6) * v  (24 MB: L2 resident unless flushed)
    n1 = 10**6
    d, v, res = ctx.uniform(n1, 1), ctx.uniform(n1, 2), ctx.empty(n1)
    D = lo.opDiagonal(d)
    line("cfg1 opDiagonal(n=1e6)*v, L2 flushed between iterations", timeit(lambda: lo.mul_(res, D, v), 50, flush=flush), 24.0 * n1)
    line("cfg1 opDiagonal(n=1e6)*v, L2 warm", timeit(lambda: lo.mul_(res, D, v), 200), 24.0 * n1)
    # elementwise leaves at n=1e8
    d, v, res = ctx.uniform(n, 1), ctx.uniform(n, 2), ctx.empty(n)
    D, E, Z, O = lo.opDiagonal(d), lo.opEye(n), lo.opZeros(n, n), lo.opOnes(n, n)
    line("opDiagonal(n=%g)*v" % n, timeit(lambda: lo.mul_(res, D, v), 30), 24.0 * n)
    line("opDiagonal 5-arg (alpha=2,beta=2)", timeit(lambda: lo.mul_(res, D, v, 2.0, 2.0), 30), 32.0 * n)
    line("opEye(n)*v", timeit(lambda: lo.mul_(res, E, v), 30), 16.0 * n)
    line("opZeros(n)*v", timeit(lambda: lo.mul_(res, Z, v), 30), 8.0 * n)
    line("opOnes(n,n)*v (sum + fill: 2 launches)", timeit(lambda: lo.mul_(res, O, v), 30), 16.0 * n)
    h = ctx.uniform(n, 3)
    h /= float(np.sqrt(ctx.dot(h, h)))
    H = lo.opHouseholder(h)
    line("opHouseholder(h)*v (one cooperative launch)", timeit(lambda: lo.mul_(res, H, v), 30), 40.0 * n)
    # cfg3
    dd = ctx.uniform(n, 4, 0.5, 1.5)
    tree = lo.opHouseholder(h) * lo.opDiagonal(dd) + 0.1 * lo.opEye(n)
    fused = lo.fuse(tree)
    l0 = ctx.launch_count()
    lo.mul_(res, tree, v)
    tree_launches = ctx.launch_count() - l0
    line("cfg3 (opHouseholder*opDiagonal + 0.1*opEye)*v, closure tree", timeit(lambda: lo.mul_(res, tree, v), 30), 56.0 * n,
         launches_per_apply=tree_launches, reference_traffic_bytes=88.0 * n)
    line("cfg3 fused, ONE launch, NVRTC-specialised", timeit(lambda: lo.mul_(res, fused, v), 30), 56.0 * n, launches_per_apply=1,
         jit=fused.info()["jit"])
    ctx.set_option("graph_jit", 0)
    line("cfg3 fused, ONE launch, interpreter kernel", timeit(lambda: lo.mul_(res, fused, v), 30), 56.0 * n, launches_per_apply=1)
    ctx.set_option("graph_jit", 1)
    del tree, fused, H, dd, h
    # gather / scatter
    k = n // 4
    idx = torch.randint(1, n + 1, (k,), device="cuda", dtype=torch.int64).cpu().numpy()
    P = lo.opRestriction(idx, n)
    rk, uk = ctx.empty(k), ctx.uniform(k, 9)
    line("opRestriction random k=n/4 (gather)", timeit(lambda: lo.mul_(rk, P, v), 20), 24.0 * k, note="random 8-byte gathers are sector (32 B) granular")
    line("opExtension random k=n/4 (memset + scatter)", timeit(lambda: lo.mul_(res, lo.transpose(P), uk), 20), 8.0 * n + 24.0 * k)
    idx = np.arange(1, n + 1, 4, dtype=np.int64)
    P = lo.opRestriction(idx, n)
    rk = ctx.empty(idx.shape[0])
    line("opRestriction 1:4:n (strided gather)", timeit(lambda: lo.mul_(rk, P, v), 20), 24.0 * idx.shape[0])
    del P, rk, uk, idx, d, D
    torch.cuda.empty_cache()
    # quasi-Newton applies
    for name, mk, m, npr in (("LSR1Operator(n, mem=10)", lambda: lo.LSR1Operator(n, mem=10, ctx=ctx), 10, 2 * 10 + 3),
                             ("cfg5 per-GPU slab: InverseLBFGSOperator(n, mem=20)", lambda: lo.InverseLBFGSOperator(n, mem=20, ctx=ctx), 20, 8 * 20 + 2),
                             ("cfg2 LBFGSOperator(n, mem=10)", lambda: lo.LBFGSOperator(n, mem=10, ctx=ctx), 10, 4 * 10 + 3)):
        op = mk()
        t0 = time.perf_counter()
        for i in range(m):
            if "LSR1" in name:
                s = ctx.uniform(n, 300 + i, -1.0, 1.0)
                y = 2.0 * s + 0.3 * ctx.uniform(n, 400 + i, -1.0, 1.0)
            else:
                s = ctx.uniform(n, 100 + i)
                y = s + 0.1 * ctx.uniform(n, 200 + i)
            lo.push_(op, s, y)
            assert op.last_push_accepted
        torch.cuda.synchronize()
        push_s = (time.perf_counter() - t0) / m
        del s, y
        line(name + " apply", timeit(lambda: lo.mul_(res, op, v), 20), npr * 8.0 * n, push_seconds_avg=round(push_s, 4))
        line(name + " apply 5-arg beta=0.5", timeit(lambda: lo.mul_(res, op, v, 2.0, 0.5), 10), (npr + 1) * 8.0 * n)
        del op
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
