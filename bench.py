#!/usr/bin/env python
"""bench.py -- the BASELINE.json headline: LBFGSOperator(n=1e8, mem=10) Float64 `op*v` on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n ROWS] [--mem M] [--workload fwd|inv]

A "step" is ONE apply `mul!(res, B, x, 1, 0)` (one persistent kernel launch) on synthetic state: `mem` accepted
pushes of s_i ~ U[0,1)^n (seed 100+i), y_i = s_i + 0.1 U[0,1)^n (seed 200+i), x ~ U[0,1) seed 7.
`value` is whole-job algorithmic GB/s = (4m+3)*8*n bytes * applies/s summed over ranks (weak scaling: every rank
holds n rows; the 2m dots are all-reduced inside the kernel over the NVLink peer mailbox).  Inputs are 16 GB >> the
126 MB L2, so no L2 flush is needed between iterations.  `e2e` runs the same apply through the host-buffer C-ABI
entry (pinned host x -> H2D -> kernel -> D2H res) inside the timed region.  `cpu_baseline` / `--impl reference` time
the CPU restatement of the reference (oracle/, "port": Julia is not installed anywhere) at the SAME n when host
memory allows (the n actually timed is printed in config.workload), with every CPU this process may use.

Under torchrun (N > 1) the line also carries
  "parity": a small row-partitioned LBFGS / InverseLBFGS push!+apply checked against the oracle on rank 0, and the
            mailbox all-reduce compared with NCCL's (results and inner products),
  "cfg5":   BASELINE config 5, InverseLBFGSOperator(n = N x 1e8, mem=20) row-partitioned: two-loop with the mailbox
            (one launch per GPU), with NCCL (one launch + one all-reduce per inner product), and the compact form,
  "per_rank_ms", "mailbox": per-rank device times and the time CTA 0 spent waiting / exchanging per step."""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

METRIC = "LBFGS(m=10,n=1e8) op*v algorithmic GB/s"
NOMINAL_HBM_GBS = 8000.0
FALLBACK_HBM_GBS = 6650.0


def alg_bytes(n, m, inverse=False, compact=False):
    """forward (4m+3)*8*n, inverse two-loop (8m+2)*8*n (beta = 0)."""
    return ((8 * m + 2) if (inverse and not compact) else (4 * m + 3)) * 8.0 * n


def workload_name(n, mem, inverse):
    return "%sLBFGSOperator(n=%d, mem=%d) Float64 apply, alpha=1 beta=0" % ("Inverse" if inverse else "", n, mem)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        self.n0 = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = self.rows[getattr(self, "n0", 0):] or self.rows
        sm = sorted(float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "power_w_max": max(pw) if pw else None}


def host_mem_available():
    """bytes this process may still allocate: MemAvailable capped by the cgroup limit (containers)."""
    avail = None
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                avail = int(line.split()[1]) * 1024
    except Exception:
        pass
    try:
        lim = open("/sys/fs/cgroup/memory.max").read().strip()
        cur = int(open("/sys/fs/cgroup/memory.current").read().strip())
        if lim != "max":
            room = int(lim) - cur
            avail = room if avail is None else min(avail, room)
    except Exception:
        pass
    return avail


def cpu_reference(n_want, mem, steps, warmup, inverse=False, budget_s=90.0, single_thread=True):
    """Reference arm: the CPU restatement of lbfgs_multiply (src/lbfgs.jl:117-154 / :173-202), one pass per reference
    statement, every CPU this process may use (affinity mask and cgroup quota; OMP_NUM_THREADS -- which torchrun sets to 1
    -- is deliberately ignored: rank 0 is the only rank that runs this), at n_want rows when the host has the memory
    (state columns filled directly -- the apply cost is data-independent), else the largest n that fits."""
    import numpy as np
    import oracle
    oracle.build()
    avail = oracle.usable_cpus()
    need = lambda n: (2 * mem + 4) * 8.0 * n * 1.15          # touched pages: 2m columns + x, res, q (+ slack)
    room = host_mem_available()
    n_cpu = n_want
    while room is not None and need(n_cpu) > room and n_cpu > 10**6:
        n_cpu //= 2
    op = None
    while op is None:
        try:
            op = oracle.LBFGS(n_cpu, mem=mem, inverse=inverse)
            x = np.empty(n_cpu)
            res = np.empty(n_cpu)
        except MemoryError:
            op, n_cpu = None, n_cpu // 2
    oracle.set_mode(False, avail)
    for k in range(mem):
        for which in (("s", "y") if inverse else ("a", "b")):
            oracle.lib().orc_fill_uniform(op.col(which, k).ctypes.data, n_cpu, 1000 + 10 * k + ord(which[0]), 0.0, 1.0e-3)
        op.ys[k] = 1.0
    op.set_state(1, 0.5)
    oracle.lib().orc_fill_uniform(x.ctypes.data, n_cpu, 7, 0.0, 1.0)
    # "all the host threads it can use": calibrate (all / half of the usable CPUs -- SMT siblings make "all" slower on
    # some hosts) with one apply each, then time the best
    best_t, best_dt = avail, None
    for t in sorted({t for t in (avail, avail // 2) if 1 <= t <= avail}, reverse=True):
        oracle.set_mode(False, t)
        dts = []
        for _ in range(2):
            t0 = time.perf_counter()
            op.apply(x, res=res)
            dts.append(time.perf_counter() - t0)
            if dts[-1] > 10.0:
                break
        if best_dt is None or min(dts) < best_dt:
            best_t, best_dt = t, min(dts)
    threads = best_t
    dt1 = None
    if single_thread:
        # SURVEY §8d figure (i): reference-faithful single thread (Julia broadcasts and one BLAS thread), one apply of a
        # 1/8 row slice (the single-thread rate does not depend on n)
        n1 = max(1, n_cpu // 8)
        op1 = oracle.LBFGS(n1, mem=mem, inverse=inverse)
        for k in range(mem):
            for which in (("s", "y") if inverse else ("a", "b")):
                op1.col(which, k)[:] = op.col(which, k)[:n1]
            op1.ys[k] = 1.0
        op1.set_state(1, 0.5)
        oracle.set_mode(False, 1)
        r1 = np.empty(n1)
        t0 = time.perf_counter()
        op1.apply(x[:n1], res=r1)
        dt1 = time.perf_counter() - t0
        del op1, r1
    oracle.set_mode(False, threads)
    note = ""
    if best_dt * (steps + warmup) > budget_s:
        steps_new = max(1, int(budget_s / best_dt) - 1)
        note = "; %d of the %d requested steps timed to stay inside %d s" % (steps_new, steps, budget_s)
        steps, warmup = steps_new, min(warmup, 1)
    for _ in range(warmup):
        op.apply(x, res=res)
    t0 = time.perf_counter()
    for _ in range(steps):
        op.apply(x, res=res)
    dt = (time.perf_counter() - t0) / steps
    oracle.set_mode(True, 1)
    out = {"value": alg_bytes(n_cpu, mem, inverse) / dt / 1e9, "unit": "GB/s", "cores": threads, "kind": "port", "n": n_cpu,
           "sample": "%s apply at n=%d rows (%s), mem=%d, %d timed applies, %.3f s/apply, %d threads of %d usable CPUs "
                     "(faster of all/half)%s" % ("inverse" if inverse else "forward", n_cpu,
                                                "the full workload" if n_cpu == n_want else "1/%.1f of the workload: host memory" % (n_want / n_cpu),
                                                mem, steps, dt, threads, avail, note),
           "applies_per_s": 1.0 / dt, "ms_per_apply": dt * 1e3, "steps_timed": steps}
    if dt1 is not None:
        out["single_thread"] = {"value": alg_bytes(n1, mem, inverse) / dt1 / 1e9, "unit": "GB/s", "cores": 1, "rows": n1,
                                "note": "same restatement on one thread (the reference's broadcasts are single-threaded)"}
    return out


# ----------------------------------------------------------------------------------------------------------------------
def multi_gpu_parity(lo, ctx, dist, rank, world, have_mailbox):
    """Row-partitioned push! + apply on a small problem (n = world x 1e5, mem = 5) against the oracle on rank 0, forward and
    inverse; then the same applies with NCCL instead of the mailbox: results and inner products compared bit for bit."""
    import numpy as np
    import torch
    n_loc, mem = 100_000, 5
    n = n_loc * world
    out = {"n": n, "mem": mem, "ranks": world}
    worst, eq_res, eq_dots, rel_mn = 0.0, True, True, 0.0

    def gather(t):
        buf = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(buf, t)
        return torch.cat(buf).cpu().numpy()

    for inverse in (False, True):
        g = lo.LBFGSOperator(n_loc, mem=mem, inverse=inverse, ctx=ctx)
        pairs = []
        for i in range(mem + 2):
            s = ctx.uniform(n_loc, 1000 * rank + 100 + i)
            y = s + 0.1 * ctx.uniform(n_loc, 1000 * rank + 200 + i)
            lo.push_(g, s, y)
            pairs.append((gather(s), gather(y)))
        x = ctx.uniform(n_loc, 1000 * rank + 7)
        xs = gather(x)
        res = {}
        dots = {}
        for mode in (["mailbox", "nccl"] if have_mailbox else ["nccl"]):
            if have_mailbox:
                ctx.set_option("use_mailbox", 1 if mode == "mailbox" else 0)
            r = g * x
            dots[mode] = ctx.debug_read(512 if inverse else 0, 2 * mem) if (mode == "mailbox" or not inverse) else None
            res[mode] = gather(r)
        if have_mailbox:
            ctx.set_option("use_mailbox", 1)
        if rank == 0:
            import oracle
            oracle.build()
            oracle.set_mode(True, 1)
            o = oracle.LBFGS(n, mem=mem, inverse=inverse)
            for s, y in pairs:
                o.push(s, y)
            ref = o.apply(xs)
            for mode, r in res.items():
                worst = max(worst, float(np.linalg.norm(r - ref) / np.linalg.norm(ref)))
            if have_mailbox:
                eq_res &= bool(np.array_equal(res["mailbox"], res["nccl"]))
                rel_mn = max(rel_mn, float(np.linalg.norm(res["mailbox"] - res["nccl"]) / np.linalg.norm(ref)))
                if dots["nccl"] is not None:
                    eq_dots &= dots["mailbox"] == dots["nccl"]
        del g
    out["multi_gpu_rel_err"] = worst
    out["tolerance"] = 1e-12
    out["ok"] = worst <= 1e-12
    if have_mailbox:
        out["mailbox_eq_nccl"] = eq_res
        out["mailbox_dots_eq_nccl"] = eq_dots
        out["mailbox_vs_nccl_rel"] = rel_mn
        out["note"] = ("mailbox sums the ranks' partials in rank order on every GPU; NCCL's order is its own (ring/tree), "
                       "so bitwise equality is guaranteed only at 2 ranks")
    return out


def time_applies(lo, ctx, op, res, x, steps, warmup, barrier, world, dist):
    """W untimed + K timed applies between two events on the context stream; returns (max-over-ranks ms per apply,
    per-rank ms per apply, launches per apply, mailbox wait / exchange us per apply on this rank)."""
    import torch
    for _ in range(warmup):
        lo.mul_(res, op, x)
    barrier()
    lo.mul_(res, op, x)        # untimed: the ranks meet on the DEVICE here, so host launch skew stays out of the events
    d0 = ctx.debug_read(768, 3)
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        lo.mul_(res, op, x)
    e1.record()
    barrier()
    launches = (ctx.launch_count() - l0) / steps
    d1 = ctx.debug_read(768, 3)
    ms = e0.elapsed_time(e1) / steps
    per_rank = [ms]
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        buf = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(buf, t)
        per_rank = [float(b.item()) for b in buf]
    epochs = max(1.0, d1[2] - d0[2])
    mb = {"wait_local_ctas_us_per_apply": (d1[0] - d0[0]) / 1e3 / steps, "exchange_us_per_apply": (d1[1] - d0[1]) / 1e3 / steps,
          "exchanges_per_apply": (d1[2] - d0[2]) / steps, "exchange_us_each": (d1[1] - d0[1]) / 1e3 / epochs} if d1[2] > d0[2] else None
    return max(per_rank), per_rank, launches, mb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=10**8, help="rows per GPU")
    ap.add_argument("--mem", type=int, default=10)
    ap.add_argument("--workload", default="fwd", choices=["fwd", "inv"])
    ap.add_argument("--cpu-n", type=int, default=0, help="rows of the CPU arm (0 = same as --n when host memory allows)")
    ap.add_argument("--tile-rows", type=int, default=0)
    ap.add_argument("--stages", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-cfg5", action="store_true")
    ap.add_argument("--no-mailbox", action="store_true", help="multi-GPU: NCCL all-reduce per inner product instead of the NVLink mailbox")
    args = ap.parse_args()
    inverse = args.workload == "inv"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3)
    cpu_n = args.cpu_n or args.n

    if args.impl == "reference":
        if rank != 0:
            return
        cb = cpu_reference(cpu_n, args.mem, max(1, args.steps), max(1, min(args.warmup, 5)), inverse, single_thread=False)
        config = {"workload": workload_name(cb["n"], args.mem, inverse), "rows": cb["n"], "mem": args.mem,
                  "parallelism": "host CPUs, %d threads" % cb["cores"],
                  "l2": "inputs (%.1f GB of columns) >> any cache" % (2 * args.mem * cb["n"] * 8 / 1e9)}
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "GB/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_apply"],
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                          "config": config, "cpu_baseline": cb,
                          "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    import torch
    import linearoperators_jl_b200 as lo
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = lo.default_context(local_rank)
    have_mailbox = False
    if world > 1:
        ctx.init_comm_from_torch()
        try:
            ctx.connect_mailbox()          # in-kernel all-reduce over NVLink peer memory; NCCL stays for push!
            have_mailbox = True
        except Exception as e:              # e.g. IPC not permitted: the NCCL path (kernel + all-reduce per dot) is used
            if rank == 0:
                print("mailbox unavailable, using NCCL per inner product: %s" % e, file=sys.stderr)
    if args.tile_rows:
        ctx.set_option("tile_rows", args.tile_rows)
    if args.stages:
        ctx.set_option("stages", args.stages)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    parity = multi_gpu_parity(lo, ctx, dist, rank, world, have_mailbox) if world > 1 else None
    use_mailbox = have_mailbox and not args.no_mailbox
    if have_mailbox:
        ctx.set_option("use_mailbox", 1 if use_mailbox else 0)
    config = {"workload": workload_name(args.n, args.mem, inverse),
              "rows_per_gpu": args.n, "mem": args.mem,
              "parallelism": ("row-partition x%d, dots all-reduced %s" % (world, "in-kernel over the NVLink peer mailbox" if use_mailbox else "over NCCL")) if world > 1 else "single GPU",
              "l2": "inputs (%.1f GB of columns) >> 126 MB L2, no flush needed" % (2 * args.mem * args.n * 8 / 1e9)}

    n, m = args.n, args.mem

    def build_op(mem, inv):
        op = lo.LBFGSOperator(n, mem=mem, inverse=inv, ctx=ctx)
        for i in range(mem):
            s = ctx.uniform(n, 1000 * rank + 100 + i)
            y = s + 0.1 * ctx.uniform(n, 1000 * rank + 200 + i)
            lo.push_(op, s, y)
            assert op.last_push_accepted
        return op

    B = build_op(m, inverse)
    x = ctx.uniform(n, 1000 * rank + 7)
    res = ctx.empty(n)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                     # before the barrier: forking nvidia-smi must not skew rank 0 against the others
    for _ in range(args.warmup):
        lo.mul_(res, B, x)
    barrier()
    if rank == 0:
        sampler.mark()
    ms_step, per_rank_ms, launches_per_step, mb = time_applies(lo, ctx, B, res, x, args.steps, 0, barrier, world, dist)
    clocks = sampler.stop() if rank == 0 else None
    bytes_step = alg_bytes(n, m, inverse)
    value = world * bytes_step / (ms_step * 1e-3) / 1e9

    # ---- end to end through the host-buffer C-ABI entry: pinned host x -> H2D -> apply -> D2H res, every step.
    # The pinned buffers come from b2o_host_alloc (on the GPU's own NUMA node when the platform exposes it).
    xh = ctx.host_empty(n)
    rh = ctx.host_empty(n)
    xh.copy_(x.cpu())
    e2e_steps = max(3, min(args.steps, 10))
    B.apply_host(rh, xh)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        B.apply_host(rh, xh)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    dlt = rh.to(res.device) - res
    e2e_rel = float((dlt.norm() / res.norm()).item())
    e2e_ndiff = int((dlt != 0).sum().item())
    del dlt
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": world * bytes_step / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * 8,
           "ms_per_step": e2e_s * 1e3, "applies_per_s": world / e2e_s, "rel_diff_vs_device_apply": e2e_rel, "rows_differing_from_device_apply": e2e_ndiff,
           "host_buffers": "pinned, b2o_host_alloc, NUMA node %d" % ctx.numa_node()}

    # ---- BASELINE config 5: InverseLBFGSOperator(n = world x 1e8 rows, mem = 20), row-partitioned
    cfg5 = None
    if not args.no_cfg5 and not inverse and n == 10**8:
        del B
        m5 = 20
        H = build_op(m5, True)
        b5 = alg_bytes(n, m5, True)
        peak, _ = measured_peak()
        cfg5 = {"workload": "InverseLBFGSOperator(n=%d, mem=%d) Float64 apply, %d rows per GPU x %d GPUs" % (n * world, m5, n, world),
                "algorithmic_bytes_per_gpu": b5}

        def leg(name, nbytes, steps):
            ms, per_rank, lps, mbx = time_applies(lo, ctx, H, res, x, steps, 2, barrier, world, dist)
            cfg5[name] = {"ms_per_apply": ms, "gbs_per_gpu": nbytes / (ms * 1e-3) / 1e9, "frac_of_measured_peak": nbytes / (ms * 1e-3) / 1e9 / peak,
                          "total_gbs": world * nbytes / (ms * 1e-3) / 1e9, "launches_per_apply": lps, "per_rank_ms": per_rank}
            if mbx:
                cfg5[name]["mailbox"] = mbx

        leg("two_loop_mailbox" if use_mailbox else ("two_loop_nccl" if world > 1 else "two_loop"), b5, 5)
        if use_mailbox:
            ctx.set_option("use_mailbox", 0)
            leg("two_loop_nccl", b5, 3)
            ctx.set_option("use_mailbox", 1)
        ref_res = res.clone()
        H.set_option("inverse_mode", 1)
        leg("compact", alg_bytes(n, m5, True, compact=True), 5)
        dlt = res - ref_res
        cfg5["compact"]["rel_diff_vs_two_loop"] = float((dlt.norm() / ref_res.norm()).item())
        cfg5["compact"]["rows_differing_from_two_loop"] = int((dlt != 0).sum().item())
        cfg5["compact"]["result_norm_this_rank"] = float(ref_res.norm().item())
        del dlt
        cfg5["compact"]["two_loop_equivalent_gbs_per_gpu"] = b5 / (cfg5["compact"]["ms_per_apply"] * 1e-3) / 1e9
        del H, ref_res

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = measured_peak()
    per_gpu = bytes_step / (ms_step * 1e-3) / 1e9
    traffic, traffic_src = None, None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        tp = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                key = "inv" if inverse else "fwd"
                # the ncu capture is for the default shapes (n=1e8; fwd mem=10, inv mem=20): report it only for that workload
                if abs(tj.get("algorithmic", {}).get(key, -1.0) - bytes_step) < 1.0:
                    traffic = tj.get(key)
                    traffic_src = "profiles/%s: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch of this kernel on this workload (static file, not re-measured in this run)" % name
                    break
            except Exception:
                pass
    out = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config, "applies_per_s": world / (ms_step * 1e-3),
        "frac_of_nominal_8TBs_per_gpu": per_gpu / NOMINAL_HBM_GBS,
        "roofline": {"bound": "hbm", "achieved": per_gpu, "peak": peak, "unit": "GB/s", "frac": per_gpu / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": "of " + peak_kind,
                     "kernel": "qn_twoloop_kernel<2048, double>" if inverse else "qn_compact_kernel<2048, 0 (LBFGS_FWD), double>",
                     "algorithmic_bytes_per_launch": bytes_step},
        "e2e": e2e, "gpu_launches": int(round(launches_per_step * args.steps)), "clocks": clocks,
        "per_rank_ms": per_rank_ms,
    }
    if mb:
        out["mailbox"] = mb
    if parity is not None:
        out["parity"] = parity
    if cfg5 is not None:
        out["cfg5"] = cfg5
    if not args.no_cpu and world == 1:
        out["cpu_baseline"] = cpu_reference(cpu_n, m, 3, 1, inverse, budget_s=30.0)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
