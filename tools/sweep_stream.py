#!/usr/bin/env python
"""Sweep the streaming-kernel knobs (tile rows x ring stages) on the cfg2 / cfg5-slab workloads; one process, state built once."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import linearoperators_jl_b200 as lo  # noqa: E402


def main():
    ctx = lo.default_context(0)
    n = 10**8
    steps = 60
    for kind, m, per_row in (("fwd", 10, 4 * 10 + 3), ("inv", 20, 8 * 20 + 2)):
        op = lo.LBFGSOperator(n, mem=m, inverse=kind == "inv", ctx=ctx)
        for i in range(m):
            s = ctx.uniform(n, 100 + i)
            y = s + 0.1 * ctx.uniform(n, 200 + i)
            lo.push_(op, s, y)
        del s, y
        x, res = ctx.uniform(n, 7), ctx.empty(n)
        for tr in (1024, 2048, 4096):
            for st in (3, 4, 5, 6, 7, 8, 10, 12, 16, 24):
                if tr * 8 * st > 200 * 1024:
                    continue
                ctx.set_option("tile_rows", tr)
                ctx.set_option("stages", st)
                for _ in range(3):
                    lo.mul_(res, op, x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps if kind == "fwd" else steps // 3):
                    lo.mul_(res, op, x)
                e1.record()
                e1.synchronize()
                ms = e0.elapsed_time(e1) / (steps if kind == "fwd" else steps // 3)
                print(json.dumps({"kind": kind, "tile_rows": tr, "stages": st, "ms": round(ms, 4),
                                  "GBps": round(per_row * 8.0 * n / ms / 1e6, 1)}), flush=True)
        del op, x, res
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
