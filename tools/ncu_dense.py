#!/usr/bin/env python
"""One profiled launch of each matrix-leaf kernel (ncu --profile-from-start off; the cudaProfilerStart/Stop window below):
dense N / T, Float64 / Float32, on a 32768 x 32768 matrix; sparse N / T on 2^21 rows x 24 entries (12 banded + 12 random)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import linearoperators_jl_b200 as lo  # noqa: E402


def main():
    lo.default_context(0)
    todo = []
    for dtype in (torch.float64, torch.float32):
        m = k = 32768
        A = (torch.rand((k, m), dtype=dtype, device="cuda") * 2 - 1).t()
        op = lo.LinearOperator(A)
        v, u = torch.rand(k, dtype=dtype, device="cuda"), torch.rand(m, dtype=dtype, device="cuda")
        r, rt = torch.empty(m, dtype=dtype, device="cuda"), torch.empty(k, dtype=dtype, device="cuda")
        todo.append((op, v, u, r, rt))
    dev = "cuda"
    for dtype in (torch.float64, torch.float32):
        nr, per_row = 1 << 21, 24
        gen = torch.Generator(device=dev).manual_seed(11)
        rows = torch.arange(nr, device=dev, dtype=torch.int64)
        band = (rows[:, None] + torch.arange(-6, 6, device=dev)[None, :]) % nr
        rnd = torch.randint(0, nr, (nr, per_row - 12), generator=gen, device=dev, dtype=torch.int64)
        cols = torch.sort(torch.cat([band, rnd], dim=1), dim=1).values.reshape(-1)
        vals = (torch.rand(nr * per_row, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(dtype)
        crow = torch.arange(0, nr * per_row + 1, per_row, device=dev, dtype=torch.int64)
        op = lo.LinearOperator(torch.sparse_csr_tensor(crow, cols, vals, size=(nr, nr), device=dev))
        v, u = torch.rand(nr, dtype=dtype, device=dev), torch.rand(nr, dtype=dtype, device=dev)
        todo.append((op, v, u, torch.empty(nr, dtype=dtype, device=dev), torch.empty(nr, dtype=dtype, device=dev)))
        del rows, band, rnd
    for op, v, u, r, rt in todo:       # warm-up
        lo.mul_(r, op, v)
        lo.mul_(rt, lo.transpose(op), u)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for op, v, u, r, rt in todo:
        lo.mul_(r, op, v)
        lo.mul_(rt, lo.transpose(op), u)
    ctx = lo.default_context(0)
    ctx.set_option("sparse_kernel", 1)          # the sparse row kernel beside the (default) tile kernel
    for op, v, u, r, rt in todo[2:]:
        lo.mul_(r, op, v)
    ctx.set_option("sparse_kernel", 0)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("NCU_DENSE_DONE")


if __name__ == "__main__":
    main()
