#!/usr/bin/env python
"""One profiled launch of the sparse-matrix kernels (ncu --profile-from-start off; cudaProfilerStart/Stop window below):
pipelined row kernel (the default), plain row kernel and TMA-staged tile kernel, Float64, on (a) the 5-point Laplacian of a 2048 x 2048 grid (local gathers) and
(b) 2^21 rows x 24 entries, 12 banded + 12 random columns (scattered gathers)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import linearoperators_jl_b200 as lo  # noqa: E402


def laplace2d(g, dtype, dev):
    nr = g * g
    ii = torch.arange(nr, device=dev, dtype=torch.int64)
    gx, gy = ii % g, ii // g
    cand = torch.stack([ii - g, ii - 1, ii, ii + 1, ii + g], dim=1)
    ok = torch.stack([gy > 0, gx > 0, gx >= 0, gx < g - 1, gy < g - 1], dim=1)
    vals_full = torch.tensor([-1.0, -1.0, 4.0, -1.0, -1.0], device=dev, dtype=dtype).repeat(nr, 1)
    crow = torch.zeros(nr + 1, device=dev, dtype=torch.int64)
    crow[1:] = torch.cumsum(ok.sum(dim=1), 0)
    return torch.sparse_csr_tensor(crow, cand[ok], vals_full[ok], size=(nr, nr), device=dev), nr


def band_rand(nr, per_row, dtype, dev):
    gen = torch.Generator(device=dev).manual_seed(11)
    rows = torch.arange(nr, device=dev, dtype=torch.int64)
    band = (rows[:, None] + torch.arange(-6, 6, device=dev)[None, :]) % nr
    rnd = torch.randint(0, nr, (nr, per_row - 12), generator=gen, device=dev, dtype=torch.int64)
    cols = torch.sort(torch.cat([band, rnd], dim=1), dim=1).values.reshape(-1)
    vals = (torch.rand(nr * per_row, generator=gen, device=dev, dtype=torch.float64) * 2 - 1).to(dtype)
    crow = torch.arange(0, nr * per_row + 1, per_row, device=dev, dtype=torch.int64)
    return torch.sparse_csr_tensor(crow, cols, vals, size=(nr, nr), device=dev), nr


def main():
    ctx = lo.default_context(0)
    dev, dtype = "cuda", torch.float64
    todo = []
    for M, nr in (laplace2d(2048, dtype, dev), band_rand(1 << 21, 24, dtype, dev)):
        op = lo.LinearOperator(M)
        todo.append((op, torch.rand(nr, dtype=dtype, device=dev), torch.empty(nr, dtype=dtype, device=dev)))
    for kern in (0, 1, 2):
        ctx.set_option("sparse_kernel", kern)
        for op, v, r in todo:
            lo.mul_(r, op, v)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for kern in (0, 1, 2):                   # launch order: pipe/laplace, pipe/band, rows/laplace, rows/band, tiles/laplace, tiles/band
        ctx.set_option("sparse_kernel", kern)
        for op, v, r in todo:
            lo.mul_(r, op, v)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    ctx.set_option("sparse_kernel", 0)
    print("NCU_SPARSE_DONE")


if __name__ == "__main__":
    main()
