#!/usr/bin/env python
"""One profiled launch of each dense-matrix kernel (N / T, Float64 / Float32) on a 32768 x 32768 matrix
(ncu --profile-from-start off; the cudaProfilerStart/Stop window below)."""
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
    for op, v, u, r, rt in todo:       # warm-up
        lo.mul_(r, op, v)
        lo.mul_(rt, lo.transpose(op), u)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for op, v, u, r, rt in todo:
        lo.mul_(r, op, v)
        lo.mul_(rt, lo.transpose(op), u)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("NCU_DENSE_DONE")


if __name__ == "__main__":
    main()
