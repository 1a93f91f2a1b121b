"""world_size-2 tests of the row-partitioned path: gloo on CPU (always), NCCL on >= 2 GPUs (-m gpu)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_workers(backend, nproc, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py"), backend]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DIST_OK %s %d" % (backend, nproc) in r.stdout


def test_row_partition_gloo_world2(orc):
    run_workers("gloo", 2, 29611)


def test_row_slab():
    from linearoperators_jl_b200.partition import row_slab
    for n, w in [(10, 3), (8 * 10**8, 8), (7, 8), (0, 2)]:
        slabs = [row_slab(n, r, w) for r in range(w)]
        assert slabs[0][0] == 0 and slabs[-1][1] == n
        assert all(slabs[i][1] == slabs[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in slabs) - min(b - a for a, b in slabs) <= 1


@pytest.mark.gpu
def test_row_partition_nccl(orc):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    run_workers("nccl", min(torch.cuda.device_count(), 4), 29612)
