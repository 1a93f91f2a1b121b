"""Row partition of vectors over ranks: rank r owns the contiguous slab [lo, hi) of every vector and of
every quasi-Newton state column; elementwise work is local, every inner product is all-reduced."""
import ctypes


def row_slab(n, rank, world):
    """contiguous, balanced slab bounds: the first n % world ranks get one extra row"""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_bytes(payload, nbytes, group=None, device=None):
    """rank 0's `payload` (bytes of length nbytes) to every rank through torch.distributed (gloo or nccl)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    t = torch.zeros(nbytes, dtype=torch.uint8)
    if rank == 0:
        t = torch.tensor(list(payload), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        t = t.cuda(device)
    dist.broadcast(t, src=0, group=group)
    return bytes(t.cpu().tolist())


def allreduce_sum(values, group=None, device=None):
    """sum a small list of Python floats over ranks (host-side twin of the library's NCCL all-reduce)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64)
    if dist.get_backend(group) == "nccl":
        t = t.cuda(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().tolist()
