"""Execution context: one CUDA device + stream + libb2o workspace (b2o_ctx).

torch is used here for what the task allows it for -- device memory (vectors are torch CUDA tensors
whose raw pointers go across the C ABI), streams and torch.distributed rendezvous -- nothing else."""
import ctypes

from . import _lib

_default = {}


def _torch():
    import torch
    return torch


class Context:
    def __init__(self, device=0, stream=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise _lib.B2OError("no CUDA device: the B200 operator-apply engine has no CPU fallback")
        self.lib = _lib.load()
        self.device = int(device)
        torch.cuda.set_device(self.device)
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        self.stream = stream
        h = ctypes.c_void_p()
        _lib.check(self.lib.b2o_ctx_create(self.device, ctypes.c_void_p(stream.cuda_stream), ctypes.byref(h)))
        self.handle = h
        self.nranks, self.rank = 1, 0
        self.mailbox = False

    # -- plumbing
    def sync(self):
        _lib.check(self.lib.b2o_ctx_sync(self.handle))

    def set_option(self, key, value):
        _lib.check(self.lib.b2o_ctx_set_option(self.handle, key.encode(), int(value)))

    def launch_count(self):
        n = ctypes.c_int64()
        _lib.check(self.lib.b2o_ctx_launch_count(self.handle, ctypes.byref(n)))
        return n.value

    def kernel_time(self, reset=False):
        ms, n = ctypes.c_double(), ctypes.c_int64()
        _lib.check(self.lib.b2o_ctx_kernel_time(self.handle, int(reset), ctypes.byref(ms), ctypes.byref(n)))
        return ms.value, n.value

    def debug_read(self, offset, count):
        """raw read of workspace scalars: [0, ncols) the reduced inner products of the last forward / L-SR1 / compact apply,
        [512, 512+2A) those of the last fused two-loop apply in sweep order, [768, 771) the mailbox timing accumulators
        (ns CTA 0 waited for its own GPU's CTAs, ns in the NVLink exchange, number of exchanges)."""
        out = (ctypes.c_double * int(count))()
        _lib.check(self.lib.b2o_ctx_debug_read(self.handle, int(offset), int(count), out))
        return list(out)

    def empty(self, n, dtype=None):
        torch = _torch()
        return torch.empty(int(n), dtype=dtype or torch.float64, device="cuda:%d" % self.device)

    def host_empty(self, n):
        """pinned Float64 host vector from b2o_host_alloc (placed on the GPU's own NUMA node when the platform says which),
        as a torch CPU tensor over that memory.  Freed with the context (or host_free)."""
        torch = _torch()
        p = ctypes.c_void_p()
        _lib.check(self.lib.b2o_host_alloc(self.handle, ctypes.c_size_t(int(n) * 8), ctypes.byref(p)))
        buf = (ctypes.c_double * int(n)).from_address(p.value)
        t = torch.frombuffer(buf, dtype=torch.float64, count=int(n))
        self._host_bufs = getattr(self, "_host_bufs", [])
        self._host_bufs.append((p, buf))
        return t

    def numa_node(self):
        node = ctypes.c_int(-1)
        _lib.check(self.lib.b2o_ctx_numa_node(self.handle, ctypes.byref(node)))
        return node.value

    def zeros(self, n, dtype=None):
        torch = _torch()
        return torch.zeros(int(n), dtype=dtype or torch.float64, device="cuda:%d" % self.device)

    def fill_uniform(self, x, seed, lo=0.0, hi=1.0):
        """x[i] = lo + (hi-lo)*u(seed,i) with the generator shared with the oracle."""
        torch = _torch()
        dt = _lib.B2O_F64 if x.dtype == torch.float64 else _lib.B2O_F32
        _lib.check(self.lib.b2o_fill_uniform(self.handle, dt, ctypes.c_void_p(x.data_ptr()), x.numel(), int(seed),
                                             float(lo), float(hi)))
        return x

    def uniform(self, n, seed, lo=0.0, hi=1.0):
        return self.fill_uniform(self.empty(n), seed, lo, hi)

    def dot(self, a, b):
        out = ctypes.c_double()
        _lib.check(self.lib.b2o_dot(self.handle, _lib.B2O_F64, ctypes.c_void_p(a.data_ptr()),
                                    ctypes.c_void_p(b.data_ptr()), a.numel(), ctypes.byref(out)))
        return out.value

    # -- row-partitioned multi-GPU: torch.distributed is only the rendezvous for the NCCL unique id
    def init_comm_from_torch(self, group=None):
        import torch.distributed as dist
        from .partition import broadcast_bytes
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        buf = (ctypes.c_ubyte * 128)()
        if rank == 0:
            _lib.check(self.lib.b2o_comm_unique_id(buf))
        raw = broadcast_bytes(bytes(buf), 128, group, self.device)
        idbuf = (ctypes.c_ubyte * 128).from_buffer_copy(raw)
        _lib.check(self.lib.b2o_comm_init(self.handle, idbuf, world, rank))
        self.nranks, self.rank = world, rank

    def connect_mailbox(self, group=None):
        """NVLink peer mailbox: exchange CUDA-IPC handles of the per-GPU mailboxes so that the quasi-Newton applies all-reduce
        their dots inside ONE persistent launch per GPU (stores over NVLink) instead of a kernel + NCCL call per inner product."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        mine = (ctypes.c_ubyte * 64)()
        _lib.check(self.lib.b2o_mbox_local_handle(self.handle, mine))
        gathered = [None] * world
        dist.all_gather_object(gathered, bytes(mine), group=group)
        allh = (ctypes.c_ubyte * (64 * world)).from_buffer_copy(b"".join(gathered))
        st = self.lib.b2o_mbox_connect(self.handle, allh, world, rank)
        err = None if st == 0 else _lib.last_error()
        # all-or-nothing: a rank that could not map its peers would otherwise wait on NCCL while the others spin on the mailbox
        oks = [None] * world
        dist.all_gather_object(oks, err, group=group)
        if any(e is not None for e in oks):
            self.lib.b2o_mbox_disconnect(self.handle)
            raise _lib.B2OError("NVLink mailbox unavailable on some rank: %s" % [e for e in oks if e][0])
        dist.barrier(group)
        self.mailbox = True

    def disconnect_mailbox(self):
        _lib.check(self.lib.b2o_mbox_disconnect(self.handle))
        self.mailbox = False

    def close(self):
        for p, _ in getattr(self, "_host_bufs", []):
            self.lib.b2o_host_free(self.handle, p)
        self._host_bufs = []
        if self.handle:
            self.lib.b2o_ctx_destroy(self.handle)
            self.handle = None


def default_context(device=None):
    torch = _torch()
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    ctx = _default.get(device)
    if ctx is None:
        ctx = _default[device] = Context(device)
    return ctx
