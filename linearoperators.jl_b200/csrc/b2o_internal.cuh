// b2o_internal.cuh -- shared host/device plumbing of libb2o (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/b2o.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb2o is written for sm_100a (B200) only"
#endif

// ------------------------------------------------------------------ errors
void b2o_set_error(const char *fmt, ...);
#define B2O_FAIL(code, ...)      \
  do {                           \
    b2o_set_error(__VA_ARGS__);  \
    return (code);               \
  } while (0)
#define B2O_CUDA(expr)                                                                          \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      b2o_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,     \
                    cudaGetErrorString(_e));                                                    \
      return B2O_ECUDA;                                                                         \
    }                                                                                           \
  } while (0)
#define B2O_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != B2O_OK) return _s; \
  } while (0)

// kernel launch spelled as a macro so that tests/emu can run the same launch logic under the host SIMT emulator
#define B2O_STREAM_T cudaStream_t
#define B2O_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)

// ------------------------------------------------------------------ context
#include "b2o_shared_defs.h"

struct b2o_ctx_s {
  int device = 0;
  int num_sms = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  // workspace (allocated once at create; nothing allocates per call)
  double *d_partials = nullptr;             // [B2O_MAX_GRID][B2O_MAX_COLS]
  double *d_dots = nullptr;                 // [B2O_WS_DOTS] reduced scalars
  unsigned long long *d_bar = nullptr;      // [0] grid-barrier counter (monotonic), [1] arrival counter (split mode)
  unsigned long long bar_base = 0;          // host mirror of d_bar[0]
  double *h_scal = nullptr;                 // pinned, B2O_WS_DOTS doubles
  // staging for *_host entry points (grown on first use, then reused)
  void *stage_x = nullptr, *stage_res = nullptr;
  size_t stage_bytes = 0;
  // host-buffer pipeline: copy-in / copy-out streams and per-chunk events (created on first use)
  cudaStream_t s_in = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[16] = {}, ev_k[16] = {}, ev_start = nullptr;
  int host_chunks = 8;
  // tuning
  int tile_rows = 2048;
  int stages = 0;        // 0 -> as many as shared memory allows
  int grid = 0;          // 0 -> one CTA per SM
  int kron_debug = 0;    // record a %globaltimer timeline of CTA 0 of the kron kernel into d_dots[448..464)
  int graph_jit = 1;     // fused trees: use the NVRTC-specialised kernel when NVRTC + driver are present (else the interpreter)
  int graph_interp = 0;  // fused trees, interpreter: 0 = pick the machine by program size, 2 = always the general (2 rows per dispatch) one
  int graph_blocks = 3;  // resident CTAs per SM the fused-graph kernel is compiled for (occupancy hides the dispatch latency)
  int dense_scalar = 0;  // dense-matrix leaf: force the scalar (unvectorised) kernels (testing)
  int sparse_kernel = 0; // sparse-matrix leaf: 0 / 3 software-pipelined row kernel (default), 1 plain row kernel, 2 TMA-staged tile kernel
  int extend_form = 0;   // opExtension: 0 gather form through the inverse map when the index set is dense enough, 1 always memset + scatter
  int multi_mma = 0;     // block apply with 5..8 right-hand sides: 0 = the SIMT kernel (default: measured faster), 1 = FP64 tensor-core kernel (DMMA)
  int twoloop_block = 1; // matrix right-hand sides of the two-loop inverse: 1 = block recursion (4 / 8 columns per sweep), 0 = column by column
  int sparse_lanes = -1; // sparse-matrix leaf: force 2^k lanes per row (k = 0..5), -1 = from the mean row length
  // accounting
  int64_t launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool time_kernels = false;
  double kern_ms = 0.0;
  int64_t kern_n = 0;
  // multi-GPU (row partition); comm is an ncclComm_t, resolved lazily through dlopen
  void *nccl_comm = nullptr;
  int nranks = 1, rank = 0;
  // NVLink peer mailbox (in-kernel all-reduce of the small dot vectors): local buffer + IPC-mapped peers
  void *mbox = nullptr;
  void *mbox_peers[8] = {};
  int mbox_ready = 0;      // in use (option "use_mailbox" toggles it on a connected context)
  int mbox_connected = 0;  // peers mapped
  int numa_local_host = 1; // b2o_host_alloc prefers the GPU's own NUMA node
  unsigned long long mbox_epoch = 0;
};

void b2o_mbox_fill(b2o_ctx *ctx, MboxDev *m);   // nranks = 1 when the mailbox is not connected

int b2o_allreduce_sum_f64(b2o_ctx *ctx, double *dptr, int count);  // no-op when nranks == 1

static inline int b2o_check_dtype_f64(int dtype) {
  if (dtype != B2O_F64) B2O_FAIL(B2O_EUNSUPPORTED, "dtype %d not supported by this entry point (Float64 only)", dtype);
  return B2O_OK;
}

// generic helpers implemented in b2o_ctx.cu / b2o_util.cu
int b2o_pair_dots(b2o_ctx *ctx, int npairs, const double *const *u, const double *const *v, int64_t n, double *d_out);
int b2o_read_scalars(b2o_ctx *ctx, const double *d_src, int count, double *h_dst);  // D2H + sync via pinned

#ifdef __CUDACC__
// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// orders this thread's generic-proxy shared-memory accesses before later async-proxy (bulk copy) accesses to the same bytes
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// streaming (evict-first) 16-byte global load/store for user vectors
__device__ __forceinline__ double2 ldg_stream2(const double *p) {
  double2 r;
  asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream2(double *p, double2 v) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// grid-wide barrier for co-resident (cooperatively launched) persistent kernels.  `ctr` only ever
// grows; `target` = value it must reach.  Every thread of every CTA calls it.
__device__ __forceinline__ void grid_barrier(unsigned long long *ctr, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1ULL);
    while (ld_acquire_u64(ctr) < target) { __nanosleep(32); }
    __threadfence();
  }
  __syncthreads();
}

// ---- NVLink peer mailbox: all-reduce (sum) of nv <= 128 doubles held in shared memory, executed by ONE warp per GPU.
// Every rank stores its values into its slot of every peer's mailbox (plain stores over NVLink), publishes them with a
// system-scope release of the epoch number, waits for the peers' epochs, and sums the slots in rank order -- the same order
// on every GPU, so all ranks obtain bit-identical results.  Slots are double-buffered by epoch parity: a rank can be at most
// one all-reduce ahead of the slowest peer.
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void mbox_allreduce_warp(const MboxDev &mb, unsigned long long epoch, double *vals, int nv) {
  const int lane = threadIdx.x & 31;
  const size_t slot = ((size_t)(epoch & 1ULL) * MBOX_MAXR + mb.rank) * MBOX_MAXV;
  for (int r = 0; r < mb.nranks; ++r)
    for (int v = lane; v < nv; v += 32) mb.vals[r][slot + v] = vals[v];
  __threadfence_system();
  __syncwarp();
  if (lane < mb.nranks) st_release_sys_u64(mb.flags[lane] + mb.rank, epoch);
  if (lane < mb.nranks)
    while (ld_acquire_sys_u64(mb.flags[mb.rank] + lane) < epoch) { __nanosleep(20); }
  __syncwarp();
  const double *mine = mb.vals[mb.rank] + (size_t)(epoch & 1ULL) * MBOX_MAXR * MBOX_MAXV;
  for (int v = lane; v < nv; v += 32) {
    double s = 0.0;
    for (int r = 0; r < mb.nranks; ++r) s += ld_relaxed_sys_f64(mine + (size_t)r * MBOX_MAXV + v);
    vals[v] = s;
  }
  __syncwarp();
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif  // __CUDACC__
