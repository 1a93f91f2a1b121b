// b2o_ctx.cu -- context, errors, memory helpers, synthetic data, generic small reductions, NCCL glue.
#include "b2o_internal.cuh"
#include <ctype.h>
#include <dlfcn.h>
#include <stdarg.h>
#include <sys/syscall.h>
#include <unistd.h>

// ------------------------------------------------------------------ errors
static thread_local char g_err[1024] = "";
void b2o_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
extern "C" const char *b2o_last_error(void) { return g_err; }
extern "C" int b2o_version(void) { return 100; }  // 0.1.0

// ------------------------------------------------------------------ context
extern "C" int b2o_ctx_create(int device, void *stream, b2o_ctx **out) {
  if (!out) B2O_FAIL(B2O_EARG, "b2o_ctx_create: out is NULL");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    B2O_FAIL(B2O_ECUDA, "b2o_ctx_create: no CUDA device (%s) -- libb2o has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) B2O_FAIL(B2O_EARG, "b2o_ctx_create: device %d out of range [0,%d)", device, ndev);
  B2O_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2O_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) B2O_FAIL(B2O_EUNSUPPORTED, "libb2o is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  b2o_ctx *c = new b2o_ctx_s();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  // stream == NULL is CUDA's legacy default stream (what torch / CUDA.jl use unless told otherwise)
  c->stream = (cudaStream_t)stream;
  c->own_stream = false;
  B2O_CUDA(cudaMalloc(&c->d_partials, sizeof(double) * B2O_MAX_GRID * B2O_MAX_COLS));
  B2O_CUDA(cudaMalloc(&c->d_dots, sizeof(double) * B2O_WS_DOTS));
  B2O_CUDA(cudaMalloc(&c->d_bar, sizeof(unsigned long long) * 8));
  B2O_CUDA(cudaMemset(c->d_bar, 0, sizeof(unsigned long long) * 8));
  B2O_CUDA(cudaMemset(c->d_dots, 0, sizeof(double) * B2O_WS_DOTS));
  B2O_CUDA(cudaMallocHost(&c->h_scal, sizeof(double) * B2O_WS_DOTS));
  B2O_CUDA(cudaEventCreate(&c->ev0));
  B2O_CUDA(cudaEventCreate(&c->ev1));
  *out = c;
  return B2O_OK;
}

extern "C" int b2o_ctx_destroy(b2o_ctx *c) {
  if (!c) return B2O_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  b2o_comm_destroy(c);
  b2o_mbox_disconnect(c);
  if (c->mbox) cudaFree(c->mbox);
  cudaFree(c->d_partials);
  cudaFree(c->d_dots);
  cudaFree(c->d_bar);
  cudaFreeHost(c->h_scal);
  if (c->stage_x) cudaFree(c->stage_x);
  if (c->stage_res) cudaFree(c->stage_res);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  for (int i = 0; i < 16; ++i) {
    if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
  }
  if (c->ev_start) cudaEventDestroy(c->ev_start);
  cudaEventDestroy(c->ev0);
  cudaEventDestroy(c->ev1);
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
  return B2O_OK;
}

extern "C" int b2o_ctx_sync(b2o_ctx *c) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  B2O_CUDA(cudaStreamSynchronize(c->stream));
  return B2O_OK;
}

extern "C" int b2o_ctx_set_option(b2o_ctx *c, const char *key, int64_t value) {
  if (!c || !key) B2O_FAIL(B2O_EARG, "null argument");
  if (!strcmp(key, "tile_rows")) {
    if (value != 1024 && value != 2048 && value != 4096) B2O_FAIL(B2O_EARG, "tile_rows must be 1024, 2048 or 4096");
    c->tile_rows = (int)value;
  } else if (!strcmp(key, "stages")) {
    if (value < 0 || value > 64) B2O_FAIL(B2O_EARG, "stages out of range");
    c->stages = (int)value;
  } else if (!strcmp(key, "grid")) {
    if (value < 0 || value > B2O_MAX_GRID) B2O_FAIL(B2O_EARG, "grid out of range");
    c->grid = (int)value;
  } else if (!strcmp(key, "kron_debug")) {
    c->kron_debug = (int)value;
  } else if (!strcmp(key, "host_chunks")) {
    if (value < 1 || value > 16) B2O_FAIL(B2O_EARG, "host_chunks must be 1..16");
    c->host_chunks = (int)value;
  } else if (!strcmp(key, "graph_jit")) {
    if (value < 0 || value > 2) B2O_FAIL(B2O_EARG, "graph_jit must be 0 (interpreter), 1 (ahead-of-time table, then NVRTC) or 2 (ahead-of-time table only)");
    c->graph_jit = (int)value;
  } else if (!strcmp(key, "graph_interp")) {
    if (value != 0 && value != 2) B2O_FAIL(B2O_EARG, "graph_interp must be 0 (by program size) or 2 (general machine)");
    c->graph_interp = (int)value;
  } else if (!strcmp(key, "dense_scalar")) {
    c->dense_scalar = value != 0;
  } else if (!strcmp(key, "sparse_kernel")) {
    if (value < 0 || value > 3)
      B2O_FAIL(B2O_EARG, "sparse_kernel must be 0 / 3 (pipelined row kernel), 1 (plain row kernel) or 2 (TMA-staged tile kernel)");
    c->sparse_kernel = (int)value;
  } else if (!strcmp(key, "extend_form")) {
    c->extend_form = value != 0;
  } else if (!strcmp(key, "twoloop_block")) {
    c->twoloop_block = value != 0;
  } else if (!strcmp(key, "multi_mma")) {
    c->multi_mma = value != 0;
  } else if (!strcmp(key, "sparse_lanes")) {
    if (value < -1 || value > 5) B2O_FAIL(B2O_EARG, "sparse_lanes must be -1 (auto) or 0..5 (2^k lanes per row)");
    c->sparse_lanes = (int)value;
  } else if (!strcmp(key, "graph_blocks")) {
    if (value < 1 || value > 4) B2O_FAIL(B2O_EARG, "graph_blocks must be 1..4");
    c->graph_blocks = (int)value;
  } else if (!strcmp(key, "time_kernels")) {
    c->time_kernels = value != 0;
  } else if (!strcmp(key, "use_mailbox")) {
    // switch between the in-kernel NVLink mailbox and NCCL for the inner products of a connected context (every rank must switch
    // at the same point of its call sequence; the peers stay mapped and the epoch counter keeps running)
    if (value != 0 && !c->mbox_connected) B2O_FAIL(B2O_ESTATE, "mailbox not connected (b2o_mbox_connect)");
    B2O_CUDA(cudaStreamSynchronize(c->stream));
    c->mbox_ready = value != 0 ? 1 : 0;
  } else if (!strcmp(key, "l2_fetch_granularity")) {
    // cudaLimitMaxL2FetchGranularity (a device-wide hint: 32, 64 or 128 bytes fetched from DRAM per L2 miss); random gathers move
    // fewer unused bytes with a small value, streams are insensitive to it (profiles/r2_index.jsonl)
    if (value != 32 && value != 64 && value != 128) B2O_FAIL(B2O_EARG, "l2_fetch_granularity must be 32, 64 or 128");
    B2O_CUDA(cudaSetDevice(c->device));
    B2O_CUDA(cudaStreamSynchronize(c->stream));
    B2O_CUDA(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
  } else if (!strcmp(key, "numa_local_host")) {
    c->numa_local_host = value != 0;
  } else {
    B2O_FAIL(B2O_EARG, "unknown option '%s'", key);
  }
  return B2O_OK;
}

// raw read of `count` workspace doubles starting at `offset` (debug timelines)
extern "C" int b2o_ctx_debug_read(b2o_ctx *c, int offset, int count, double *out) {
  if (!c || !out || offset < 0 || count < 1 || offset + count > B2O_WS_DOTS) B2O_FAIL(B2O_EARG, "bad argument");
  return b2o_read_scalars(c, c->d_dots + offset, count, out);
}

extern "C" int b2o_ctx_launch_count(b2o_ctx *c, int64_t *out) {
  if (!c || !out) B2O_FAIL(B2O_EARG, "null argument");
  *out = c->launches;
  return B2O_OK;
}

extern "C" int b2o_ctx_kernel_time(b2o_ctx *c, int reset, double *ms_total, int64_t *launches) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  if (ms_total) *ms_total = c->kern_ms;
  if (launches) *launches = c->kern_n;
  if (reset) {
    c->kern_ms = 0.0;
    c->kern_n = 0;
  }
  return B2O_OK;
}

// ------------------------------------------------------------------ memory helpers
extern "C" int b2o_malloc(b2o_ctx *c, size_t bytes, void **dptr) {
  if (!c || !dptr) B2O_FAIL(B2O_EARG, "null argument");
  B2O_CUDA(cudaSetDevice(c->device));
  cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 16);
  if (e != cudaSuccess) {
    cudaGetLastError();
    B2O_FAIL(B2O_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  }
  return B2O_OK;
}
extern "C" int b2o_free(b2o_ctx *c, void *dptr) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  B2O_CUDA(cudaStreamSynchronize(c->stream));
  B2O_CUDA(cudaFree(dptr));
  return B2O_OK;
}
// NUMA node the GPU hangs off (sysfs; -1 = unknown / single node).  Pinned staging buffers that live on the GPU's own node keep
// the H2D / D2H streams of 8 ranks off the inter-socket link (the round-1 end-to-end collapse at N=8).
static int gpu_numa_node(int device) {
  char bus[32] = {0};
  if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  for (char *p = bus; *p; ++p) *p = (char)tolower(*p);
  char path[128];
  snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
  FILE *f = fopen(path, "r");
  if (!f) return -1;
  int node = -1;
  if (fscanf(f, "%d", &node) != 1) node = -1;
  fclose(f);
  return node;
}
extern "C" int b2o_ctx_numa_node(b2o_ctx *c, int *node) {
  if (!c || !node) B2O_FAIL(B2O_EARG, "null argument");
  *node = gpu_numa_node(c->device);
  return B2O_OK;
}
extern "C" int b2o_host_alloc(b2o_ctx *c, size_t bytes, void **hptr) {
  if (!c || !hptr) B2O_FAIL(B2O_EARG, "null argument");
  // MPOL_PREFERRED on the GPU's node for the duration of the allocation: cudaMallocHost populates and pins the pages in this
  // thread, so they land on that node when it has room (and anywhere otherwise -- never a failure)
  const int node = c->numa_local_host ? gpu_numa_node(c->device) : -1;
  bool policy_set = false;
  if (node >= 0 && node < 1024) {
    unsigned long mask[16] = {0};
    mask[node / (8 * sizeof(unsigned long))] |= 1UL << (node % (8 * sizeof(unsigned long)));
    policy_set = syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, (unsigned long)(sizeof(mask) * 8)) == 0;
  }
  cudaError_t e = cudaMallocHost(hptr, bytes ? bytes : 16);
  if (policy_set) syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0UL);
  if (e != cudaSuccess) {
    cudaGetLastError();
    B2O_FAIL(B2O_ENOMEM, "cudaMallocHost(%zu) failed: %s", bytes, cudaGetErrorString(e));
  }
  return B2O_OK;
}
extern "C" int b2o_host_free(b2o_ctx *c, void *hptr) {
  (void)c;
  B2O_CUDA(cudaFreeHost(hptr));
  return B2O_OK;
}
extern "C" int b2o_memcpy_h2d(b2o_ctx *c, void *dst, const void *src, size_t bytes) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  B2O_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  B2O_CUDA(cudaStreamSynchronize(c->stream));
  return B2O_OK;
}
extern "C" int b2o_memcpy_d2h(b2o_ctx *c, void *dst, const void *src, size_t bytes) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  B2O_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  B2O_CUDA(cudaStreamSynchronize(c->stream));
  return B2O_OK;
}
extern "C" int b2o_memset_zero(b2o_ctx *c, void *dptr, size_t bytes) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  B2O_CUDA(cudaMemsetAsync(dptr, 0, bytes, c->stream));
  return B2O_OK;
}

int b2o_read_scalars(b2o_ctx *c, const double *d_src, int count, double *h_dst) {
  if (count > B2O_WS_DOTS) B2O_FAIL(B2O_EARG, "too many scalars");
  B2O_CUDA(cudaMemcpyAsync(c->h_scal, d_src, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
  B2O_CUDA(cudaStreamSynchronize(c->stream));
  memcpy(h_dst, c->h_scal, sizeof(double) * count);
  return B2O_OK;
}

// ------------------------------------------------------------------ synthetic data
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(uint64_t seed, uint64_t i) {
  uint64_t z = mix64((i + 1) * 0x9E3779B97F4A7C15ULL + seed * 0xD1B54A32D192ED03ULL);
  return (double)(z >> 11) * 0x1.0p-53;
}
template <typename T>
__global__ void fill_uniform_kernel(T *x, int64_t n, uint64_t seed, double lo, double w) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    x[i] = (T)__dadd_rn(lo, __dmul_rn(w, u01(seed, (uint64_t)i)));
}
extern "C" int b2o_fill_uniform(b2o_ctx *c, int dtype, void *dptr, int64_t n, uint64_t seed, double lo, double hi) {
  if (!c || (!dptr && n > 0)) B2O_FAIL(B2O_EARG, "null argument");
  if (n <= 0) return B2O_OK;
  int blocks = (int)((n + 255) / 256 < (int64_t)c->num_sms * 16 ? (n + 255) / 256 : (int64_t)c->num_sms * 16);
  if (dtype == B2O_F64)
    fill_uniform_kernel<double><<<blocks, 256, 0, c->stream>>>((double *)dptr, n, seed, lo, hi - lo);
  else if (dtype == B2O_F32)
    fill_uniform_kernel<float><<<blocks, 256, 0, c->stream>>>((float *)dptr, n, seed, lo, hi - lo);
  else
    B2O_FAIL(B2O_EUNSUPPORTED, "fill_uniform: dtype %d", dtype);
  c->launches++;
  B2O_CUDA(cudaGetLastError());
  return B2O_OK;
}

// ------------------------------------------------------------------ generic pair dots
// out[p] = sum_i u_p[i]*v_p[i] for up to 8 pairs in one pass; deterministic: fixed grid, per-block
// partials, last-arriving block sums them in block order.
struct PairDotArgs {
  const double *u[8];
  const double *v[8];
  int npairs;
  int vec;                     // every vector 16-byte aligned: vector loads
  int64_t n;
  double *partials;            // [grid][8]
  double *out;                 // [8]
  unsigned long long *arrive;  // counter, reset by the finishing block
};
__global__ void __launch_bounds__(256) pair_dots_kernel(PairDotArgs a) {
  __shared__ double sred[8][8];
  __shared__ bool is_last;
  double acc[8];
#pragma unroll
  for (int p = 0; p < 8; ++p) acc[p] = 0.0;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (a.vec) {
    // 16-byte loads, two row pairs per pair of vectors in flight per thread (round 1: scalar loads, 3.6 TB/s)
    const int64_t npair = a.n >> 1;
    // many pairs (update_gram: 6-8): one row pair per trip keeps the loads of a trip within the register budget (two per trip
    // measured 25 % slower there: 128 live registers of loads); few pairs: two row pairs in flight
    const int64_t step = a.npairs > 4 ? stride : 2 * stride;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npair; i += step) {
      const int64_t j = i + stride;
      const bool hj = a.npairs <= 4 && j < npair;
#pragma unroll
      for (int p = 0; p < 8; ++p)
        if (p < a.npairs) {
          const double2 u0 = ldg_stream2(a.u[p] + 2 * i);
          const double2 v0 = a.v[p] ? ldg_stream2(a.v[p] + 2 * i) : make_double2(1.0, 1.0);
          double2 u1 = make_double2(0.0, 0.0), v1 = make_double2(0.0, 0.0);
          if (hj) {
            u1 = ldg_stream2(a.u[p] + 2 * j);
            v1 = a.v[p] ? ldg_stream2(a.v[p] + 2 * j) : make_double2(1.0, 1.0);
          }
          acc[p] = fma(u0.x, v0.x, acc[p]);
          acc[p] = fma(u0.y, v0.y, acc[p]);
          acc[p] = fma(u1.x, v1.x, acc[p]);
          acc[p] = fma(u1.y, v1.y, acc[p]);
        }
    }
    if ((a.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      const int64_t i = a.n - 1;
#pragma unroll
      for (int p = 0; p < 8; ++p)
        if (p < a.npairs) acc[p] = fma(a.u[p][i], a.v[p] ? a.v[p][i] : 1.0, acc[p]);
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
#pragma unroll
      for (int p = 0; p < 8; ++p)
        if (p < a.npairs) acc[p] = fma(a.u[p][i], a.v[p] ? a.v[p][i] : 1.0, acc[p]);   // v == null: plain sum
    }
  }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < 8; ++p) {
    double s = warp_sum(acc[p]);
    if (lane == 0) sred[p][warp] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sred[threadIdx.x][w];
    a.partials[(size_t)blockIdx.x * 8 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = atomicAdd(a.arrive, 1ULL);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (warp < 8 && warp < a.npairs) {
      double s = 0.0;
      for (int b = lane; b < gridDim.x; b += 32) s += a.partials[(size_t)b * 8 + warp];
      s = warp_sum(s);
      if (lane == 0) a.out[warp] = s;
    }
    if (threadIdx.x == 0) *a.arrive = 0ULL;
  }
}

int b2o_pair_dots(b2o_ctx *c, int npairs, const double *const *u, const double *const *v, int64_t n, double *d_out) {
  if (npairs < 1 || npairs > 8) B2O_FAIL(B2O_EARG, "pair_dots: 1..8 pairs");
  PairDotArgs a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < npairs; ++p) {
    a.u[p] = u[p];
    a.v[p] = v[p];
  }
  a.npairs = npairs;
  a.n = n;
  a.vec = 1;
  for (int p = 0; p < npairs; ++p)
    if (((uintptr_t)u[p] % 16) || (v[p] && ((uintptr_t)v[p] % 16))) a.vec = 0;
  a.partials = c->d_partials;
  a.out = d_out;
  a.arrive = c->d_bar + 1;
  int64_t want = (n + 255) / 256;
  int grid = (int)(want < 1 ? 1 : (want > (int64_t)c->num_sms * 6 ? (int64_t)c->num_sms * 6 : want));
  pair_dots_kernel<<<grid, 256, 0, c->stream>>>(a);
  c->launches++;
  B2O_CUDA(cudaGetLastError());
  B2O_TRY(b2o_allreduce_sum_f64(c, d_out, npairs));
  return B2O_OK;
}

extern "C" int b2o_dot(b2o_ctx *c, int dtype, const void *a, const void *b, int64_t n, double *out) {
  if (!c || !out) B2O_FAIL(B2O_EARG, "null argument");
  B2O_TRY(b2o_check_dtype_f64(dtype));
  const double *u[1] = {(const double *)a}, *v[1] = {(const double *)b};
  B2O_TRY(b2o_pair_dots(c, 1, u, v, n, c->d_dots + 256));
  return b2o_read_scalars(c, c->d_dots + 256, 1, out);
}

// ------------------------------------------------------------------ NCCL (resolved lazily; libb2o loads without it)
typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*fn_ncclGetUniqueId)(nccl_uid_t *);
typedef int (*fn_ncclCommInitRank)(void **, int, nccl_uid_t, int);
typedef int (*fn_ncclCommDestroy)(void *);
typedef int (*fn_ncclAllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*fn_ncclGetErrorString)(int);
static struct {
  void *lib = nullptr;
  fn_ncclGetUniqueId GetUniqueId = nullptr;
  fn_ncclCommInitRank CommInitRank = nullptr;
  fn_ncclCommDestroy CommDestroy = nullptr;
  fn_ncclAllReduce AllReduce = nullptr;
  fn_ncclGetErrorString GetErrorString = nullptr;
} g_nccl;

static int nccl_load() {
  if (g_nccl.lib) return B2O_OK;
  // If the host process already holds an NCCL (torch ships one) the SONAME lookup reuses it.
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) B2O_FAIL(B2O_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
  g_nccl.GetUniqueId = (fn_ncclGetUniqueId)dlsym(h, "ncclGetUniqueId");
  g_nccl.CommInitRank = (fn_ncclCommInitRank)dlsym(h, "ncclCommInitRank");
  g_nccl.CommDestroy = (fn_ncclCommDestroy)dlsym(h, "ncclCommDestroy");
  g_nccl.AllReduce = (fn_ncclAllReduce)dlsym(h, "ncclAllReduce");
  g_nccl.GetErrorString = (fn_ncclGetErrorString)dlsym(h, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce)
    B2O_FAIL(B2O_ENCCL, "libnccl is missing required symbols");
  g_nccl.lib = h;
  return B2O_OK;
}
#define B2O_NCCL(expr)                                                                            \
  do {                                                                                            \
    int _r = (expr);                                                                              \
    if (_r != 0) {                                                                                \
      b2o_set_error("NCCL error %d (%s) at %s:%d", _r,                                            \
                    g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?", __FILE__, __LINE__); \
      return B2O_ENCCL;                                                                           \
    }                                                                                             \
  } while (0)

extern "C" int b2o_comm_unique_id(void *id128) {
  if (!id128) B2O_FAIL(B2O_EARG, "null argument");
  B2O_TRY(nccl_load());
  nccl_uid_t id;
  B2O_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, 128);
  return B2O_OK;
}
extern "C" int b2o_comm_init(b2o_ctx *c, const void *id128, int nranks, int rank) {
  if (!c || !id128) B2O_FAIL(B2O_EARG, "null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) B2O_FAIL(B2O_EARG, "bad rank %d / nranks %d", rank, nranks);
  B2O_TRY(nccl_load());
  B2O_CUDA(cudaSetDevice(c->device));
  nccl_uid_t id;
  memcpy(&id, id128, 128);
  void *comm = nullptr;
  B2O_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  c->nccl_comm = comm;
  c->nranks = nranks;
  c->rank = rank;
  return B2O_OK;
}
extern "C" int b2o_comm_destroy(b2o_ctx *c) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  if (c->nccl_comm && g_nccl.CommDestroy) {
    cudaStreamSynchronize(c->stream);
    g_nccl.CommDestroy(c->nccl_comm);
  }
  c->nccl_comm = nullptr;
  c->nranks = 1;
  c->rank = 0;
  return B2O_OK;
}
// ------------------------------------------------------------------ NVLink peer mailbox (CUDA IPC between the per-GPU processes)
extern "C" int b2o_mbox_local_handle(b2o_ctx *c, void *handle64) {
  if (!c || !handle64) B2O_FAIL(B2O_EARG, "null argument");
  B2O_CUDA(cudaSetDevice(c->device));
  if (!c->mbox) {
    B2O_CUDA(cudaMalloc(&c->mbox, MBOX_BYTES));
    B2O_CUDA(cudaMemset(c->mbox, 0, MBOX_BYTES));
  }
  cudaIpcMemHandle_t h;
  B2O_CUDA(cudaIpcGetMemHandle(&h, c->mbox));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  return B2O_OK;
}
// handles: nranks x 64 bytes in rank order (this rank's own entry is ignored)
extern "C" int b2o_mbox_connect(b2o_ctx *c, const void *handles, int nranks, int rank) {
  if (!c || !handles) B2O_FAIL(B2O_EARG, "null argument");
  if (nranks < 1 || nranks > MBOX_MAXR || rank < 0 || rank >= nranks) B2O_FAIL(B2O_EARG, "mailbox supports up to %d ranks", MBOX_MAXR);
  if (!c->mbox) B2O_FAIL(B2O_EARG, "call b2o_mbox_local_handle first");
  B2O_CUDA(cudaSetDevice(c->device));
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) {
      c->mbox_peers[r] = c->mbox;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles + 64 * r, 64);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      B2O_FAIL(B2O_ECUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
    }
    c->mbox_peers[r] = p;
  }
  c->nranks = nranks;
  c->rank = rank;
  c->mbox_epoch = 0;
  c->mbox_ready = 1;
  c->mbox_connected = 1;
  return B2O_OK;
}
extern "C" int b2o_mbox_disconnect(b2o_ctx *c) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  if (c->mbox_connected) {
    cudaStreamSynchronize(c->stream);
    for (int r = 0; r < c->nranks; ++r)
      if (r != c->rank && c->mbox_peers[r]) cudaIpcCloseMemHandle(c->mbox_peers[r]);
  }
  c->mbox_ready = 0;
  c->mbox_connected = 0;
  return B2O_OK;
}
void b2o_mbox_fill(b2o_ctx *c, MboxDev *m) {
  memset(m, 0, sizeof(*m));
  m->nranks = 1;
  m->ready = c->d_bar + 2;
  if (!c->mbox_ready || c->nranks <= 1) return;
  m->nranks = c->nranks;
  m->rank = c->rank;
  for (int r = 0; r < c->nranks; ++r) {
    m->vals[r] = (double *)c->mbox_peers[r];
    m->flags[r] = (unsigned long long *)((char *)c->mbox_peers[r] + MBOX_FLAGS_OFF);
  }
  m->epoch_base = c->mbox_epoch;
}

int b2o_allreduce_sum_f64(b2o_ctx *c, double *dptr, int count) {
  if (c->nranks <= 1 || !c->nccl_comm) return B2O_OK;
  // ncclFloat64 = 8, ncclSum = 0
  B2O_NCCL(g_nccl.AllReduce(dptr, dptr, (size_t)count, 8, 0, c->nccl_comm, c->stream));
  return B2O_OK;
}
