// b2o_qn.cu -- LBFGSOperator / InverseLBFGSOperator / LSR1Operator handles (src/lbfgs.jl, src/lsr1.jl):
// state, apply (persistent kernels in b2o_qn_kernels.cuh), push!, diag!, reset!, state access.
#include "b2o_qn_multi_mma.cuh"
#include <float.h>
#include <math.h>
#include <algorithm>

constexpr int64_t B2O_PITCH_ALIGN = 4096;
// 227 KB opt-in limit per CTA minus the kernels' static shared memory (flags, a few doubles)
constexpr size_t B2O_MAX_DYN_SMEM = 227 * 1024 - 1024;  // rows; every supported tile size divides it

struct b2o_qn_s {
  b2o_ctx *ctx = nullptr;
  int kind = 0;  // 0 = L-BFGS, 1 = L-SR1
  int esize = 8; // element size: 8 = Float64 (everything below), 4 = Float32 (b2o_qn_f32.inc: the column slabs then hold floats)
  int64_t n = 0, pitch = 0;
  int mem = 1;
  bool scaling = true, damped = false, inverse = false;
  double gamma = 1.0, sigma2 = 0.99, sigma3 = 10.0, opnorm_ub = 1.0;
  double *S = nullptr, *Y = nullptr, *A = nullptr, *B = nullptr;  // [mem][pitch], zero padded
  double *q = nullptr, *tmp = nullptr;                            // [pitch]
  double *shifted_p = nullptr;                                    // [2*mem][pitch] work matrix of solve_shifted_system! (lazy)
  double *q_multi = nullptr;                                      // [q_multi_cols][pitch] work vectors of the block two-loop recursion (lazy)
  int q_multi_cols = 0;
  double *d_alpha = nullptr;                                      // [mem] device copy of data.α (inverse)
  std::vector<double> ys, aux;  // aux: inverse L-BFGS α (host mirror, lazily), forward norm_b, L-SR1 as
  int ins0 = 0;
  // optional compact-representation inverse apply: Gram matrices by ring slot, device middle matrix
  bool inv_compact = false, w_dirty = true;
  bool push_streamed = true;    // push!: rebuild the a_k with the streaming apply kernel (false: generic multi-dot + lincomb passes)
  bool fwd_compact = false;     // forward operator in compact form: state = S, Y + Gram matrices; push! is O(m) dots, not O(m^2) passes
  std::vector<double> SY, YY, SS;   // [mem*mem]: SY[i*mem+j] = s_i·y_j, YY[i*mem+j] = y_i·y_j, SS[i*mem+j] = s_i·s_j
  double *d_W = nullptr;        // [(2*mem)^2]
  double *h_W = nullptr;        // pinned staging
  double *col(double *base, int k0) const { return base + (size_t)k0 * (size_t)pitch; }   // Float64 handles only
  void *colv(double *base, int k0) const { return reinterpret_cast<char *>(base) + (size_t)k0 * (size_t)pitch * (size_t)esize; }
};

// Float32 operators (b2o_qn_f32.inc, included at the end of this file)
static int qn32_apply(b2o_qn *q, float *res, const float *x, double alpha, double beta);
static int qn32_push(b2o_qn *q, const float *s, const float *y, int *accepted);
static int qn32_push_damped(b2o_qn *q, const float *s, float *y, bool inverse_form, double alpha, const float *g, float *Bs, int *accepted);
static int qn32_diag(b2o_qn *q, float *d);
static int qn32_apply_multi(b2o_qn *q, float *res, int64_t ldr, const float *x, int64_t ldx, int nrhs, double alpha, double beta);
#define B2O_F64_ONLY(q, what)                                                                                                    \
  do {                                                                                                                           \
    if ((q)->esize != 8) B2O_FAIL(B2O_EUNSUPPORTED, what " is not built for Float32 quasi-Newton operators (Float64 only)");      \
  } while (0)

static inline int pmod(int a, int m) {
  int r = a % m;
  return r < 0 ? r + m : r;
}

// ------------------------------------------------------------------ small elementwise helpers (push!/diag!, not the hot path)
// One shape for the three elementwise helpers of push!/diag!: 16-byte vector loads, four independent pairs in flight per thread
// (loads first, then the arithmetic and the streaming stores), scalar tail / scalar path for 8-byte-aligned user views.
// Round 1 ran these as scalar grid-stride loops at 3.6-3.8 TB/s (profiles/r2_launches_fwd.csv: div_sqrt_dev_kernel 0.42 ms at n=1e8).
// F: out[i] = f(x[i], y[i]) with the reference's rounding (the library is built with -fmad=false).
template <typename F>
__device__ __forceinline__ void ew_stream(double *__restrict__ out, const double *__restrict__ x, const double *__restrict__ y, int64_t n, bool vec,
                                          F f) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t npair = n >> 1;
    for (int64_t base = tid; base < npair; base += 4 * nthr) {
      double2 xv[4], yv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = base + u * nthr;
        xv[u] = yv[u] = make_double2(0.0, 0.0);
        if (i < npair) {
          xv[u] = ldg_stream2(x + 2 * i);
          if (y) yv[u] = ldg_stream2(y + 2 * i);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = base + u * nthr;
        if (i < npair) stg_stream2(out + 2 * i, make_double2(f(xv[u].x, yv[u].x), f(xv[u].y, yv[u].y)));
      }
    }
    if ((n & 1) && tid == 0) out[n - 1] = f(x[n - 1], y ? y[n - 1] : 0.0);
  } else {
    for (int64_t i = tid; i < n; i += nthr) out[i] = f(x[i], y ? y[i] : 0.0);
  }
}
static inline bool ew_vec_ok(const void *a, const void *b, const void *c) {
  return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) % 16) == 0;
}
// out = a*x + b*y   (y may be null -> out = a*x);  rounding as the reference's broadcasts (no fma)
__global__ void __launch_bounds__(256) axpby_kernel(double *out, double a, const double *x, double b, const double *y, int64_t n, int vec) {
  if (y) ew_stream(out, x, y, n, vec != 0, [a, b](double xi, double yi) { return a * xi + b * yi; });
  else ew_stream(out, x, y, n, vec != 0, [a](double xi, double) { return a * xi; });
}
// out = x / d
__global__ void __launch_bounds__(256) div_kernel(double *out, const double *x, double d, int64_t n, int vec) {
  ew_stream(out, x, nullptr, n, vec != 0, [d](double xi, double) { return xi / d; });
}
// out = x / sqrt(*dscal)   (device scalar)
__global__ void __launch_bounds__(256) div_sqrt_dev_kernel(double *out, const double *x, const double *dscal, int64_t n, int vec) {
  const double d = sqrt(*dscal);
  ew_stream(out, x, nullptr, n, vec != 0, [d](double xi, double) { return xi / d; });
}

// *out = partials[0] + ... + partials[count-1] in index order (one warp)
// out[o] = partials[o] + partials[2+o] + ... (count pairs, index order), o = warp index (launch with 32*nout threads)
__global__ void sum_partials_kernel(const double *partials, int count, double *out) {
  const int o = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double s = 0.0;
  for (int b = lane; b < count; b += 32) s += partials[2 * b + o];
  s = warp_sum(s);
  if (lane == 0) out[o] = s;
}

// out[j] = base(j), then sequentially out[j] (+|-)= coef_t * col_t[j];  coef_t = dots[t] / cdiv[t].
// base: mode 0: P1[j]/g ; mode 1: P0[j] - P1[j]/g ; mode 2: g*P1[j].   coef_t = dots[t]/cdiv[t] (or dots[t]*cdiv[t] if cmul).
// Fused reductions: red[0] = out·u0 ; red[1] = out·out (or out·u1 when u1 is given).
struct LincombArgs {
  const double *cols[B2O_MAX_COLS];
  double cdiv[B2O_MAX_COLS];
  signed char sign[B2O_MAX_COLS];
  int nterms;
  int base_mode;
  const double *P0, *P1;
  double g;
  const double *dots;  // device coefficients
  const double *u0;    // may be null
  const double *u1;    // may be null: red[1] = out·u1 instead of out·out
  int cmul;            // coef_t = dots[t] * cdiv[t]
  double *out;
  int64_t n;
  double *partials;    // [grid][2]
  double *red;         // [2]
  unsigned long long *arrive;
};
__global__ void __launch_bounds__(256) lincomb_kernel(const __grid_constant__ LincombArgs a) {
  __shared__ double scoef[B2O_MAX_COLS];
  __shared__ double sred[2][8];
  __shared__ bool is_last;
  for (int t = threadIdx.x; t < a.nterms; t += blockDim.x) scoef[t] = a.cmul ? a.cdiv[t] * a.dots[t] : a.dots[t] / a.cdiv[t];
  __syncthreads();
  double r0 = 0.0, r1 = 0.0;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    double v = a.base_mode == 0 ? a.P1[i] / a.g : a.base_mode == 1 ? a.P0[i] - a.P1[i] / a.g : a.g * a.P1[i];
    for (int t = 0; t < a.nterms; ++t) {
      double c = scoef[t] * a.cols[t][i];
      v = a.sign[t] > 0 ? v + c : v - c;
    }
    a.out[i] = v;
    if (a.u0) r0 = fma(v, a.u0[i], r0);
    r1 = fma(v, a.u1 ? a.u1[i] : v, r1);
  }
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  r0 = warp_sum(r0);
  r1 = warp_sum(r1);
  if (lane == 0) {
    sred[0][warp] = r0;
    sred[1][warp] = r1;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sred[threadIdx.x][w];
    a.partials[(size_t)blockIdx.x * 2 + threadIdx.x] = s;
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
    if (warp < 2) {
      double s = 0.0;
      for (int b = lane; b < gridDim.x; b += 32) s += __ldcg(&a.partials[(size_t)b * 2 + warp]);
      s = warp_sum(s);
      if (lane == 0) a.red[warp] = s;
    }
    if (threadIdx.x == 0) *a.arrive = 0ULL;
  }
}

// diag!: d = 1 (/γ) ; d += b_k^2 - a_k^2  (L-BFGS :379-395)  |  d += a_k^2/as_k  (L-SR1 :196-211)
template <typename T>
struct DiagArgsT {
  const T *c0[B2O_MAX_MEM];
  const T *c1[B2O_MAX_MEM];
  double cdiv[B2O_MAX_MEM];
  int nact, kind, scaling;
  double gamma;
  T *d;
  int64_t n;
};
using DiagArgs = DiagArgsT<double>;
template <typename T>
__global__ void qn_diag_kernel(const __grid_constant__ DiagArgsT<T> a) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const T gamma = (T)a.gamma;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    T d = (T)1;
    if (a.scaling) d = d / gamma;
    for (int k = 0; k < a.nact; ++k) {
      if (a.kind == 0) {
        T av = a.c0[k][i], bv = a.c1[k][i];
        d = d + (bv * bv - av * av);
      } else {
        T av = a.c0[k][i];
        d = d + (av * av) / (T)a.cdiv[k];
      }
    }
    a.d[i] = d;
  }
}

static inline int ew_grid(b2o_ctx *c, int64_t n) {
  int64_t want = (n + 255) / 256;
  int64_t cap = (int64_t)c->num_sms * 8;
  return (int)std::max<int64_t>(1, std::min(want, cap));
}
static int ew_axpby(b2o_ctx *c, double *out, double a, const double *x, double b, const double *y, int64_t n) {
  if (n <= 0) return B2O_OK;
  axpby_kernel<<<ew_grid(c, n), 256, 0, c->stream>>>(out, a, x, b, y, n, ew_vec_ok(out, x, y));
  c->launches++;
  B2O_CUDA(cudaGetLastError());
  return B2O_OK;
}

// ------------------------------------------------------------------ create / destroy
static int qn_alloc(b2o_qn *q) {
  b2o_ctx *c = q->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  const size_t colb = (size_t)q->pitch * (size_t)q->esize;
  auto alloc0 = [&](double **p, size_t bytes) -> int {
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      B2O_FAIL(B2O_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    }
    B2O_CUDA(cudaMemsetAsync(*p, 0, bytes, c->stream));
    return B2O_OK;
  };
  B2O_TRY(alloc0(&q->S, colb * q->mem));
  B2O_TRY(alloc0(&q->Y, colb * q->mem));
  if (q->kind == 1 || !q->inverse) B2O_TRY(alloc0(&q->A, colb * q->mem));
  if (q->kind == 0 && !q->inverse) B2O_TRY(alloc0(&q->B, colb * q->mem));
  B2O_TRY(alloc0(&q->q, colb));
  B2O_TRY(alloc0(&q->tmp, colb));
  B2O_TRY(alloc0(&q->d_alpha, sizeof(double) * B2O_MAX_MEM));
  q->ys.assign(q->mem, 0.0);
  q->aux.assign(q->mem, 0.0);
  return B2O_OK;
}

extern "C" int b2o_qn_destroy(b2o_qn *q) {
  if (!q) return B2O_OK;
  cudaSetDevice(q->ctx->device);
  cudaStreamSynchronize(q->ctx->stream);
  cudaFree(q->S);
  cudaFree(q->Y);
  cudaFree(q->A);
  cudaFree(q->B);
  cudaFree(q->q);
  cudaFree(q->tmp);
  cudaFree(q->d_alpha);
  if (q->shifted_p) cudaFree(q->shifted_p);
  if (q->q_multi) cudaFree(q->q_multi);
  if (q->d_W) cudaFree(q->d_W);
  if (q->h_W) cudaFreeHost(q->h_W);
  delete q;
  return B2O_OK;
}

static int qn_create_common(b2o_ctx *ctx, int dtype, int64_t n, int mem, b2o_qn **out, b2o_qn **made) {
  if (!ctx || !out) B2O_FAIL(B2O_EARG, "null argument");
  if (dtype != B2O_F64 && dtype != B2O_F32) B2O_FAIL(B2O_EUNSUPPORTED, "quasi-Newton operators: dtype %d not supported (Float64 or Float32)", dtype);
  if (n < 0) B2O_FAIL(B2O_EARG, "n must be >= 0");
  if (mem < 1) mem = 1;  // LBFGSData clamps mem to max(mem,1) (src/lbfgs.jl:37)
  if (mem > B2O_MAX_MEM) B2O_FAIL(B2O_EUNSUPPORTED, "mem=%d exceeds the built maximum %d", mem, B2O_MAX_MEM);
  b2o_qn *q = new b2o_qn_s();
  q->ctx = ctx;
  q->n = n;
  q->pitch = std::max<int64_t>(B2O_PITCH_ALIGN, (n + B2O_PITCH_ALIGN - 1) / B2O_PITCH_ALIGN * B2O_PITCH_ALIGN);
  q->mem = mem;
  q->esize = dtype == B2O_F32 ? 4 : 8;
  *made = q;
  return B2O_OK;
}

extern "C" int b2o_lbfgs_create(b2o_ctx *ctx, int dtype, int64_t n, int mem, int scaling, int damped, double sigma2,
                                double sigma3, int inverse, b2o_qn **out) {
  b2o_qn *q = nullptr;
  B2O_TRY(qn_create_common(ctx, dtype, n, mem, out, &q));
  q->kind = 0;
  q->scaling = scaling != 0;
  q->damped = damped != 0;
  q->inverse = inverse != 0;
  q->sigma2 = sigma2;
  q->sigma3 = sigma3;
  int st = qn_alloc(q);
  if (st != B2O_OK) {
    b2o_qn_destroy(q);
    return st;
  }
  *out = q;
  return B2O_OK;
}

extern "C" int b2o_lsr1_create(b2o_ctx *ctx, int dtype, int64_t n, int mem, int scaling, b2o_qn **out) {
  b2o_qn *q = nullptr;
  B2O_TRY(qn_create_common(ctx, dtype, n, mem, out, &q));
  q->kind = 1;
  q->scaling = scaling != 0;
  int st = qn_alloc(q);
  if (st != B2O_OK) {
    b2o_qn_destroy(q);
    return st;
  }
  *out = q;
  return B2O_OK;
}

// ------------------------------------------------------------------ launch plumbing
struct LaunchCfg {
  int R, stages, grid, group;
  SmemLayout L;
};
static int plan_launch(b2o_ctx *c, int64_t ntiles, int ncols, bool need_accs, LaunchCfg *cfg) {
  cfg->R = c->tile_rows;
  cfg->group = need_accs ? std::max(1, std::min(ncols, 40)) : 0;
  const size_t max_smem = B2O_MAX_DYN_SMEM;
  int stages = 32;
  for (;; --stages) {
    cfg->L = smem_layout(cfg->R, stages, cfg->group);
    if (cfg->L.total <= max_smem) break;
    if (stages <= 2) B2O_FAIL(B2O_ECUDA, "shared memory plan does not fit");
  }
  // measured on B200 (tools/sweep_stream.py, profiles/r1_sweep.jsonl): ~112-128 KB of column data in flight per SM is the
  // sweet spot (2048x7, 4096x4, 1024x16); deeper rings lose 3-4%.
  const int auto_stages = std::max(3, (int)((114688 + cfg->R * 4) / (cfg->R * 8)));
  stages = std::min(stages, c->stages > 0 ? std::max(2, c->stages) : auto_stages);
  cfg->stages = stages;
  cfg->L = smem_layout(cfg->R, stages, cfg->group);
  int g = c->grid > 0 ? c->grid : c->num_sms;
  g = std::min(g, c->num_sms);  // co-residency: 1 CTA per SM
  cfg->grid = (int)std::max<int64_t>(1, std::min<int64_t>(g, ntiles));
  return B2O_OK;
}

template <typename K, typename Args>
static int launch_persistent(b2o_ctx *c, K kern, const LaunchCfg &cfg, Args &args, bool cooperative) {
  static thread_local const void *configured[64];
  static thread_local int nconfigured = 0;
  bool seen = false;
  for (int i = 0; i < nconfigured; ++i) seen |= (configured[i] == (const void *)kern);
  if (!seen) {
    B2O_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)B2O_MAX_DYN_SMEM));
    if (nconfigured < 64) configured[nconfigured++] = (const void *)kern;
  }
  if (c->time_kernels) B2O_CUDA(cudaEventRecord(c->ev0, c->stream));
  void *kargs[] = {(void *)&args};
  if (cooperative) {
    B2O_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3(cfg.grid), dim3(B2O_NTHREADS), kargs, cfg.L.total,
                                         c->stream));
  } else {
    B2O_CUDA(cudaLaunchKernel((const void *)kern, dim3(cfg.grid), dim3(B2O_NTHREADS), kargs, cfg.L.total, c->stream));
  }
  c->launches++;
  if (c->time_kernels) {
    B2O_CUDA(cudaEventRecord(c->ev1, c->stream));
    B2O_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    B2O_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->kern_ms += ms;
    c->kern_n++;
  }
  return B2O_OK;
}

template <int OP>
static int launch_compact_R(b2o_ctx *c, const LaunchCfg &cfg, CompactArgs &a, bool coop) {
  switch (cfg.R) {
    case 1024: return launch_persistent(c, qn_compact_kernel<1024, OP>, cfg, a, coop);
    case 2048: return launch_persistent(c, qn_compact_kernel<2048, OP>, cfg, a, coop);
    case 4096: return launch_persistent(c, qn_compact_kernel<4096, OP>, cfg, a, coop);
  }
  B2O_FAIL(B2O_EARG, "bad tile_rows");
}
static int launch_twoloop_R(b2o_ctx *c, const LaunchCfg &cfg, TwoLoopArgs &a, bool coop) {
  switch (cfg.R) {
    case 1024: return launch_persistent(c, qn_twoloop_kernel<1024>, cfg, a, coop);
    case 2048: return launch_persistent(c, qn_twoloop_kernel<2048>, cfg, a, coop);
    case 4096: return launch_persistent(c, qn_twoloop_kernel<4096>, cfg, a, coop);
  }
  B2O_FAIL(B2O_EARG, "bad tile_rows");
}

// ordered list of active ring slots, oldest -> newest: k = mod(insert+i-2, mem)+1, ys[k] != 0
// (src/lbfgs.jl:189-191, :141-143; src/lsr1.jl:98-100)
static int active_old_to_new(const b2o_qn *q, int *slots) {
  int na = 0;
  for (int i = 1; i <= q->mem; ++i) {
    int k = pmod(q->ins0 + i - 1, q->mem);
    if (q->ys[k] != 0) slots[na++] = k;
  }
  return na;
}

// ------------------------------------------------------------------ apply
// fills the column table (reference order) and scalars; row range / launch geometry are set by the launcher
static void compact_columns(const b2o_qn *q, CompactArgs &a, double alpha, double beta) {
  int slots[B2O_MAX_MEM];
  const int na = active_old_to_new(q, slots);
  memset(&a, 0, sizeof(a));
  if (q->kind == 0 && (q->inverse || q->fwd_compact)) {
    // compact forms: columns [s_old..s_new, y_old..y_new]
    for (int i = 0; i < na; ++i) {
      a.cols[i] = q->col(q->S, slots[i]);
      a.cols[na + i] = q->col(q->Y, slots[i]);
      a.cdiv[i] = a.cdiv[na + i] = 1.0;
    }
    a.ncols = 2 * na;
    a.W = q->d_W;
    a.base_div = q->inverse ? 0 : 1;
  } else if (q->kind == 0) {
    for (int i = 0; i < na; ++i) {
      a.cols[2 * i] = q->col(q->A, slots[i]);
      a.cols[2 * i + 1] = q->col(q->B, slots[i]);
      a.cdiv[2 * i] = a.cdiv[2 * i + 1] = 1.0;
    }
    a.ncols = 2 * na;
  } else {
    for (int i = 0; i < na; ++i) {
      a.cols[i] = q->col(q->A, slots[i]);
      a.cdiv[i] = q->aux[slots[i]];
    }
    a.ncols = na;
  }
  a.alpha = alpha;
  a.beta = beta;
  a.gamma = q->gamma;
  a.scaling = q->scaling ? 1 : 0;
}

// one launch of the compact kernel over rows [r0, r1) (r0 a multiple of the pitch alignment)
static int compact_launch_rows(b2o_qn *q, const CompactArgs &base, double *res, const double *x, int64_t r0, int64_t r1, int mode,
                               int accumulate, int push_op = -1, int *grid_out = nullptr) {
  b2o_ctx *c = q->ctx;
  CompactArgs a = base;
  for (int i = 0; i < a.ncols; ++i) a.cols[i] += r0;
  a.x = x + r0;
  a.res = res + r0;
  if (a.y2) a.y2 += r0;
  a.n = r1 - r0;
  LaunchCfg cfg;
  cfg.R = c->tile_rows;
  a.ntiles = (a.n + cfg.R - 1) / cfg.R;
  B2O_TRY(plan_launch(c, a.ntiles, a.ncols, true, &cfg));
  a.x_al16 = ((uintptr_t)a.x % 16) == 0;
  a.res_al16 = ((uintptr_t)a.res % 16) == 0;
  a.partials = c->d_partials;
  a.dots = c->d_dots;
  a.bar = c->d_bar;
  a.arrive = c->d_bar + 1;
  a.stages = cfg.stages;
  a.group = std::max(1, cfg.group);
  a.accs_off = (uint32_t)cfg.L.accs_off;
  a.coef_off = (uint32_t)cfg.L.coef_off;
  a.bar_off = (uint32_t)cfg.L.bar_off;
  a.mode = mode;
  a.accumulate = accumulate;
  if (a.n <= 0) return B2O_OK;
  const bool coop = mode == MODE_FUSED;
  if (coop) a.bar_target = c->bar_base + (unsigned long long)cfg.grid;
  b2o_mbox_fill(c, &a.mbox);
  if (!coop) a.mbox.nranks = 1;
  a.dbg = c->d_dots + B2O_WS_QNDBG;
  if (grid_out) *grid_out = cfg.grid;
  int st = push_op == OP_PUSH_A                                     ? launch_compact_R<OP_PUSH_A>(c, cfg, a, coop)
           : push_op == OP_PUSH_L                                   ? launch_compact_R<OP_PUSH_L>(c, cfg, a, coop)
           : (q->kind == 0 && (q->inverse || q->fwd_compact)) ? launch_compact_R<OP_INV_COMPACT>(c, cfg, a, coop)
           : q->kind == 0             ? launch_compact_R<OP_LBFGS_FWD>(c, cfg, a, coop)
                                      : launch_compact_R<OP_LSR1>(c, cfg, a, coop);
  if (st == B2O_OK && coop) {
    c->bar_base += (unsigned long long)cfg.grid;
    if (a.mbox.nranks > 1) c->mbox_epoch += 1;
  }
  return st;
}

static int qn_apply_compact(b2o_qn *q, double *res, const double *x, double alpha, double beta) {
  b2o_ctx *c = q->ctx;
  if (q->n == 0) return B2O_OK;
  CompactArgs a;
  compact_columns(q, a, alpha, beta);
  if (a.ncols == 0) return compact_launch_rows(q, a, res, x, 0, q->n, MODE_PHASE2, 0);
  if (c->nranks <= 1 || c->mbox_ready) return compact_launch_rows(q, a, res, x, 0, q->n, MODE_FUSED, 0);   // mailbox: still ONE launch
  // row-partitioned without the mailbox: local dots -> one NCCL all-reduce of ncols scalars -> combine
  B2O_TRY(compact_launch_rows(q, a, res, x, 0, q->n, MODE_PHASE1, 0));
  B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots, a.ncols));
  return compact_launch_rows(q, a, res, x, 0, q->n, MODE_PHASE2, 0);
}

// A == 0 (fresh / reset operator): q = x; scaling && q *= γ; res = α q (+ β res)
__global__ void twoloop_empty_kernel(double *res, const double *x, double alpha, double beta, double gamma, int scaling,
                                     int64_t n) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    double q = x[i];
    if (scaling) q = q * gamma;
    res[i] = (beta != 0.0) ? alpha * q + beta * res[i] : alpha * q;
  }
}

static int qn_apply_twoloop(b2o_qn *q, double *res, const double *x, double alpha, double beta) {
  b2o_ctx *c = q->ctx;
  int o2n[B2O_MAX_MEM];
  const int na = active_old_to_new(q, o2n);
  if (q->n == 0) return B2O_OK;
  if (na == 0) {
    twoloop_empty_kernel<<<ew_grid(c, q->n), 256, 0, c->stream>>>(res, x, alpha, beta, q->gamma, q->scaling ? 1 : 0, q->n);
    c->launches++;
    B2O_CUDA(cudaGetLastError());
    return B2O_OK;
  }
  TwoLoopArgs a;
  memset(&a, 0, sizeof(a));
  // loop 1 runs newest -> oldest: k = mod(insert-i-1, mem)+1 (src/lbfgs.jl:130-131) = reverse of o2n
  for (int i = 0; i < na; ++i) {
    int k = o2n[na - 1 - i];
    a.s[i] = q->col(q->S, k);
    a.y[i] = q->col(q->Y, k);
    a.ys[i] = q->ys[k];
  }
  a.nact = na;
  a.x = x;
  a.res = res;
  a.q = q->q;
  a.alpha_out = q->d_alpha;
  a.n = q->n;
  LaunchCfg cfg;
  cfg.R = c->tile_rows;
  a.ntiles = (q->n + cfg.R - 1) / cfg.R;
  B2O_TRY(plan_launch(c, a.ntiles, 0, false, &cfg));
  a.alpha = alpha;
  a.beta = beta;
  a.gamma = q->gamma;
  a.scaling = q->scaling ? 1 : 0;
  a.x_al16 = ((uintptr_t)x % 16) == 0;
  a.res_al16 = ((uintptr_t)res % 16) == 0;
  a.partials = c->d_partials;
  a.dots = c->d_dots;
  a.bar = c->d_bar;
  a.arrive = c->d_bar + 1;
  a.stages = cfg.stages;
  a.coef_off = (uint32_t)cfg.L.coef_off;
  a.bar_off = (uint32_t)cfg.L.bar_off;
  const int nsweeps = 2 * na + 1;
  b2o_mbox_fill(c, &a.mbox);
  a.sweep_dots = c->d_dots + B2O_WS_SWEEP;
  a.dbg = c->d_dots + B2O_WS_QNDBG;
  if (c->nranks <= 1 || c->mbox_ready) {
    // single GPU, or row-partitioned with the NVLink mailbox: all 2A+1 sweeps and all 2A all-reduces in ONE launch
    a.sweep_begin = 0;
    a.sweep_end = nsweeps;
    a.bar_target = c->bar_base + (unsigned long long)cfg.grid;
    B2O_TRY(launch_twoloop_R(c, cfg, a, true));
    c->bar_base += (unsigned long long)cfg.grid * (unsigned long long)(2 * na);
    if (a.mbox.nranks > 1) c->mbox_epoch += (unsigned long long)(2 * na);
  } else {
    a.mbox.nranks = 1;
    // row-partitioned: one launch + one NCCL all-reduce per inner product (north_star / SURVEY §8e)
    for (int w = 0; w < nsweeps; ++w) {
      a.sweep_begin = w;
      a.sweep_end = w + 1;
      B2O_TRY(launch_twoloop_R(c, cfg, a, false));
      if (w < nsweeps - 1) B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots, 1));
    }
  }
  return B2O_OK;
}

// W' = diag(I, γI) · [[R^{-T}(D + γ YᵀY) R^{-1}, -R^{-T}], [-R^{-1}, 0]] · diag(I, γI) for the active pairs, oldest -> newest
// (Byrd, Nocedal, Schnabel 1994, eq. 4.x: H = γI + [S γY] W [Sᵀ; γYᵀ]); small (2A x 2A), computed on the host in long double.
static int build_inverse_W(b2o_qn *q) {
  b2o_ctx *c = q->ctx;
  int sl[B2O_MAX_MEM];
  const int A = active_old_to_new(q, sl), m = q->mem, N = 2 * A;
  if (A == 0) {
    q->w_dirty = false;
    return B2O_OK;
  }
  const long double g = q->scaling ? (long double)q->gamma : 1.0L;
  std::vector<long double> R((size_t)A * A, 0.0L), Ri((size_t)A * A, 0.0L), M((size_t)A * A, 0.0L), T((size_t)A * A, 0.0L);
  for (int i = 0; i < A; ++i)
    for (int j = i; j < A; ++j) R[(size_t)i * A + j] = q->SY[(size_t)sl[i] * m + sl[j]];
  for (int j = 0; j < A; ++j) {       // Ri = R^{-1} (upper triangular), column by column
    for (int i = A - 1; i >= 0; --i) {
      long double s = (i == j) ? 1.0L : 0.0L;
      for (int k = i + 1; k < A; ++k) s -= R[(size_t)i * A + k] * Ri[(size_t)k * A + j];
      Ri[(size_t)i * A + j] = s / R[(size_t)i * A + i];
    }
  }
  for (int i = 0; i < A; ++i)
    for (int j = 0; j < A; ++j)
      M[(size_t)i * A + j] = g * (long double)q->YY[(size_t)sl[i] * m + sl[j]] + (i == j ? (long double)q->SY[(size_t)sl[i] * m + sl[i]] : 0.0L);
  for (int i = 0; i < A; ++i)         // T = M * Ri
    for (int j = 0; j < A; ++j) {
      long double s = 0.0L;
      for (int k = 0; k < A; ++k) s += M[(size_t)i * A + k] * Ri[(size_t)k * A + j];
      T[(size_t)i * A + j] = s;
    }
  double *W = q->h_W;
  for (int i = 0; i < A; ++i)
    for (int j = 0; j < A; ++j) {
      long double s = 0.0L;           // (Ri^T * T)[i][j]
      for (int k = 0; k < A; ++k) s += Ri[(size_t)k * A + i] * T[(size_t)k * A + j];
      W[(size_t)i * N + j] = (double)s;                                    // S-S block
      W[(size_t)i * N + A + j] = (double)(-g * Ri[(size_t)j * A + i]);     // S-Y block: -R^{-T} (times γ on the Y side)
      W[(size_t)(A + i) * N + j] = (double)(-g * Ri[(size_t)i * A + j]);   // Y-S block: -R^{-1}
      W[(size_t)(A + i) * N + A + j] = 0.0;
    }
  B2O_CUDA(cudaMemcpyAsync(q->d_W, W, sizeof(double) * N * N, cudaMemcpyHostToDevice, c->stream));
  B2O_CUDA(cudaStreamSynchronize(c->stream));   // h_W is reused by the next rebuild
  q->w_dirty = false;
  return B2O_OK;
}

// in-place Gauss-Jordan inverse with partial pivoting (long double); false if singular
static bool invert_ld(std::vector<long double> &M, std::vector<long double> &Inv, int N) {
  Inv.assign((size_t)N * N, 0.0L);
  for (int i = 0; i < N; ++i) Inv[(size_t)i * N + i] = 1.0L;
  for (int col = 0; col < N; ++col) {
    int piv = col;
    for (int r = col + 1; r < N; ++r)
      if (fabsl(M[(size_t)r * N + col]) > fabsl(M[(size_t)piv * N + col])) piv = r;
    if (M[(size_t)piv * N + col] == 0.0L) return false;
    if (piv != col)
      for (int k = 0; k < N; ++k) {
        std::swap(M[(size_t)piv * N + k], M[(size_t)col * N + k]);
        std::swap(Inv[(size_t)piv * N + k], Inv[(size_t)col * N + k]);
      }
    const long double d = M[(size_t)col * N + col];
    for (int k = 0; k < N; ++k) {
      M[(size_t)col * N + k] /= d;
      Inv[(size_t)col * N + k] /= d;
    }
    for (int r = 0; r < N; ++r) {
      if (r == col) continue;
      const long double f = M[(size_t)r * N + col];
      if (f == 0.0L) continue;
      for (int k = 0; k < N; ++k) {
        M[(size_t)r * N + k] -= f * M[(size_t)col * N + k];
        Inv[(size_t)r * N + k] -= f * Inv[(size_t)col * N + k];
      }
    }
  }
  return true;
}

// middle matrix of the compact forward form for the active pairs (age order): M = [[SᵀS/γ, L], [Lᵀ, -D]]
static void forward_middle(const b2o_qn *q, const int *sl, int A, long double g, std::vector<long double> &M) {
  const int m = q->mem, N = 2 * A;
  M.assign((size_t)N * N, 0.0L);
  for (int i = 0; i < A; ++i)
    for (int j = 0; j < A; ++j) {
      M[(size_t)i * N + j] = (long double)q->SS[(size_t)sl[i] * m + sl[j]] / g;
      const long double Lij = (i > j) ? (long double)q->SY[(size_t)sl[i] * m + sl[j]] : 0.0L;
      M[(size_t)i * N + A + j] = Lij;
      M[(size_t)(A + j) * N + i] = Lij;
      M[(size_t)(A + i) * N + A + j] = (i == j) ? -(long double)q->SY[(size_t)sl[i] * m + sl[i]] : 0.0L;
    }
}

// Compact FORWARD form (Byrd, Nocedal, Schnabel 1994, Thm 2.3) with B0 = I/γ:
//   B = B0 - [B0 S  Y] [[SᵀB0S, L], [Lᵀ, -D]]^{-1} [SᵀB0; Yᵀ],   L_ij = s_i·y_j (i > j), D = diag(s_i·y_i)
// so  B x = x/γ + [S Y] W' [Sᵀx; Yᵀx]  with  W' = -diag(1/γ, 1) M^{-1} diag(1/γ, 1).  M is 2A x 2A, inverted on the host
// (Gauss-Jordan with partial pivoting, long double).
static int build_forward_W(b2o_qn *q) {
  b2o_ctx *c = q->ctx;
  int sl[B2O_MAX_MEM];
  const int A = active_old_to_new(q, sl), m = q->mem, N = 2 * A;
  if (A == 0) {
    q->w_dirty = false;
    return B2O_OK;
  }
  const long double g = q->scaling ? (long double)q->gamma : 1.0L;
  std::vector<long double> M, Inv;
  forward_middle(q, sl, A, g, M);
  if (!invert_ld(M, Inv, N)) B2O_FAIL(B2O_ESTATE, "compact forward L-BFGS: singular middle matrix");
  double *W = q->h_W;
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) {
      long double v = -Inv[(size_t)i * N + j];
      if (i < A) v /= g;
      if (j < A) v /= g;
      W[(size_t)i * N + j] = (double)v;
    }
  B2O_CUDA(cudaMemcpyAsync(q->d_W, W, sizeof(double) * N * N, cudaMemcpyHostToDevice, c->stream));
  B2O_CUDA(cudaStreamSynchronize(c->stream));
  q->w_dirty = false;
  return B2O_OK;
}

static int qn_apply_dev(b2o_qn *q, double *res, const double *x, double alpha, double beta) {
  if (q->kind == 0 && !q->inverse && q->fwd_compact && q->w_dirty) B2O_TRY(build_forward_W(q));
  if (q->kind == 0 && q->inverse) {
    if (!q->inv_compact) return qn_apply_twoloop(q, res, x, alpha, beta);
    if (q->w_dirty) B2O_TRY(build_inverse_W(q));
  }
  return qn_apply_compact(q, res, x, alpha, beta);
}

// ------------------------------------------------------------------ block apply: NR right-hand sides per column pass
template <int NR, typename T = double>
static int multi_launch(b2o_qn *q, const CompactArgsT<T> &base, T *res, int64_t ldr, const T *x, int64_t ldx, int nrhs) {
  b2o_ctx *c = q->ctx;
  constexpr int R = MultiTile<NR, T>::R;
  MultiArgsT<T> a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < base.ncols; ++i) {
    a.cols[i] = base.cols[i];
    a.cdiv[i] = base.cdiv[i];
  }
  a.ncols = base.ncols;
  a.x = x;
  a.res = res;
  a.ldx = ldx;
  a.ldr = ldr;
  a.nrhs = nrhs;
  a.n = q->n;
  a.ntiles = (a.n + R - 1) / R;
  a.alpha = base.alpha;
  a.beta = base.beta;
  a.gamma = base.gamma;
  a.scaling = base.scaling;
  a.W = base.W;
  a.base_div = base.base_div;
  constexpr int64_t VN = 16 / (int64_t)sizeof(T);   // elements per 16-byte vector: every column must start on one
  a.x_al16 = ((uintptr_t)x % 16) == 0 && (nrhs == 1 || ldx % VN == 0);
  a.res_al16 = ((uintptr_t)res % 16) == 0 && (nrhs == 1 || ldr % VN == 0);
  const int nv = a.ncols * NR;
  LaunchCfg cfg;
  cfg.R = R;
  cfg.group = 0;
  const size_t fixed = ((size_t)B2O_CONS_WARPS * a.ncols * 32 + nv) * sizeof(double) + B2O_NCONS * sizeof(unsigned);
  const int tile_bytes = R * (int)sizeof(T);
  int stages = c->stages > 0 ? std::max(2, c->stages) : std::max(3, (114688 + tile_bytes / 2) / tile_bytes);
  while (stages > 2 && (size_t)stages * tile_bytes + fixed + 2 * stages * sizeof(uint64_t) + 16 > B2O_MAX_DYN_SMEM) --stages;
  cfg.stages = stages;
  cfg.L.ring_off = 0;
  cfg.L.accs_off = (size_t)stages * tile_bytes;
  cfg.L.coef_off = cfg.L.accs_off + (size_t)B2O_CONS_WARPS * a.ncols * 32 * sizeof(double);
  cfg.L.bar_off = cfg.L.coef_off + (size_t)nv * sizeof(double);
  a.landed_off = (uint32_t)(cfg.L.bar_off + (size_t)2 * stages * sizeof(uint64_t));
  cfg.L.total = a.landed_off + B2O_NCONS * sizeof(unsigned);
  if (cfg.L.total > B2O_MAX_DYN_SMEM) B2O_FAIL(B2O_ECUDA, "shared memory plan does not fit");
  int g = c->grid > 0 ? c->grid : c->num_sms;
  g = std::min(g, c->num_sms);
  cfg.grid = (int)std::max<int64_t>(1, std::min<int64_t>(g, a.ntiles));
  if ((size_t)cfg.grid * nv > (size_t)B2O_MAX_GRID * B2O_MAX_COLS) B2O_FAIL(B2O_EUNSUPPORTED, "too many columns for the block apply");
  a.partials = c->d_partials;
  a.dots = c->d_dots;
  a.bar = c->d_bar;
  a.stages = stages;
  a.wacc_off = (uint32_t)cfg.L.accs_off;
  a.coef_off = (uint32_t)cfg.L.coef_off;
  a.bar_off = (uint32_t)cfg.L.bar_off;
  a.bar_target = c->bar_base + (unsigned long long)cfg.grid;
  b2o_mbox_fill(c, &a.mbox);
  int st;
  if constexpr (sizeof(T) == 8) {
    st = (q->kind == 0 && (q->inverse || q->fwd_compact)) ? launch_persistent(c, qn_multi_kernel<NR, OP_INV_COMPACT, T>, cfg, a, true)
         : q->kind == 0                                   ? launch_persistent(c, qn_multi_kernel<NR, OP_LBFGS_FWD, T>, cfg, a, true)
                                                          : launch_persistent(c, qn_multi_kernel<NR, OP_LSR1, T>, cfg, a, true);
  } else {   // Float32 handles have no compact forms
    st = q->kind == 0 ? launch_persistent(c, qn_multi_kernel<NR, OP_LBFGS_FWD, T>, cfg, a, true)
                      : launch_persistent(c, qn_multi_kernel<NR, OP_LSR1, T>, cfg, a, true);
  }
  if (st == B2O_OK) {
    c->bar_base += (unsigned long long)cfg.grid;
    if (a.mbox.nranks > 1) c->mbox_epoch += (unsigned long long)((nv + MBOX_MAXV - 1) / MBOX_MAXV);
  }
  return st;
}

// block apply for 5..8 right-hand sides on the FP64 tensor cores (b2o_qn_multi_mma.cuh)
static int multi_mma_launch(b2o_qn *q, const CompactArgs &base, double *res, int64_t ldr, const double *x, int64_t ldx, int nrhs) {
  b2o_ctx *c = q->ctx;
  constexpr int NR = 8;
  MultiArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < base.ncols; ++i) {
    a.cols[i] = base.cols[i];
    a.cdiv[i] = base.cdiv[i];
  }
  a.ncols = base.ncols;
  a.x = x;
  a.res = res;
  a.ldx = ldx;
  a.ldr = ldr;
  a.nrhs = nrhs;
  a.n = q->n;
  a.ntiles = (a.n + MM_R - 1) / MM_R;
  a.alpha = base.alpha;
  a.beta = base.beta;
  a.gamma = base.gamma;
  a.scaling = base.scaling;
  a.W = base.W;
  a.base_div = base.base_div;
  const int nv = a.ncols * NR;
  LaunchCfg cfg;
  cfg.R = MM_R;
  cfg.group = 0;
  cfg.stages = MM_STAGES;
  cfg.L.ring_off = 0;
  cfg.L.accs_off = (size_t)MM_STAGES * MM_SLOT;                                            // wacc [8 warps][ncols][8]
  cfg.L.coef_off = cfg.L.accs_off + (size_t)B2O_CONS_WARPS * nv * sizeof(double);
  cfg.L.bar_off = cfg.L.coef_off + (size_t)nv * sizeof(double);
  a.landed_off = (uint32_t)(cfg.L.bar_off + (size_t)2 * MM_STAGES * sizeof(uint64_t));
  cfg.L.total = a.landed_off + B2O_NCONS * sizeof(unsigned);
  if (cfg.L.total > B2O_MAX_DYN_SMEM) B2O_FAIL(B2O_ECUDA, "shared memory plan does not fit");
  int g = c->grid > 0 ? c->grid : c->num_sms;
  g = std::min(g, c->num_sms);
  cfg.grid = (int)std::max<int64_t>(1, std::min<int64_t>(g, a.ntiles));
  if ((size_t)cfg.grid * nv > (size_t)B2O_MAX_GRID * B2O_MAX_COLS) B2O_FAIL(B2O_EUNSUPPORTED, "too many columns for the block apply");
  a.partials = c->d_partials;
  a.dots = c->d_dots;
  a.bar = c->d_bar;
  a.stages = MM_STAGES;
  a.wacc_off = (uint32_t)cfg.L.accs_off;
  a.coef_off = (uint32_t)cfg.L.coef_off;
  a.bar_off = (uint32_t)cfg.L.bar_off;
  a.bar_target = c->bar_base + (unsigned long long)cfg.grid;
  b2o_mbox_fill(c, &a.mbox);
  int st = (q->kind == 0 && (q->inverse || q->fwd_compact)) ? launch_persistent(c, qn_multi_mma_kernel<OP_INV_COMPACT>, cfg, a, true)
           : q->kind == 0             ? launch_persistent(c, qn_multi_mma_kernel<OP_LBFGS_FWD>, cfg, a, true)
                                      : launch_persistent(c, qn_multi_mma_kernel<OP_LSR1>, cfg, a, true);
  if (st == B2O_OK) {
    c->bar_base += (unsigned long long)cfg.grid;
    if (a.mbox.nranks > 1) c->mbox_epoch += (unsigned long long)((nv + MBOX_MAXV - 1) / MBOX_MAXV);
  }
  return st;
}

// ------------------------------------------------------------------ block two-loop recursion (matrix right-hand sides of H)
template <int NR>
static int twoloop_multi_launch(b2o_qn *q, double *res, int64_t ldr, const double *x, int64_t ldx, int nrhs, double alpha, double beta) {
  b2o_ctx *c = q->ctx;
  int o2n[B2O_MAX_MEM];
  const int na = active_old_to_new(q, o2n);
  if (q->q_multi_cols < NR) {     // work vectors q_r, allocated on the first block apply and kept
    if (q->q_multi) B2O_CUDA(cudaFree(q->q_multi));
    q->q_multi = nullptr;
    q->q_multi_cols = 0;
    cudaError_t e = cudaMalloc(&q->q_multi, sizeof(double) * (size_t)NR * (size_t)q->pitch);
    if (e != cudaSuccess) {
      cudaGetLastError();
      B2O_FAIL(B2O_ENOMEM, "block two-loop: cudaMalloc of %d work vectors failed: %s", NR, cudaGetErrorString(e));
    }
    B2O_CUDA(cudaMemsetAsync(q->q_multi, 0, sizeof(double) * (size_t)NR * (size_t)q->pitch, c->stream));
    q->q_multi_cols = NR;
  }
  TwoLoopMultiArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < na; ++i) {   // newest -> oldest (src/lbfgs.jl:130-131)
    const int k = o2n[na - 1 - i];
    a.s[i] = q->col(q->S, k);
    a.y[i] = q->col(q->Y, k);
    a.ys[i] = q->ys[k];
  }
  a.nact = na;
  a.x = x;
  a.res = res;
  a.ldx = ldx;
  a.ldr = ldr;
  a.nrhs = nrhs;
  a.q = q->q_multi;
  a.qpitch = q->pitch;
  a.n = q->n;
  LaunchCfg cfg;
  cfg.R = c->tile_rows;
  a.ntiles = (q->n + cfg.R - 1) / cfg.R;
  B2O_TRY(plan_launch(c, a.ntiles, 0, false, &cfg));
  // ring | alphas [B2O_MAX_MEM][NR], warp partials [8][NR], dots [NR] | barriers
  const size_t scal = (size_t)(B2O_MAX_MEM + B2O_CONS_WARPS + 1) * NR * sizeof(double);
  int stages = cfg.stages;
  const size_t landed = B2O_NCONS * sizeof(unsigned);
  while (stages > 2 && (size_t)stages * cfg.R * sizeof(double) + scal + landed + 2 * stages * sizeof(uint64_t) > B2O_MAX_DYN_SMEM) --stages;
  cfg.stages = stages;
  cfg.L.accs_off = (size_t)stages * cfg.R * sizeof(double);
  cfg.L.bar_off = cfg.L.accs_off + scal;
  a.landed_off = (uint32_t)(cfg.L.bar_off + (size_t)2 * stages * sizeof(uint64_t));
  cfg.L.total = a.landed_off + landed;
  a.stages = stages;
  a.scal_off = (uint32_t)cfg.L.accs_off;
  a.bar_off = (uint32_t)cfg.L.bar_off;
  a.alpha = alpha;
  a.beta = beta;
  a.gamma = q->gamma;
  a.scaling = q->scaling ? 1 : 0;
  a.x_al16 = ((uintptr_t)x % 16) == 0 && (nrhs == 1 || ldx % 2 == 0);
  a.res_al16 = ((uintptr_t)res % 16) == 0 && (nrhs == 1 || ldr % 2 == 0);
  a.partials = c->d_partials;
  a.bar = c->d_bar;
  a.bar_target = c->bar_base + (unsigned long long)cfg.grid;
  int st;
  switch (cfg.R) {
    case 1024: st = launch_persistent(c, qn_twoloop_multi_kernel<1024, NR>, cfg, a, true); break;
    case 2048: st = launch_persistent(c, qn_twoloop_multi_kernel<2048, NR>, cfg, a, true); break;
    case 4096: st = launch_persistent(c, qn_twoloop_multi_kernel<4096, NR>, cfg, a, true); break;
    default: B2O_FAIL(B2O_EARG, "bad tile_rows");
  }
  if (st == B2O_OK) c->bar_base += (unsigned long long)cfg.grid * (unsigned long long)(2 * na);
  return st;
}

// mul!(Res, op, X, α, β) with n x nrhs column-major matrices (leading dimensions ldr, ldx)
extern "C" int b2o_qn_apply_multi(b2o_qn *q, void *res_, int64_t ldr, const void *x_, int64_t ldx, int64_t len, int nrhs,
                                  double alpha, double beta) {
  if (!q) B2O_FAIL(B2O_EARG, "null operator");
  if (len != q->n) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if (nrhs < 0) B2O_FAIL(B2O_EARG, "nrhs must be >= 0");
  if (nrhs == 0 || q->n == 0) return B2O_OK;
  if (!res_ || !x_) B2O_FAIL(B2O_EARG, "null matrix");
  if (nrhs > 1 && (ldr < q->n || ldx < q->n)) B2O_FAIL(B2O_EARG, "leading dimension smaller than n");
  if (((uintptr_t)res_ | (uintptr_t)x_) % (uintptr_t)q->esize) B2O_FAIL(B2O_EARG, "matrices must be aligned to the element size");
  b2o_ctx *c = q->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  if (q->esize == 4) return qn32_apply_multi(q, (float *)res_, ldr, (const float *)x_, ldx, nrhs, alpha, beta);
  double *res = (double *)res_;
  const double *x = (const double *)x_;
  const bool twoloop = q->kind == 0 && q->inverse && !q->inv_compact;
  const bool split_ranks = c->nranks > 1 && !c->mbox_ready;
  CompactArgs base;
  if (!twoloop) {
    if (q->kind == 0 && !q->inverse && q->fwd_compact && q->w_dirty) B2O_TRY(build_forward_W(q));
    if (q->kind == 0 && q->inverse && q->w_dirty) B2O_TRY(build_inverse_W(q));
    compact_columns(q, base, alpha, beta);
  }
  if (twoloop && nrhs > 1 && c->nranks <= 1 && c->twoloop_block) {
    int o2n[B2O_MAX_MEM];
    if (active_old_to_new(q, o2n) > 0) {
      // block two-loop recursion: the update / dot columns of a sweep are staged once for up to 8 right-hand sides
      for (int r0 = 0; r0 < nrhs;) {
        const int left = nrhs - r0, k = std::min(left, left > 4 ? 8 : 4);
        if (left == 1) B2O_TRY(qn_apply_dev(q, res + (int64_t)r0 * ldr, x + (int64_t)r0 * ldx, alpha, beta));
        else if (k > 4) B2O_TRY(twoloop_multi_launch<8>(q, res + (int64_t)r0 * ldr, ldr, x + (int64_t)r0 * ldx, ldx, k, alpha, beta));
        else B2O_TRY(twoloop_multi_launch<4>(q, res + (int64_t)r0 * ldr, ldr, x + (int64_t)r0 * ldx, ldx, k, alpha, beta));
        r0 += k;
      }
      return B2O_OK;
    }
  }
  if (nrhs == 1 || twoloop || split_ranks || base.ncols == 0 || base.ncols * 4 > B2O_MULTI_MAXV) {
    // row-partitioned two-loop handles (dependent all-reduces) and NCCL mode (split launches): column by column
    for (int r = 0; r < nrhs; ++r) B2O_TRY(qn_apply_dev(q, res + (int64_t)r * ldr, x + (int64_t)r * ldx, alpha, beta));
    return B2O_OK;
  }
  const bool can8 = base.ncols * 8 <= B2O_MULTI_MAXV;
  for (int r0 = 0; r0 < nrhs;) {
    const int left = nrhs - r0;
    if (can8 && left > 4) {
      const int k = std::min(left, 8);
      if (c->multi_mma && base.ncols <= 8 * MM_MAXG) B2O_TRY(multi_mma_launch(q, base, res + (int64_t)r0 * ldr, ldr, x + (int64_t)r0 * ldx, ldx, k));
      else B2O_TRY(multi_launch<8>(q, base, res + (int64_t)r0 * ldr, ldr, x + (int64_t)r0 * ldx, ldx, k));
      r0 += k;
    } else if (left > 2) {
      const int k = std::min(left, 4);
      B2O_TRY(multi_launch<4>(q, base, res + (int64_t)r0 * ldr, ldr, x + (int64_t)r0 * ldx, ldx, k));
      r0 += k;
    } else {
      B2O_TRY(multi_launch<2>(q, base, res + (int64_t)r0 * ldr, ldr, x + (int64_t)r0 * ldx, ldx, left));   // 1 or 2 columns left
      r0 += left;
    }
  }
  return B2O_OK;
}

// Gram entries involving ring slot `k` (after s_k, y_k were stored): s_k·y_j, s_j·y_k, y_k·y_j (and s_k·s_j) for every slot j
static int update_gram(b2o_qn *q, int k) {
  b2o_ctx *c = q->ctx;
  const int m = q->mem;
  const double *sk = q->col(q->S, k), *yk = q->col(q->Y, k);
  for (int j0 = 0; j0 < m; j0 += 2) {
    const double *u[8], *v[8];
    int np = 0, idx[8][2];
    for (int j = j0; j < std::min(m, j0 + 2); ++j) {
      u[np] = sk; v[np] = q->col(q->Y, j); idx[np][0] = 0; idx[np][1] = j; np++;   // s_k·y_j
      u[np] = q->col(q->S, j); v[np] = yk; idx[np][0] = 1; idx[np][1] = j; np++;   // s_j·y_k
      u[np] = yk; v[np] = q->col(q->Y, j); idx[np][0] = 2; idx[np][1] = j; np++;   // y_k·y_j
      if (q->fwd_compact) { u[np] = sk; v[np] = q->col(q->S, j); idx[np][0] = 3; idx[np][1] = j; np++; }   // s_k·s_j
    }
    double h[8];
    B2O_TRY(b2o_pair_dots(c, np, u, v, q->n, c->d_dots + 300));
    B2O_TRY(b2o_read_scalars(c, c->d_dots + 300, np, h));
    for (int p = 0; p < np; ++p) {
      const int j = idx[p][1];
      if (idx[p][0] == 0) q->SY[(size_t)k * m + j] = h[p];
      else if (idx[p][0] == 1) q->SY[(size_t)j * m + k] = h[p];
      else if (idx[p][0] == 2) q->YY[(size_t)k * m + j] = q->YY[(size_t)j * m + k] = h[p];
      else q->SS[(size_t)k * m + j] = q->SS[(size_t)j * m + k] = h[p];
    }
  }
  q->w_dirty = true;
  return B2O_OK;
}

// "inverse_mode": 0 = two-loop recursion (reference algorithm, default), 1 = compact representation (half the DRAM traffic,
// ONE all-reduce instead of 2m dependent ones; different rounding, same operator)
extern "C" int b2o_qn_set_option(b2o_qn *q, const char *key, int64_t value) {
  if (!q || !key) B2O_FAIL(B2O_EARG, "null argument");
  if (strcmp(key, "push_mode")) B2O_F64_ONLY(q, "this option");
  if (!strcmp(key, "inverse_mode")) {
    if (!(q->kind == 0 && q->inverse)) B2O_FAIL(B2O_EARG, "inverse_mode applies to InverseLBFGSOperator");
    if (value != 0 && value != 1) B2O_FAIL(B2O_EARG, "inverse_mode must be 0 (two-loop) or 1 (compact)");
    b2o_ctx *c = q->ctx;
    B2O_CUDA(cudaSetDevice(c->device));
    if (value == 1 && !q->inv_compact) {
      const int m = q->mem;
      if (!q->d_W) {
        B2O_CUDA(cudaMalloc(&q->d_W, sizeof(double) * 4 * m * m));
        B2O_CUDA(cudaMallocHost(&q->h_W, sizeof(double) * 4 * m * m));
      }
      q->SY.assign((size_t)m * m, 0.0);
      q->YY.assign((size_t)m * m, 0.0);
      q->inv_compact = true;
      for (int k = 0; k < m; ++k)
        if (q->ys[k] != 0) B2O_TRY(update_gram(q, k));   // pairs pushed before the mode was enabled
      q->w_dirty = true;
    } else if (value == 0) {
      q->inv_compact = false;
    }
    return B2O_OK;
  }
  if (!strcmp(key, "push_mode")) {
    // 1 (default) = the a_k rebuild of push! (L-BFGS forward form and L-SR1) runs on the streaming apply kernel;
    // 0 = generic multi-dot + linear-combination passes
    if (value != 0 && value != 1) B2O_FAIL(B2O_EARG, "push_mode must be 0 or 1");
    q->push_streamed = value == 1;
    return B2O_OK;
  }
  if (!strcmp(key, "forward_mode")) {
    // 0 = the reference's a_k/b_k form (default); 1 = compact form: push! costs O(m) dots instead of O(m^2) vector passes
    if (!(q->kind == 0 && !q->inverse)) B2O_FAIL(B2O_EARG, "forward_mode applies to the forward LBFGSOperator");
    if (value != 0 && value != 1) B2O_FAIL(B2O_EARG, "forward_mode must be 0 or 1");
    b2o_ctx *c = q->ctx;
    B2O_CUDA(cudaSetDevice(c->device));
    if (value == 1 && !q->fwd_compact) {
      const int m = q->mem;
      if (!q->d_W) {
        B2O_CUDA(cudaMalloc(&q->d_W, sizeof(double) * 4 * m * m));
        B2O_CUDA(cudaMallocHost(&q->h_W, sizeof(double) * 4 * m * m));
      }
      q->SY.assign((size_t)m * m, 0.0);
      q->YY.assign((size_t)m * m, 0.0);
      q->SS.assign((size_t)m * m, 0.0);
      q->fwd_compact = true;
      for (int k = 0; k < m; ++k)
        if (q->ys[k] != 0) B2O_TRY(update_gram(q, k));
      q->w_dirty = true;
    } else if (value == 0 && q->fwd_compact) {
      for (int k = 0; k < q->mem; ++k)
        if (q->ys[k] != 0) B2O_FAIL(B2O_EUNSUPPORTED, "cannot leave the compact forward mode once pairs were pushed (a_k, b_k were not built); reset! first");
      q->fwd_compact = false;
    }
    return B2O_OK;
  }
  B2O_FAIL(B2O_EARG, "unknown option '%s'", key);
}

extern "C" int b2o_qn_apply(b2o_qn *q, void *res, int64_t res_len, const void *x, int64_t x_len, double alpha,
                            double beta) {
  if (!q) B2O_FAIL(B2O_EARG, "null operator");
  if (x_len != q->n || res_len != q->n) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if ((!res || !x) && q->n > 0) B2O_FAIL(B2O_EARG, "null vector");
  if (((uintptr_t)res | (uintptr_t)x) % (uintptr_t)q->esize) B2O_FAIL(B2O_EARG, "vectors must be aligned to the element size");
  B2O_CUDA(cudaSetDevice(q->ctx->device));
  if (q->esize == 4) return qn32_apply(q, (float *)res, (const float *)x, alpha, beta);
  return qn_apply_dev(q, (double *)res, (const double *)x, alpha, beta);
}

static int ensure_stage(b2o_ctx *c, size_t bytes) {
  if (c->stage_bytes >= bytes) return B2O_OK;
  if (c->stage_x) cudaFree(c->stage_x);
  if (c->stage_res) cudaFree(c->stage_res);
  c->stage_x = c->stage_res = nullptr;
  c->stage_bytes = 0;
  cudaError_t e1 = cudaMalloc(&c->stage_x, bytes), e2 = cudaMalloc(&c->stage_res, bytes);
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    cudaGetLastError();
    B2O_FAIL(B2O_ENOMEM, "staging allocation of %zu bytes failed", bytes);
  }
  c->stage_bytes = bytes;
  return B2O_OK;
}

static int ensure_pipeline(b2o_ctx *c) {
  if (c->s_in) return B2O_OK;
  B2O_CUDA(cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking));
  B2O_CUDA(cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < 16; ++i) {
    B2O_CUDA(cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming));
    B2O_CUDA(cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming));
  }
  B2O_CUDA(cudaEventCreateWithFlags(&c->ev_start, cudaEventDisableTiming));
  return B2O_OK;
}

// Host-buffer apply.  For the two-phase operators (forward L-BFGS, L-SR1) the transfers are pipelined with the kernels in
// row chunks: chunk c's dots run while chunk c+1 is still crossing PCIe, and res chunks stream back while later chunks combine
// (phase 2 needs all dots, so the H2D and D2H halves cannot overlap each other).  Chunk results are bit-identical to a single
// launch's only per chunk order: the dots are accumulated chunk by chunk in a fixed order (deterministic).
extern "C" int b2o_qn_apply_host(b2o_qn *q, void *res_host, const void *x_host, int64_t len, double alpha, double beta) {
  if (!q) B2O_FAIL(B2O_EARG, "null operator");
  if (len != q->n) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  b2o_ctx *c = q->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  if (q->esize == 4) {
    // Float32 handles: staged copy-in, apply, copy-out (no chunked transfer / compute pipeline)
    const size_t b32 = (size_t)q->n * sizeof(float);
    B2O_TRY(ensure_stage(c, std::max<size_t>(b32, 16)));
    B2O_CUDA(cudaMemcpyAsync(c->stage_x, x_host, b32, cudaMemcpyHostToDevice, c->stream));
    if (beta != 0.0) B2O_CUDA(cudaMemcpyAsync(c->stage_res, res_host, b32, cudaMemcpyHostToDevice, c->stream));
    B2O_TRY(qn32_apply(q, (float *)c->stage_res, (const float *)c->stage_x, alpha, beta));
    B2O_CUDA(cudaMemcpyAsync(res_host, c->stage_res, b32, cudaMemcpyDeviceToHost, c->stream));
    B2O_CUDA(cudaStreamSynchronize(c->stream));
    return B2O_OK;
  }
  const size_t bytes = (size_t)q->n * sizeof(double);
  B2O_TRY(ensure_stage(c, std::max<size_t>(bytes, 16)));
  double *dx = (double *)c->stage_x, *dres = (double *)c->stage_res;
  const bool two_phase = !(q->kind == 0 && q->inverse);
  const int64_t align = B2O_PITCH_ALIGN;
  int nch = c->host_chunks;
  if (!two_phase || q->n < 4 * align * nch) nch = 1;
  if (nch == 1) {
    B2O_CUDA(cudaMemcpyAsync(dx, x_host, bytes, cudaMemcpyHostToDevice, c->stream));
    if (beta != 0.0) B2O_CUDA(cudaMemcpyAsync(dres, res_host, bytes, cudaMemcpyHostToDevice, c->stream));
    B2O_TRY(qn_apply_dev(q, dres, dx, alpha, beta));
    B2O_CUDA(cudaMemcpyAsync(res_host, dres, bytes, cudaMemcpyDeviceToHost, c->stream));
    B2O_CUDA(cudaStreamSynchronize(c->stream));
    return B2O_OK;
  }
  B2O_TRY(ensure_pipeline(c));
  // the chunked path launches the phases itself: the compact forward form needs its middle matrix first (qn_apply_dev does this
  // for the single-launch path)
  if (q->kind == 0 && !q->inverse && q->fwd_compact && q->w_dirty) B2O_TRY(build_forward_W(q));
  CompactArgs a;
  compact_columns(q, a, alpha, beta);
  const int64_t rows_per = ((q->n + nch - 1) / nch + align - 1) / align * align;
  auto lo = [&](int ch) { return std::min<int64_t>(q->n, (int64_t)ch * rows_per); };
  // copy-in stream starts after whatever is already queued on the compute stream (previous users of the staging buffers)
  B2O_CUDA(cudaEventRecord(c->ev_start, c->stream));
  B2O_CUDA(cudaStreamWaitEvent(c->s_in, c->ev_start, 0));
  B2O_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_start, 0));
  for (int ch = 0; ch < nch; ++ch) {
    const int64_t r0 = lo(ch), r1 = lo(ch + 1);
    if (r1 > r0) {
      B2O_CUDA(cudaMemcpyAsync(dx + r0, (const double *)x_host + r0, (size_t)(r1 - r0) * 8, cudaMemcpyHostToDevice, c->s_in));
      if (beta != 0.0)
        B2O_CUDA(cudaMemcpyAsync(dres + r0, (const double *)res_host + r0, (size_t)(r1 - r0) * 8, cudaMemcpyHostToDevice, c->s_in));
    }
    B2O_CUDA(cudaEventRecord(c->ev_in[ch], c->s_in));
  }
  if (a.ncols > 0) {
    for (int ch = 0; ch < nch; ++ch) {
      B2O_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in[ch], 0));
      B2O_TRY(compact_launch_rows(q, a, dres, dx, lo(ch), lo(ch + 1), MODE_PHASE1, ch > 0));
    }
    B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots, a.ncols));
  } else {
    B2O_CUDA(cudaStreamWaitEvent(c->stream, c->ev_in[nch - 1], 0));
  }
  for (int ch = 0; ch < nch; ++ch) {
    const int64_t r0 = lo(ch), r1 = lo(ch + 1);
    B2O_TRY(compact_launch_rows(q, a, dres, dx, r0, r1, MODE_PHASE2, 0));
    B2O_CUDA(cudaEventRecord(c->ev_k[ch], c->stream));
    B2O_CUDA(cudaStreamWaitEvent(c->s_out, c->ev_k[ch], 0));
    if (r1 > r0)
      B2O_CUDA(cudaMemcpyAsync((double *)res_host + r0, dres + r0, (size_t)(r1 - r0) * 8, cudaMemcpyDeviceToHost, c->s_out));
  }
  B2O_CUDA(cudaStreamSynchronize(c->s_out));
  B2O_CUDA(cudaStreamSynchronize(c->stream));
  return B2O_OK;
}

extern "C" int b2o_qn_apply_bytes(b2o_qn *q, double beta, double *bytes) {
  if (!q || !bytes) B2O_FAIL(B2O_EARG, "null argument");
  int slots[B2O_MAX_MEM];
  const int na = active_old_to_new(q, slots);
  double per_row;
  if (q->kind == 0 && q->inverse && q->inv_compact) per_row = na > 0 ? 4.0 * na + 3.0 : 2.0;   // compact: (4m+3) n E
  else if (q->kind == 0 && q->inverse) per_row = na > 0 ? 8.0 * na + 2.0 : 2.0;   // (8m+2) n E   SURVEY App. A
  else if (q->kind == 0) per_row = na > 0 ? 4.0 * na + 3.0 : 2.0;            // (4m+3) n E
  else per_row = na > 0 ? 2.0 * na + 3.0 : 2.0;                              // (2m+3) n E
  if (beta != 0.0) per_row += 1.0;
  *bytes = per_row * (double)q->esize * (double)q->n;
  return B2O_OK;
}

// ------------------------------------------------------------------ push!
static int lincomb_launch(b2o_ctx *c, LincombArgs &a) {
  a.partials = c->d_partials;
  a.arrive = c->d_bar + 1;
  int grid = ew_grid(c, a.n);
  lincomb_kernel<<<grid, 256, 0, c->stream>>>(a);
  c->launches++;
  B2O_CUDA(cudaGetLastError());
  return B2O_OK;
}

// dots of `ncols` columns against v into d_out[0..ncols) (8 per pass)
static int multi_dots(b2o_ctx *c, int ncols, const double *const *cols, const double *v, int64_t n, double *d_out) {
  for (int off = 0; off < ncols; off += 8) {
    int np = std::min(8, ncols - off);
    const double *u[8], *w[8];
    for (int p = 0; p < np; ++p) {
      u[p] = cols[off + p];
      w[p] = v;
    }
    B2O_TRY(b2o_pair_dots(c, np, u, w, n, d_out + off));
  }
  return B2O_OK;
}

// push_common!  src/lbfgs.jl:210-255.  s, y device vectors (y may be the damped y); ys, yy host scalars.
static int lbfgs_push_common(b2o_qn *q, const double *s, const double *y, double ys, double yy) {
  b2o_ctx *c = q->ctx;
  const int64_t n = q->n;
  const int mem = q->mem, ins = q->ins0;
  const size_t bytes = (size_t)n * sizeof(double);
  B2O_CUDA(cudaMemcpyAsync(q->col(q->S, ins), s, bytes, cudaMemcpyDeviceToDevice, c->stream));   // :220
  B2O_CUDA(cudaMemcpyAsync(q->col(q->Y, ins), y, bytes, cudaMemcpyDeviceToDevice, c->stream));   // :221
  q->ys[ins] = ys;                                                                              // :222
  if (q->scaling) {                                                                             // :223-227
    if (q->gamma != 0) q->opnorm_ub -= 1 / q->gamma;
    q->gamma = ys / yy;
    if (q->gamma != 0) q->opnorm_ub += 1 / q->gamma;
  }
  if (!q->inverse && q->fwd_compact) {
    // compact forward form: no a_k / b_k vectors.  ‖b‖² = y·y / ys keeps opnorm_upper_bound exact (src/lbfgs.jl:231-234).
    q->opnorm_ub -= q->aux[ins] * q->aux[ins];
    q->aux[ins] = sqrt(yy / ys);
    q->opnorm_ub += q->aux[ins] * q->aux[ins];
    B2O_TRY(update_gram(q, ins));
  } else if (!q->inverse) {
    double *bi = q->col(q->B, ins);
    q->opnorm_ub -= q->aux[ins] * q->aux[ins];                                                  // :231
    if (n > 0) {
      div_kernel<<<ew_grid(c, n), 256, 0, c->stream>>>(bi, y, sqrt(ys), n, ew_vec_ok(bi, y, nullptr));                     // :232 b = y ./ sqrt(ys)
      c->launches++;
    }
    {
      const double *u[1] = {bi}, *v[1] = {bi};
      double nb2 = 0;
      B2O_TRY(b2o_pair_dots(c, 1, u, v, n, c->d_dots + 300));
      B2O_TRY(b2o_read_scalars(c, c->d_dots + 300, 1, &nb2));
      q->aux[ins] = sqrt(nb2);                                                                  // :233 norm_b
    }
    q->opnorm_ub += q->aux[ins] * q->aux[ins];
    // rebuild every a_k, oldest -> newest (the new pair is last): k = mod(insert+i-1, mem)+1   :236-250
    int prev[B2O_MAX_MEM];
    int nprev = 0;
    for (int i = 1; i <= mem; ++i) {
      const int k = pmod(ins + i, mem);
      if (q->ys[k] == 0) continue;
      const double *sk = q->col(q->S, k);
      double *ak = q->col(q->A, k);
      LincombArgs a;
      memset(&a, 0, sizeof(a));
      const double *dcols[B2O_MAX_COLS];
      for (int j = 0; j < nprev; ++j) {
        // a[k] .+= dot(b[l], s[k]) .* b[l] ; a[k] .-= dot(a[l], s[k]) .* a[l]                   :244-245
        a.cols[2 * j] = q->col(q->B, prev[j]);
        a.sign[2 * j] = +1;
        a.cols[2 * j + 1] = q->col(q->A, prev[j]);
        a.sign[2 * j + 1] = -1;
        a.cdiv[2 * j] = a.cdiv[2 * j + 1] = 1.0;
        dcols[2 * j] = a.cols[2 * j];
        dcols[2 * j + 1] = a.cols[2 * j + 1];
      }
      a.nterms = 2 * nprev;
      if (n > 0 && 2 * nprev <= B2O_MAX_COLS && q->push_streamed) {
        // a_k = (operator truncated to the pairs older than k) * s_k: ONE launch of the streaming apply kernel (TMA ring, all
        // 2(k-1) dots + combine + the dot s_k·a_k), then the scaling pass                          :239-248
        CompactArgs ca;
        memset(&ca, 0, sizeof(ca));
        for (int j = 0; j < nprev; ++j) {
          ca.cols[2 * j] = q->col(q->A, prev[j]);
          ca.cols[2 * j + 1] = q->col(q->B, prev[j]);
          ca.cdiv[2 * j] = ca.cdiv[2 * j + 1] = 1.0;
        }
        ca.ncols = 2 * nprev;
        ca.alpha = 1.0;
        ca.beta = 0.0;
        ca.gamma = q->gamma;
        ca.scaling = 1;
        int grid = 0;
        if (ca.ncols == 0) {
          B2O_TRY(compact_launch_rows(q, ca, ak, sk, 0, n, MODE_PHASE2, 0, OP_PUSH_A, &grid));
        } else if (c->nranks <= 1 || c->mbox_ready) {
          B2O_TRY(compact_launch_rows(q, ca, ak, sk, 0, n, MODE_FUSED, 0, OP_PUSH_A, &grid));
        } else {
          B2O_TRY(compact_launch_rows(q, ca, ak, sk, 0, n, MODE_PHASE1, 0, OP_PUSH_A, &grid));
          B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots, ca.ncols));
          B2O_TRY(compact_launch_rows(q, ca, ak, sk, 0, n, MODE_PHASE2, 0, OP_PUSH_A, &grid));
        }
        sum_partials_kernel<<<1, 32, 0, c->stream>>>(c->d_partials + (size_t)grid * ca.ncols, grid, c->d_dots + 256);
        c->launches++;
        B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots + 256, 1));
        div_sqrt_dev_kernel<<<ew_grid(c, n), 256, 0, c->stream>>>(ak, ak, c->d_dots + 256, n, ew_vec_ok(ak, nullptr, nullptr));     // :248
        c->launches++;
        prev[nprev++] = k;
        continue;
      }
      B2O_TRY(multi_dots(c, a.nterms, dcols, sk, n, c->d_dots));
      a.base_mode = 0;  // a[k] .= s[k] ./ γ                                                     :239
      a.P1 = sk;
      a.g = q->gamma;
      a.dots = c->d_dots;
      a.u0 = sk;
      a.out = ak;
      a.n = n;
      a.red = c->d_dots + 256;
      if (n > 0) {
        B2O_TRY(lincomb_launch(c, a));
        B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots + 256, 2));
        div_sqrt_dev_kernel<<<ew_grid(c, n), 256, 0, c->stream>>>(ak, ak, c->d_dots + 256, n, ew_vec_ok(ak, nullptr, nullptr));   // :248
        c->launches++;
      }
      prev[nprev++] = k;
    }
    B2O_CUDA(cudaGetLastError());
  }
  if (q->inverse && q->inv_compact) B2O_TRY(update_gram(q, ins));
  q->w_dirty = true;
  q->ins0 = pmod(ins + 1, mem);                                                                 // :253
  return B2O_OK;
}

static int check_vec(const b2o_qn *q, const void *p, int64_t len) {
  if (len != q->n) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if (!p && q->n > 0) B2O_FAIL(B2O_EARG, "null vector");
  if ((uintptr_t)p % (uintptr_t)q->esize) B2O_FAIL(B2O_EARG, "vectors must be aligned to the element size");
  return B2O_OK;
}

static int lsr1_push(b2o_qn *q, const double *s, const double *y, int *accepted);

extern "C" int b2o_lbfgs_push_damped_fwd(b2o_qn *q, const void *s_, const void *y_, void *Bs_, int64_t len, int *accepted) {
  if (!q || q->kind != 0) B2O_FAIL(B2O_EARG, "not an L-BFGS operator");
  if (!q->damped) B2O_FAIL(B2O_ESTATE, "This push! should be used for damped operators");
  if (q->inverse) B2O_FAIL(B2O_ESTATE, "This function be used for forward operators. Use push!(op, s, y, α, g, Bs) instead.");
  B2O_TRY(check_vec(q, s_, len));
  B2O_TRY(check_vec(q, y_, len));
  B2O_TRY(check_vec(q, Bs_, len));
  b2o_ctx *c = q->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  if (q->esize == 4) return qn32_push_damped(q, (const float *)s_, (float *)const_cast<void *>(y_), false, 0.0, nullptr, (float *)Bs_, accepted);
  const double *s = (const double *)s_, *y = (const double *)y_;
  double *Bs = (double *)Bs_;
  const int64_t n = q->n;
  B2O_TRY(qn_apply_dev(q, Bs, s, 1.0, 0.0));                                                    // :305
  const double *u[3] = {y, s, y}, *v[3] = {s, Bs, y};
  double h[3];
  B2O_TRY(b2o_pair_dots(c, 3, u, v, n, c->d_dots + 300));
  B2O_TRY(b2o_read_scalars(c, c->d_dots + 300, 3, h));
  double ys = h[0], sBs = h[1], yy = h[2];
  bool damp = false;
  double th = 0;
  if (ys < (1 - q->sigma2) * sBs) {                                                             // :308-314
    th = q->sigma2 * sBs / (sBs - ys);
    damp = true;
  } else if (ys > (1 + q->sigma3) * sBs) {
    th = q->sigma3 * sBs / (ys - sBs);
    damp = true;
  }
  const double *yuse = y;
  if (damp) {
    B2O_TRY(ew_axpby(c, q->tmp, th, y, 1 - th, Bs, n));                                         // :316 damped y
    ys = th * ys + (1 - th) * sBs;
    yuse = q->tmp;
    if (q->scaling || q->fwd_compact) {   // yy of the DAMPED y: scaling factor, and norm_b of the compact forward form
      const double *u2[1] = {q->tmp}, *v2[1] = {q->tmp};
      B2O_TRY(b2o_pair_dots(c, 1, u2, v2, n, c->d_dots + 300));
      B2O_TRY(b2o_read_scalars(c, c->d_dots + 300, 1, &yy));
    }
  }
  B2O_TRY(lbfgs_push_common(q, s, yuse, ys, yy));
  if (accepted) *accepted = 1;
  return B2O_OK;
}

extern "C" int b2o_lbfgs_push_damped_inv(b2o_qn *q, const void *s_, void *y_, double alpha, const void *g_, void *Bs_,
                                         int64_t len, int *accepted) {
  if (!q || q->kind != 0) B2O_FAIL(B2O_EARG, "not an L-BFGS operator");
  if (!q->damped) B2O_FAIL(B2O_ESTATE, "This push! should be used for damped operators");
  if (!q->inverse) B2O_FAIL(B2O_ESTATE, "This function be used for inverse operators. Use push!(op, s, y, Bs) instead.");
  B2O_TRY(check_vec(q, s_, len));
  B2O_TRY(check_vec(q, y_, len));
  B2O_TRY(check_vec(q, g_, len));
  B2O_TRY(check_vec(q, Bs_, len));
  b2o_ctx *c = q->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  if (q->esize == 4) return qn32_push_damped(q, (const float *)s_, (float *)y_, true, alpha, (const float *)g_, (float *)Bs_, accepted);
  const double *s = (const double *)s_, *g = (const double *)g_;
  double *y = (double *)y_, *Bs = (double *)Bs_;
  const int64_t n = q->n;
  B2O_TRY(ew_axpby(c, Bs, -alpha, g, 0.0, nullptr, n));                                         // :341 Bs .= -α .* g
  const double *u[3] = {y, s, y}, *v[3] = {s, Bs, y};
  double h[3];
  B2O_TRY(b2o_pair_dots(c, 3, u, v, n, c->d_dots + 300));
  B2O_TRY(b2o_read_scalars(c, c->d_dots + 300, 3, h));
  double ys = h[0], sBs = h[1], yy = h[2];
  bool damp = false;
  double th = 0;
  if (ys < (1 - q->sigma2) * sBs) {
    th = q->sigma2 * sBs / (sBs - ys);
    damp = true;
  } else if (ys > (1 + q->sigma3) * sBs) {
    th = q->sigma3 * sBs / (ys - sBs);
    damp = true;
  }
  if (damp) {
    B2O_TRY(ew_axpby(c, y, th, y, 1 - th, Bs, n));                                              // :352 y .= θ y + (1-θ) Bs
    ys = th * ys + (1 - th) * sBs;
    if (q->scaling) {
      const double *u2[1] = {y}, *v2[1] = {y};
      B2O_TRY(b2o_pair_dots(c, 1, u2, v2, n, c->d_dots + 300));
      B2O_TRY(b2o_read_scalars(c, c->d_dots + 300, 1, &yy));
    }
  }
  B2O_TRY(lbfgs_push_common(q, s, y, ys, yy));
  if (accepted) *accepted = 1;
  return B2O_OK;
}

extern "C" int b2o_qn_push(b2o_qn *q, const void *s_, const void *y_, int64_t len, int *accepted) {
  if (!q) B2O_FAIL(B2O_EARG, "null operator");
  B2O_TRY(check_vec(q, s_, len));
  B2O_TRY(check_vec(q, y_, len));
  b2o_ctx *c = q->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  if (q->esize == 4) return qn32_push(q, (const float *)s_, (const float *)y_, accepted);
  const double *s = (const double *)s_, *y = (const double *)y_;
  if (accepted) *accepted = 0;
  if (q->kind == 1) return lsr1_push(q, s, y, accepted);
  if (q->damped) {
    // push!(op,s,y) on a damped operator forwards to push!(op,s,y,similar(s)) (src/lbfgs.jl:274-276), which
    // errors for inverse operators (:296-298).  The library scratch q->q plays the role of similar(s).
    if (q->inverse) B2O_FAIL(B2O_ESTATE, "This function be used for forward operators. Use push!(op, s, y, α, g, Bs) instead.");
    return b2o_lbfgs_push_damped_fwd(q, s_, y_, q->q, len, accepted);
  }
  const double *u[2] = {y, y}, *v[2] = {s, y};
  double h[2];
  B2O_TRY(b2o_pair_dots(c, 2, u, v, q->n, c->d_dots + 300));
  B2O_TRY(b2o_read_scalars(c, c->d_dots + 300, 2, h));
  if (h[0] <= DBL_EPSILON) return B2O_OK;                                                       // :281 rejected
  B2O_TRY(lbfgs_push_common(q, s, y, h[0], h[1]));
  if (accepted) *accepted = 1;
  return B2O_OK;
}

// push!  src/lsr1.jl:119-184
static int lsr1_push(b2o_qn *q, const double *s, const double *y, int *accepted) {
  b2o_ctx *c = q->ctx;
  const int64_t n = q->n;
  const int mem = q->mem;
  const size_t bytes = (size_t)n * sizeof(double);
  double *t = q->tmp;
  B2O_CUDA(cudaMemcpyAsync(t, y, bytes, cudaMemcpyDeviceToDevice, c->stream));                  // :124
  B2O_TRY(qn_apply_dev(q, t, s, -1.0, 1.0));                                                    // :125 ymBs = y - B s
  const double *u[5] = {y, s, y, t, t}, *v[5] = {s, s, y, s, t};
  double h[5];
  B2O_TRY(b2o_pair_dots(c, 5, u, v, n, c->d_dots + 300));
  B2O_TRY(b2o_read_scalars(c, c->d_dots + 300, 5, h));
  const double ys = h[0], sNorm = sqrt(h[1]), yy = h[2];
  const double eps = DBL_EPSILON;
  const bool well_defined = fabs(h[3]) >= eps + eps * sqrt(h[4]) * sNorm;                       // :131
  bool sufficient_curvature = true, scaling_condition = true;
  if (q->scaling) {                                                                             // :135-143
    const double yNorm = sqrt(yy);
    sufficient_curvature = fabs(ys) >= eps * yNorm * sNorm;
    if (sufficient_curvature) {
      const double sf = ys / yy;
      LincombArgs a;
      memset(&a, 0, sizeof(a));
      a.nterms = 0;
      a.base_mode = 1;  // tmp .= y .- s ./ sf
      a.P0 = y;
      a.P1 = s;
      a.g = sf;
      a.dots = c->d_dots;
      a.out = t;
      a.n = n;
      a.red = c->d_dots + 256;
      double r[2] = {0, 0};
      if (n > 0) {
        B2O_TRY(lincomb_launch(c, a));
        B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots + 256, 2));
        B2O_TRY(b2o_read_scalars(c, c->d_dots + 256, 2, r));
      }
      scaling_condition = sqrt(r[1]) >= eps * yNorm * sNorm;
    }
  }
  if (!(well_defined && sufficient_curvature && scaling_condition)) return B2O_OK;              // :145-149 rejected
  const int ins = q->ins0;
  B2O_CUDA(cudaMemcpyAsync(q->col(q->S, ins), s, bytes, cudaMemcpyDeviceToDevice, c->stream));
  B2O_CUDA(cudaMemcpyAsync(q->col(q->Y, ins), y, bytes, cudaMemcpyDeviceToDevice, c->stream));
  q->ys[ins] = ys;
  q->opnorm_ub = 1.0;                                                                           // :156
  if (q->scaling) {
    q->gamma = ys / yy;
    if (q->gamma != 0) q->opnorm_ub = 1 / fabs(q->gamma);
  }
  q->ins0 = pmod(ins + 1, mem);                                                                 // :163
  int prev[B2O_MAX_MEM];
  int nprev = 0;
  for (int i = 1; i <= mem; ++i) {                                                              // :166
    const int k = pmod(q->ins0 + i - 1, mem);
    if (q->ys[k] == 0) continue;
    const double *sk = q->col(q->S, k), *yk = q->col(q->Y, k);
    double *ak = q->col(q->A, k);
    LincombArgs a;
    memset(&a, 0, sizeof(a));
    const double *dcols[B2O_MAX_COLS];
    for (int j = 0; j < nprev; ++j) {
      a.cols[j] = q->col(q->A, prev[j]);                                                        // as = dot(a[l], s[k]) / as[l]
      a.sign[j] = -1;                                                                           // a[k] .-= as .* a[l]   :173-174
      a.cdiv[j] = q->aux[prev[j]];
      dcols[j] = a.cols[j];
    }
    a.nterms = nprev;
    if (n > 0 && q->push_streamed && nprev <= B2O_MAX_COLS) {
      // a_k = (y_k - s_k/γ) - Σ_{l<k} ((a_l·s_k)/as_l) a_l: ONE launch of the streaming kernel (all dots, the combine with the
      // reference's statements, as[k] = a_k·s_k and ‖a_k‖² on the way out)                      :169-179
      CompactArgs ca;
      memset(&ca, 0, sizeof(ca));
      for (int j = 0; j < nprev; ++j) {
        ca.cols[j] = q->col(q->A, prev[j]);
        ca.cdiv[j] = q->aux[prev[j]];
      }
      ca.ncols = nprev;
      ca.alpha = 1.0;
      ca.beta = 0.0;
      ca.gamma = q->gamma;
      ca.scaling = 1;
      ca.y2 = yk;
      int grid = 0;
      if (ca.ncols == 0) {
        B2O_TRY(compact_launch_rows(q, ca, ak, sk, 0, n, MODE_PHASE2, 0, OP_PUSH_L, &grid));
      } else if (c->nranks <= 1 || c->mbox_ready) {
        B2O_TRY(compact_launch_rows(q, ca, ak, sk, 0, n, MODE_FUSED, 0, OP_PUSH_L, &grid));
      } else {
        B2O_TRY(compact_launch_rows(q, ca, ak, sk, 0, n, MODE_PHASE1, 0, OP_PUSH_L, &grid));
        B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots, ca.ncols));
        B2O_TRY(compact_launch_rows(q, ca, ak, sk, 0, n, MODE_PHASE2, 0, OP_PUSH_L, &grid));
      }
      sum_partials_kernel<<<1, 64, 0, c->stream>>>(c->d_partials + (size_t)grid * ca.ncols, grid, c->d_dots + 256);
      c->launches++;
      B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots + 256, 2));
      double r2[2] = {0, 0};
      B2O_TRY(b2o_read_scalars(c, c->d_dots + 256, 2, r2));
      q->aux[k] = r2[0];                                                                          // :177
      if (q->aux[k] != 0) q->opnorm_ub += r2[1] / fabs(q->aux[k]);                                // :179
      prev[nprev++] = k;
      continue;
    }
    B2O_TRY(multi_dots(c, a.nterms, dcols, sk, n, c->d_dots));
    a.base_mode = 1;  // a[k] .= y[k] .- s[k] ./ γ                                               :169
    a.P0 = yk;
    a.P1 = sk;
    a.g = q->gamma;
    a.dots = c->d_dots;
    a.u0 = sk;
    a.out = ak;
    a.n = n;
    a.red = c->d_dots + 256;
    double r[2] = {0, 0};
    if (n > 0) {
      B2O_TRY(lincomb_launch(c, a));
      B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots + 256, 2));
      B2O_TRY(b2o_read_scalars(c, c->d_dots + 256, 2, r));
    }
    q->aux[k] = r[0];                                                                           // :177 as[k] = dot(a[k], s[k])
    if (q->aux[k] != 0) q->opnorm_ub += r[1] / fabs(q->aux[k]);                                 // :179
    prev[nprev++] = k;
  }
  if (accepted) *accepted = 1;
  return B2O_OK;
}

// ------------------------------------------------------------------ solve_shifted_system! / ldiv!  (src/utilities.jl:207-289)
// (B + σI) x = b for a forward L-BFGS operator: 2*mem Sherman-Morrison steps over the rank-one terms a_k a_kᵀ / b_k b_kᵀ
// (Erway, Jain, Marcia 2014).  Per step: all dots p_t·u in multi-dot passes, one combine pass that forms p_i with the two
// dots u·p_i and b·p_i fused in; x is assembled at the end in one pass with the reference's statement order.
extern "C" int b2o_lbfgs_solve_shifted(b2o_qn *q, void *x_, int64_t x_len, const void *b_, int64_t b_len, double sigma) {
  if (!q) B2O_FAIL(B2O_EARG, "null operator");
  B2O_F64_ONLY(q, "solve_shifted_system!");
  if (q->kind != 0 || q->inverse) B2O_FAIL(B2O_EARG, "solve_shifted_system! needs a forward LBFGSOperator");
  if (sigma < 0) B2O_FAIL(B2O_EARG, "σ must be nonnegative");                                   // ArgumentError :213-215
  if (q->fwd_compact) {
    // Woodbury on the compact form: B + σI = τI - Ψ M⁻¹ Ψᵀ, τ = 1/γ + σ, Ψ = [S/γ  Y]  =>
    //   x = b/τ + [S Y] W'' [Sᵀb; Yᵀb],  W'' = diag(1/γ,1) (τ²M - τG)⁻¹ diag(1/γ,1),  G = ΨᵀΨ from the Gram matrices:
    // ONE launch of the compact kernel (two passes over S, Y) instead of 2·mem Sherman-Morrison steps.
    B2O_TRY(check_vec(q, x_, x_len));
    B2O_TRY(check_vec(q, b_, b_len));
    b2o_ctx *c = q->ctx;
    B2O_CUDA(cudaSetDevice(c->device));
    if (q->n == 0) return B2O_OK;
    int sl[B2O_MAX_MEM];
    const int A = active_old_to_new(q, sl), m = q->mem, N = 2 * A;
    const long double g = (long double)q->gamma;            // B0 = I/γ (γ = 1 without scaling or after reset!)
    const long double tau = 1.0L / g + (long double)sigma;
    if (A > 0) {
      std::vector<long double> M, K((size_t)N * N), Inv;
      forward_middle(q, sl, A, g, M);
      for (int i = 0; i < A; ++i)
        for (int j = 0; j < A; ++j) {
          const long double ss = q->SS[(size_t)sl[i] * m + sl[j]], sy = q->SY[(size_t)sl[i] * m + sl[j]],
                            ys_ = q->SY[(size_t)sl[j] * m + sl[i]], yy = q->YY[(size_t)sl[i] * m + sl[j]];
          K[(size_t)i * N + j] = tau * tau * M[(size_t)i * N + j] - tau * ss / (g * g);
          K[(size_t)i * N + A + j] = tau * tau * M[(size_t)i * N + A + j] - tau * sy / g;
          K[(size_t)(A + i) * N + j] = tau * tau * M[(size_t)(A + i) * N + j] - tau * ys_ / g;
          K[(size_t)(A + i) * N + A + j] = tau * tau * M[(size_t)(A + i) * N + A + j] - tau * yy;
        }
      if (!invert_ld(K, Inv, N)) B2O_FAIL(B2O_ESTATE, "solve_shifted_system!: singular system");
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
          long double v = Inv[(size_t)i * N + j];
          if (i < A) v /= g;
          if (j < A) v /= g;
          q->h_W[(size_t)i * N + j] = (double)v;
        }
      B2O_CUDA(cudaMemcpyAsync(q->d_W, q->h_W, sizeof(double) * N * N, cudaMemcpyHostToDevice, c->stream));
      q->w_dirty = true;     // d_W now holds the solve matrix; the apply matrix is rebuilt on the next apply
    }
    CompactArgs a;
    compact_columns(q, a, 1.0, 0.0);
    a.gamma = (double)tau;
    a.scaling = 1;
    a.base_div = 1;
    const int mode = (a.ncols == 0) ? MODE_PHASE2 : MODE_FUSED;
    if (c->nranks > 1 && !c->mbox_ready && a.ncols > 0) {
      B2O_TRY(compact_launch_rows(q, a, (double *)x_, (const double *)b_, 0, q->n, MODE_PHASE1, 0));
      B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots, a.ncols));
      B2O_TRY(compact_launch_rows(q, a, (double *)x_, (const double *)b_, 0, q->n, MODE_PHASE2, 0));
    } else {
      B2O_TRY(compact_launch_rows(q, a, (double *)x_, (const double *)b_, 0, q->n, mode, 0));
    }
    B2O_CUDA(cudaStreamSynchronize(c->stream));   // h_W is reused
    return B2O_OK;
  }
  B2O_TRY(check_vec(q, x_, x_len));
  B2O_TRY(check_vec(q, b_, b_len));
  b2o_ctx *c = q->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  const int64_t n = q->n;
  const int mem = q->mem, max_i = 2 * mem;
  if (n == 0) return B2O_OK;
  if (2 * mem + 1 > B2O_MAX_COLS) B2O_FAIL(B2O_EUNSUPPORTED, "mem too large for solve_shifted_system!");
  if (!q->shifted_p) {
    cudaError_t e = cudaMalloc(&q->shifted_p, sizeof(double) * (size_t)q->pitch * max_i);
    if (e != cudaSuccess) {
      cudaGetLastError();
      B2O_FAIL(B2O_ENOMEM, "solve_shifted_system!: work matrix of %zu bytes: %s", sizeof(double) * (size_t)q->pitch * max_i, cudaGetErrorString(e));
    }
  }
  double *x = (double *)x_;
  const double *b = (const double *)b_;
  const double gamma_inv = 1 / q->gamma;                                                        // :219
  const double x_0 = 1 / (gamma_inv + sigma);
  std::vector<double> v(max_i, 0.0), cx(max_i, 0.0);
  int sign_i = 1;
  for (int i = 1; i <= max_i; ++i) {                                                            // :226
    const int jj = (i + 1) / 2;
    const int k = pmod(q->ins0 + jj, mem);                                                      // k = mod(insert + j - 1, mem) + 1
    const double *u = (sign_i == -1) ? q->col(q->B, k) : q->col(q->A, k);                       // :229
    double *pi = q->shifted_p + (size_t)(i - 1) * q->pitch;
    LincombArgs a;
    memset(&a, 0, sizeof(a));
    const double *dcols[B2O_MAX_COLS];
    int sign_t = 1;
    for (int t = 1; t <= i - 1; ++t) {                                                          // c2 = (sign_t * v[t]) * dot(p_t, u)
      a.cols[t - 1] = q->shifted_p + (size_t)(t - 1) * q->pitch;
      a.sign[t - 1] = +1;
      a.cdiv[t - 1] = sign_t * v[t - 1];
      dcols[t - 1] = a.cols[t - 1];
      sign_t = -sign_t;
    }
    a.nterms = i - 1;
    a.cmul = 1;
    B2O_TRY(multi_dots(c, a.nterms, dcols, u, n, c->d_dots));
    a.base_mode = 2;                                                                            // p_i = x_0 .* u   :231
    a.P1 = u;
    a.g = x_0;
    a.dots = c->d_dots;
    a.u0 = u;                                                                                   // red[0] = u·p_i
    a.u1 = b;                                                                                   // red[1] = p_i·b
    a.out = pi;
    a.n = n;
    a.red = c->d_dots + 256;
    double r[2];
    B2O_TRY(lincomb_launch(c, a));
    B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots + 256, 2));
    B2O_TRY(b2o_read_scalars(c, c->d_dots + 256, 2, r));
    v[i - 1] = 1 / (1 - sign_i * r[0]);                                                         // :242
    cx[i - 1] = sign_i * v[i - 1] * r[1];                                                       // :243
    sign_i = -sign_i;
  }
  // x = x_0 .* b ; x .+= cx_i .* p_i  (i = 1..2mem in order)                                    :221, :243-244
  LincombArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < max_i; ++i) {
    a.cols[i] = q->shifted_p + (size_t)i * q->pitch;
    a.sign[i] = +1;
    a.cdiv[i] = 1.0;
    c->h_scal[i] = cx[i];
  }
  B2O_CUDA(cudaMemcpyAsync(c->d_dots, c->h_scal, sizeof(double) * max_i, cudaMemcpyHostToDevice, c->stream));
  a.nterms = max_i;
  a.cmul = 1;
  a.base_mode = 2;
  a.P1 = b;
  a.g = x_0;
  a.dots = c->d_dots;
  a.out = x;
  a.n = n;
  a.red = c->d_dots + 256;
  B2O_TRY(lincomb_launch(c, a));
  B2O_CUDA(cudaStreamSynchronize(c->stream));   // h_scal is reused by later calls
  return B2O_OK;
}

// ------------------------------------------------------------------ diag! / reset! / state
extern "C" int b2o_qn_diag(b2o_qn *q, void *d, int64_t d_len) {
  if (!q) B2O_FAIL(B2O_EARG, "null operator");
  if (q->kind == 0 && q->inverse)
    B2O_FAIL(B2O_ESTATE, "only the diagonal of a forward L-BFGS approximation is available");  // src/lbfgs.jl:380-382
  if (q->fwd_compact) B2O_FAIL(B2O_EUNSUPPORTED, "diag! needs the a_k/b_k form (forward_mode 0)");
  B2O_TRY(check_vec(q, d, d_len));
  b2o_ctx *c = q->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  if (q->n == 0) return B2O_OK;
  if (q->esize == 4) return qn32_diag(q, (float *)d);
  int slots[B2O_MAX_MEM];
  const int na = active_old_to_new(q, slots);
  DiagArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < na; ++i) {
    a.c0[i] = q->col(q->A, slots[i]);
    a.c1[i] = q->kind == 0 ? q->col(q->B, slots[i]) : nullptr;
    a.cdiv[i] = q->kind == 1 ? q->aux[slots[i]] : 1.0;
  }
  a.nact = na;
  a.kind = q->kind;
  a.scaling = q->scaling ? 1 : 0;
  a.gamma = q->gamma;
  a.d = (double *)d;
  a.n = q->n;
  qn_diag_kernel<double><<<ew_grid(c, q->n), 256, 0, c->stream>>>(a);
  c->launches++;
  B2O_CUDA(cudaGetLastError());
  return B2O_OK;
}

extern "C" int b2o_qn_reset(b2o_qn *q) {
  if (!q) B2O_FAIL(B2O_EARG, "null operator");
  b2o_ctx *c = q->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  const size_t colb = (size_t)q->pitch * (size_t)q->esize * q->mem;
  B2O_CUDA(cudaMemsetAsync(q->S, 0, colb, c->stream));
  B2O_CUDA(cudaMemsetAsync(q->Y, 0, colb, c->stream));
  if (q->A) B2O_CUDA(cudaMemsetAsync(q->A, 0, colb, c->stream));
  if (q->B) B2O_CUDA(cudaMemsetAsync(q->B, 0, colb, c->stream));
  std::fill(q->ys.begin(), q->ys.end(), 0.0);
  // L-BFGS reset! zeroes α but leaves norm_b / opnorm_upper_bound (Q8, src/lbfgs.jl:401-415); L-SR1 zeroes `as`.
  if (q->kind == 1 || q->inverse) std::fill(q->aux.begin(), q->aux.end(), 0.0);
  q->gamma = 1.0;
  q->ins0 = 0;
  q->w_dirty = true;
  if (q->inv_compact || q->fwd_compact) {
    std::fill(q->SY.begin(), q->SY.end(), 0.0);
    std::fill(q->YY.begin(), q->YY.end(), 0.0);
    std::fill(q->SS.begin(), q->SS.end(), 0.0);
  }
  return B2O_OK;
}

static double *qn_base(b2o_qn *q, int which) {
  switch (which) {
    case 0: return q->S;
    case 1: return q->Y;
    case 2: return q->A;
    case 3: return q->B;
  }
  return nullptr;
}
extern "C" int b2o_qn_get_col(b2o_qn *q, int which, int k0, void *dst) {
  if (!q || !dst) B2O_FAIL(B2O_EARG, "null argument");
  double *base = qn_base(q, which);
  if (!base || k0 < 0 || k0 >= q->mem) B2O_FAIL(B2O_EARG, "no such column (which=%d, k0=%d)", which, k0);
  B2O_CUDA(cudaMemcpyAsync(dst, q->colv(base, k0), (size_t)q->n * (size_t)q->esize, cudaMemcpyDeviceToDevice, q->ctx->stream));
  return B2O_OK;
}
extern "C" int b2o_qn_set_col(b2o_qn *q, int which, int k0, const void *src) {
  if (!q || !src) B2O_FAIL(B2O_EARG, "null argument");
  double *base = qn_base(q, which);
  if (!base || k0 < 0 || k0 >= q->mem) B2O_FAIL(B2O_EARG, "no such column (which=%d, k0=%d)", which, k0);
  B2O_CUDA(cudaMemcpyAsync(q->colv(base, k0), src, (size_t)q->n * (size_t)q->esize, cudaMemcpyDeviceToDevice, q->ctx->stream));
  if ((q->inv_compact || q->fwd_compact) && (which == 0 || which == 1)) B2O_TRY(update_gram(q, k0));   // imported pair: refresh its Gram row/column
  q->w_dirty = true;
  return B2O_OK;
}
extern "C" int b2o_qn_get_scalars(b2o_qn *q, int *insert1, double *gamma, double *opnorm_ub, double *ys, double *aux) {
  if (!q) B2O_FAIL(B2O_EARG, "null operator");
  if (insert1) *insert1 = q->ins0 + 1;
  if (gamma) *gamma = q->gamma;
  if (opnorm_ub) *opnorm_ub = q->opnorm_ub;
  if (ys) memcpy(ys, q->ys.data(), sizeof(double) * q->mem);
  if (aux) {
    if (q->kind == 0 && q->inverse) {
      // data.α lives on the device in loop-1 (newest -> oldest) order; scatter back to ring slots
      int o2n[B2O_MAX_MEM];
      const int na = active_old_to_new(q, o2n);
      double h[B2O_MAX_MEM];
      B2O_TRY(b2o_read_scalars(q->ctx, q->d_alpha, B2O_MAX_MEM, h));
      for (int i = 0; i < na; ++i) q->aux[o2n[na - 1 - i]] = h[i];
    }
    memcpy(aux, q->aux.data(), sizeof(double) * q->mem);
  }
  return B2O_OK;
}
extern "C" int b2o_qn_set_scalars(b2o_qn *q, int insert1, double gamma, double opnorm_ub, const double *ys,
                                  const double *aux) {
  if (!q) B2O_FAIL(B2O_EARG, "null operator");
  if (insert1 < 1 || insert1 > q->mem) B2O_FAIL(B2O_EARG, "insert out of range");
  q->ins0 = insert1 - 1;
  q->gamma = gamma;
  q->opnorm_ub = opnorm_ub;
  if (ys) q->ys.assign(ys, ys + q->mem);
  if (aux) q->aux.assign(aux, aux + q->mem);
  q->w_dirty = true;
  return B2O_OK;
}

#include "b2o_qn_f32.inc"
