// b2o_leaf.cu -- leaf operators of src/special-operators.jl and src/linalg.jl:
// opDiagonal, opEye, opOnes, opZeros, opHouseholder (streaming, 16-byte vectorised) and
// opRestriction / opExtension (gather / scatter, bit-exact index work).
#include "b2o_internal.cuh"
#include <algorithm>
#include <unordered_set>

// ------------------------------------------------------------------ streaming elementwise kernel
enum { EW_DIAG = 0, EW_EYE = 1, EW_ZEROS = 2, EW_ONES = 3, EW_HOUSE = 4 };

struct EwArgs {
  int op;
  const double *a;     // DIAG: d ; HOUSE: h
  const double *v;     // input vector
  double *res;
  int64_t nmin;        // rows with the elementwise formula
  int64_t nrow;        // rows of res; [nmin,nrow) get `tail`
  double alpha, beta, tail;
  const double *dscal; // ONES: sum(v) ; HOUSE: dot(h,v)   (device scalar)
};

template <int OP>
__device__ __forceinline__ double ew_one(const EwArgs &p, double a, double v, double r, double scal) {
  double t;
  if (OP == EW_DIAG) t = (p.alpha * a) * v;            // α .* d .* v                  special-operators.jl:127
  else if (OP == EW_EYE) t = p.alpha * v;              // α .* v                        :38
  else if (OP == EW_ZEROS) return (p.beta != 0.0) ? r * p.beta : 0.0;  // res .*= β    :106
  else if (OP == EW_ONES) t = scal;                    // α * sum(v)                    :81
  else t = p.alpha * (v - scal * a);                   // α .* (v .- 2τ .* h)           linalg.jl:79
  return (p.beta != 0.0) ? t + p.beta * r : t;
}

constexpr int EW_UNROLL = 4;

template <int OP, bool VEC>
__global__ void __launch_bounds__(256) ew_kernel(const __grid_constant__ EwArgs p) {
  constexpr bool NEED_A = (OP == EW_DIAG || OP == EW_HOUSE);
  constexpr bool NEED_V = (OP == EW_DIAG || OP == EW_EYE || OP == EW_HOUSE);
  double scal = 0.0;
  if (OP == EW_ONES) scal = p.alpha * (*p.dscal);
  if (OP == EW_HOUSE) scal = 2 * (*p.dscal);
  const bool need_r = p.beta != 0.0;
  if (VEC) {
    const int64_t nvec = p.nmin >> 1;
    const int64_t base = (int64_t)blockIdx.x * (256 * EW_UNROLL) + threadIdx.x;
    double2 a[EW_UNROLL], v[EW_UNROLL], r[EW_UNROLL];
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int64_t i = base + u * 256;
      a[u] = v[u] = r[u] = make_double2(0.0, 0.0);
      if (i < nvec) {
        if (NEED_A) a[u] = ldg_stream2(p.a + 2 * i);
        if (NEED_V) v[u] = ldg_stream2(p.v + 2 * i);
        if (need_r) r[u] = ldg_stream2(p.res + 2 * i);
      }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int64_t i = base + u * 256;
      if (i < nvec) {
        double2 o;
        o.x = ew_one<OP>(p, a[u].x, v[u].x, r[u].x, scal);
        o.y = ew_one<OP>(p, a[u].y, v[u].y, r[u].y, scal);
        stg_stream2(p.res + 2 * i, o);
      }
    }
    // odd element of the elementwise part + the tail are handled by the last block
    if (blockIdx.x == gridDim.x - 1) {
      if ((p.nmin & 1) && threadIdx.x == 0) {
        const int64_t i = p.nmin - 1;
        p.res[i] = ew_one<OP>(p, NEED_A ? p.a[i] : 0.0, NEED_V ? p.v[i] : 0.0, need_r ? p.res[i] : 0.0, scal);
      }
      for (int64_t i = p.nmin + threadIdx.x; i < p.nrow; i += 256) p.res[i] = p.tail;
    }
  } else {
    const int64_t base = (int64_t)blockIdx.x * (256 * EW_UNROLL) + threadIdx.x;
#pragma unroll
    for (int u = 0; u < EW_UNROLL; ++u) {
      const int64_t i = base + u * 256;
      if (i < p.nmin)
        p.res[i] = ew_one<OP>(p, NEED_A ? p.a[i] : 0.0, NEED_V ? p.v[i] : 0.0, need_r ? p.res[i] : 0.0, scal);
    }
    if (blockIdx.x == gridDim.x - 1)
      for (int64_t i = p.nmin + threadIdx.x; i < p.nrow; i += 256) p.res[i] = p.tail;
  }
}

template <int OP>
static int ew_launch(b2o_ctx *c, EwArgs &p) {
  if (p.nrow <= 0) return B2O_OK;
  const bool vec = (((uintptr_t)p.a | (uintptr_t)p.v | (uintptr_t)p.res) % 16) == 0;
  const int64_t units = vec ? (p.nmin >> 1) : p.nmin;
  const int64_t per_block = 256 * EW_UNROLL;
  const int64_t grid = std::max<int64_t>(1, (units + per_block - 1) / per_block);
  if (grid > 0x7fffffff) B2O_FAIL(B2O_EARG, "vector too long");
  if (vec)
    ew_kernel<OP, true><<<(unsigned)grid, 256, 0, c->stream>>>(p);
  else
    ew_kernel<OP, false><<<(unsigned)grid, 256, 0, c->stream>>>(p);
  c->launches++;
  B2O_CUDA(cudaGetLastError());
  return B2O_OK;
}

static int check_ptrs8(const void *a, const void *b, const void *c) {
  if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) % 8) B2O_FAIL(B2O_EARG, "vectors must be 8-byte aligned");
  return B2O_OK;
}

// ------------------------------------------------------------------ opDiagonal
extern "C" int b2o_diag_apply(b2o_ctx *c, int dtype, int64_t nrow, int64_t ncol, const void *d, int64_t d_len, void *res,
                              int64_t res_len, const void *v, int64_t v_len, double alpha, double beta) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  B2O_TRY(b2o_check_dtype_f64(dtype));
  if (nrow < 0 || ncol < 0) B2O_FAIL(B2O_EARG, "negative size");
  if (v_len != ncol || res_len != nrow) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  const int64_t nmin = std::min(nrow, ncol);
  if (d_len < nmin) B2O_FAIL(B2O_EARG, "diagonal shorter than min(nrow,ncol)");
  if (nrow > 0 && (!res || (nmin > 0 && (!d || !v)))) B2O_FAIL(B2O_EARG, "null vector");
  B2O_TRY(check_ptrs8(d, v, res));
  B2O_CUDA(cudaSetDevice(c->device));
  EwArgs p;
  memset(&p, 0, sizeof(p));
  p.op = EW_DIAG;
  p.a = (const double *)d;
  p.v = (const double *)v;
  p.res = (double *)res;
  p.nmin = nmin;
  p.nrow = nrow;
  p.alpha = alpha;
  p.beta = beta;
  p.tail = 0.0;  // res[(n_min+1):end] .= 0 regardless of β                 special-operators.jl:150
  return ew_launch<EW_DIAG>(c, p);
}

// ------------------------------------------------------------------ opEye
extern "C" int b2o_eye_apply(b2o_ctx *c, int dtype, int64_t nrow, int64_t ncol, void *res, int64_t res_len, const void *v,
                             int64_t v_len, double alpha, double beta) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  B2O_TRY(b2o_check_dtype_f64(dtype));
  if (nrow < 0 || ncol < 0) B2O_FAIL(B2O_EARG, "negative size");
  if (v_len != ncol || res_len != nrow) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if (nrow > 0 && (!res || (ncol > 0 && !v))) B2O_FAIL(B2O_EARG, "null vector");
  B2O_TRY(check_ptrs8(nullptr, v, res));
  B2O_CUDA(cudaSetDevice(c->device));
  EwArgs p;
  memset(&p, 0, sizeof(p));
  p.op = EW_EYE;
  p.v = (const double *)v;
  p.res = (double *)res;
  p.nmin = std::min(nrow, ncol);
  p.nrow = nrow;
  p.alpha = alpha;
  p.beta = beta;
  p.tail = (beta == 0.0) ? 0.0 : beta;  // res[(n_min+1):end] .= β  (Q2)     special-operators.jl:39,42
  return ew_launch<EW_EYE>(c, p);
}

// ------------------------------------------------------------------ opZeros
extern "C" int b2o_zeros_apply(b2o_ctx *c, int dtype, int64_t nrow, int64_t ncol, void *res, int64_t res_len, int64_t v_len,
                               double alpha, double beta) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  B2O_TRY(b2o_check_dtype_f64(dtype));
  if (v_len != ncol || res_len != nrow) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if (nrow > 0 && !res) B2O_FAIL(B2O_EARG, "null vector");
  B2O_TRY(check_ptrs8(nullptr, nullptr, res));
  B2O_CUDA(cudaSetDevice(c->device));
  EwArgs p;
  memset(&p, 0, sizeof(p));
  p.op = EW_ZEROS;
  p.res = (double *)res;
  p.nmin = nrow;
  p.nrow = nrow;
  p.alpha = alpha;
  p.beta = beta;
  return ew_launch<EW_ZEROS>(c, p);
}

// ------------------------------------------------------------------ opOnes: sum pass + fill pass in one cooperative launch
struct OnesArgs {
  const double *v;
  double *res;
  int64_t ncol, nrow;
  double alpha, beta;
  double *partials;
  unsigned long long *bar;
  unsigned long long bar_target;
  int vvec, rvec;
};
__global__ void __launch_bounds__(512) ones_kernel(const __grid_constant__ OnesArgs p) {
  __shared__ double sred[16];
  __shared__ double s_c;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  double a0 = 0.0, a1 = 0.0;
  if (p.vvec) {
    const int64_t nvec = p.ncol >> 1;
    for (int64_t i = tid; i < nvec; i += nth) {
      double2 x = ldg_stream2(p.v + 2 * i);
      a0 += x.x;
      a1 += x.y;
    }
    if ((p.ncol & 1) && tid == 0) a0 += p.v[p.ncol - 1];
  } else {
    for (int64_t i = tid; i < p.ncol; i += nth) a0 += p.v[i];
  }
  double s = warp_sum(a0 + a1);
  if (lane == 0) sred[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sred[w];
    p.partials[blockIdx.x] = t;
  }
  grid_barrier(p.bar, p.bar_target);
  if (warp == 0) {
    double t = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(&p.partials[b]);
    t = warp_sum(t);
    if (lane == 0) s_c = p.alpha * t;                 // α * sum(v)                 special-operators.jl:81
  }
  __syncthreads();
  const double cst = s_c;
  const bool need_r = p.beta != 0.0;
  if (p.rvec) {
    const int64_t nvec = p.nrow >> 1;
    for (int64_t i = tid; i < nvec; i += nth) {
      double2 o = make_double2(cst, cst);
      if (need_r) {
        double2 r = *reinterpret_cast<const double2 *>(p.res + 2 * i);
        o.x = cst + p.beta * r.x;
        o.y = cst + p.beta * r.y;
      }
      stg_stream2(p.res + 2 * i, o);
    }
    if ((p.nrow & 1) && tid == 0) p.res[p.nrow - 1] = need_r ? cst + p.beta * p.res[p.nrow - 1] : cst;
  } else {
    for (int64_t i = tid; i < p.nrow; i += nth) p.res[i] = need_r ? cst + p.beta * p.res[i] : cst;
  }
}
static int ones_fused(b2o_ctx *c, double *res, int64_t nrow, const double *v, int64_t ncol, double alpha, double beta) {
  static thread_local int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    B2O_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, ones_kernel, 512, 0));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  OnesArgs p;
  p.v = v;
  p.res = res;
  p.ncol = ncol;
  p.nrow = nrow;
  p.alpha = alpha;
  p.beta = beta;
  p.partials = c->d_partials;
  p.bar = c->d_bar;
  p.vvec = ((uintptr_t)v % 16) == 0;
  p.rvec = ((uintptr_t)res % 16) == 0;
  const int64_t want = (std::max(nrow, ncol) / 2 + 511) / 512;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)c->num_sms * std::min(blocks_per_sm, 4)));
  grid = std::min(grid, B2O_MAX_GRID);
  p.bar_target = c->bar_base + (unsigned long long)grid;
  void *kargs[] = {(void *)&p};
  B2O_CUDA(cudaLaunchCooperativeKernel((const void *)ones_kernel, dim3(grid), dim3(512), kargs, 0, c->stream));
  c->bar_base += (unsigned long long)grid;
  c->launches++;
  return B2O_OK;
}

// ------------------------------------------------------------------ sums (opOnes) -- pair_dots with v = null means "times one"
extern "C" int b2o_ones_apply(b2o_ctx *c, int dtype, int64_t nrow, int64_t ncol, void *res, int64_t res_len, const void *v,
                              int64_t v_len, double alpha, double beta) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  B2O_TRY(b2o_check_dtype_f64(dtype));
  if (v_len != ncol || res_len != nrow) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if ((nrow > 0 && !res) || (ncol > 0 && !v)) B2O_FAIL(B2O_EARG, "null vector");
  B2O_TRY(check_ptrs8(nullptr, v, res));
  B2O_CUDA(cudaSetDevice(c->device));
  if (c->nranks <= 1 && nrow > 0) return ones_fused(c, (double *)res, nrow, (const double *)v, ncol, alpha, beta);
  const double *u[1] = {(const double *)v}, *w[1] = {nullptr};
  B2O_TRY(b2o_pair_dots(c, 1, u, w, ncol, c->d_dots + 320));   // sum(v), all-reduced when row-partitioned
  EwArgs p;
  memset(&p, 0, sizeof(p));
  p.op = EW_ONES;
  p.res = (double *)res;
  p.nmin = nrow;
  p.nrow = nrow;
  p.alpha = alpha;
  p.beta = beta;
  p.dscal = c->d_dots + 320;
  return ew_launch<EW_ONES>(c, p);
}

// ------------------------------------------------------------------ opHouseholder: one cooperative launch
struct HouseArgs {
  const double *h, *v;
  double *res;
  int64_t n;
  double alpha, beta;
  double *partials;
  unsigned long long *bar;
  unsigned long long bar_target;
  int vec;
};
__global__ void __launch_bounds__(512) householder_kernel(const __grid_constant__ HouseArgs p) {
  __shared__ double sred[16];
  __shared__ double s_t2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  double acc0 = 0.0, acc1 = 0.0;
  if (p.vec) {
    const int64_t nvec = p.n >> 1;
    for (int64_t i = tid; i < nvec; i += nth) {
      double2 a = *reinterpret_cast<const double2 *>(p.h + 2 * i), b = *reinterpret_cast<const double2 *>(p.v + 2 * i);
      acc0 = fma(a.x, b.x, acc0);
      acc1 = fma(a.y, b.y, acc1);
    }
    if ((p.n & 1) && tid == 0) acc0 = fma(p.h[p.n - 1], p.v[p.n - 1], acc0);
  } else {
    for (int64_t i = tid; i < p.n; i += nth) acc0 = fma(p.h[i], p.v[i], acc0);
  }
  double s = warp_sum(acc0 + acc1);
  if (lane == 0) sred[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sred[w];
    p.partials[blockIdx.x] = t;
  }
  grid_barrier(p.bar, p.bar_target);
  if (warp == 0) {
    double t = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32) t += __ldcg(&p.partials[b]);
    t = warp_sum(t);
    if (lane == 0) s_t2 = 2 * t;          // 2 * dot(h, v)                      linalg.jl:79
  }
  __syncthreads();
  const double t2 = s_t2;
  const bool need_r = p.beta != 0.0;
  // second pass in reverse so the tail of pass 1 is served from L2
  if (p.vec) {
    const int64_t nvec = p.n >> 1;
    for (int64_t i = nvec - 1 - tid; i >= 0; i -= nth) {
      double2 a = *reinterpret_cast<const double2 *>(p.h + 2 * i), b = *reinterpret_cast<const double2 *>(p.v + 2 * i), o;
      o.x = p.alpha * (b.x - t2 * a.x);
      o.y = p.alpha * (b.y - t2 * a.y);
      if (need_r) {
        double2 r = *reinterpret_cast<const double2 *>(p.res + 2 * i);
        o.x = o.x + p.beta * r.x;
        o.y = o.y + p.beta * r.y;
      }
      stg_stream2(p.res + 2 * i, o);
    }
    if ((p.n & 1) && tid == 0) {
      const int64_t i = p.n - 1;
      double o = p.alpha * (p.v[i] - t2 * p.h[i]);
      p.res[i] = need_r ? o + p.beta * p.res[i] : o;
    }
  } else {
    for (int64_t i = tid; i < p.n; i += nth) {
      double o = p.alpha * (p.v[i] - t2 * p.h[i]);
      p.res[i] = need_r ? o + p.beta * p.res[i] : o;
    }
  }
}

extern "C" int b2o_householder_apply(b2o_ctx *c, int dtype, int64_t n, const void *h, void *res, int64_t res_len,
                                     const void *v, int64_t v_len, double alpha, double beta) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  B2O_TRY(b2o_check_dtype_f64(dtype));
  if (v_len != n || res_len != n) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if (n > 0 && (!h || !res || !v)) B2O_FAIL(B2O_EARG, "null vector");
  B2O_TRY(check_ptrs8(h, v, res));
  if (n == 0) return B2O_OK;
  B2O_CUDA(cudaSetDevice(c->device));
  if (c->nranks > 1) {
    // row-partitioned: dot -> all-reduce -> update
    const double *u[1] = {(const double *)h}, *w[1] = {(const double *)v};
    B2O_TRY(b2o_pair_dots(c, 1, u, w, n, c->d_dots + 320));
    EwArgs p;
    memset(&p, 0, sizeof(p));
    p.op = EW_HOUSE;
    p.a = (const double *)h;
    p.v = (const double *)v;
    p.res = (double *)res;
    p.nmin = n;
    p.nrow = n;
    p.alpha = alpha;
    p.beta = beta;
    p.dscal = c->d_dots + 320;
    return ew_launch<EW_HOUSE>(c, p);
  }
  static thread_local int blocks_per_sm = 0;
  if (!blocks_per_sm) {
    B2O_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, householder_kernel, 512, 0));
    if (blocks_per_sm < 1) blocks_per_sm = 1;
  }
  HouseArgs p;
  p.h = (const double *)h;
  p.v = (const double *)v;
  p.res = (double *)res;
  p.n = n;
  p.alpha = alpha;
  p.beta = beta;
  p.partials = c->d_partials;
  p.bar = c->d_bar;
  p.vec = (((uintptr_t)h | (uintptr_t)v | (uintptr_t)res) % 16) == 0;
  int64_t want = (n / 2 + 511) / 512;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)c->num_sms * std::min(blocks_per_sm, 4)));
  grid = std::min(grid, B2O_MAX_GRID);
  p.bar_target = c->bar_base + (unsigned long long)grid;
  void *kargs[] = {(void *)&p};
  B2O_CUDA(cudaLaunchCooperativeKernel((const void *)householder_kernel, dim3(grid), dim3(512), kargs, 0, c->stream));
  c->bar_base += (unsigned long long)grid;
  c->launches++;
  return B2O_OK;
}

// ------------------------------------------------------------------ ComplexF64 leaves and the conj-sandwich primitives
// The reference is generic in the element type.  For complex operators `mul!` of an adjoint / transpose / conjugate wrapper
// (src/adjtrans.jl:128-136, 196-204) runs  conj!(res); prod!(res, conj.(v), conj(α), conj(β)); conj!(res), opDiagonal's ctprod!
// multiplies by conj.(d) (src/special-operators.jl:140) and mulHouseholder!'s `dot(h, v)` conjugates h (src/linalg.jl:79).
// Vectors are interleaved (re, im) pairs = one 16-byte element; complex products are Julia's plain
// (ac - bd) + (ad + bc)i without contraction (the library is built with -fmad=false), so the elementwise operators stay
// bit-exact against the oracle.
struct cplx {
  double re, im;
};
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return cplx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ cplx cadd(cplx a, cplx b) { return cplx{a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ cplx csub(cplx a, cplx b) { return cplx{a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ cplx cld(const double *p, int64_t i) {
  const double2 v = ldg_stream2(p + 2 * i);
  return cplx{v.x, v.y};
}
__device__ __forceinline__ void cst(double *p, int64_t i, cplx v) { stg_stream2(p + 2 * i, make_double2(v.re, v.im)); }

enum { CEW_DIAG = 0, CEW_EYE = 1, CEW_ZEROS = 2, CEW_HOUSE = 3, CEW_CONJ = 4 };
struct CewArgs {
  const double *a;      // DIAG: d ; HOUSE: h (interleaved complex)
  const double *v;
  double *res;
  int64_t nmin, nrow;
  cplx alpha, beta, tail;
  int conj_a;           // DIAG: use conj.(d) (ctprod!)
  int beta_nz;
  const double *dscal;  // HOUSE: dot(h, v) as (re, im) device scalars
};
template <int OP>
__global__ void __launch_bounds__(256) cew_kernel(const __grid_constant__ CewArgs p) {
  cplx tau{0.0, 0.0};
  if (OP == CEW_HOUSE) tau = cmul(cplx{2.0, 0.0}, cplx{p.dscal[0], p.dscal[1]});      // 2 * dot(h, v)
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.nrow; i += stride) {
    cplx out;
    if (i >= p.nmin) {
      out = p.tail;
    } else if (OP == CEW_CONJ) {
      const cplx x = cld(p.v, i);
      out = cplx{x.re, -x.im};
    } else {
      cplx t;
      if (OP == CEW_DIAG) {
        cplx d = cld(p.a, i);
        if (p.conj_a) d.im = -d.im;
        t = cmul(cmul(p.alpha, d), cld(p.v, i));                                      // α .* d .* v
      } else if (OP == CEW_EYE) {
        t = cmul(p.alpha, cld(p.v, i));                                               // α .* v
      } else if (OP == CEW_ZEROS) {
        t = cplx{0.0, 0.0};
      } else {
        t = cmul(p.alpha, csub(cld(p.v, i), cmul(tau, cld(p.a, i))));                 // α .* (v .- 2 dot(h,v) .* h)
      }
      if (OP == CEW_ZEROS) out = p.beta_nz ? cmul(cld(p.res, i), p.beta) : t;        // res .*= β
      else out = p.beta_nz ? cadd(t, cmul(p.beta, cld(p.res, i))) : t;
    }
    cst(p.res, i, out);
  }
}
// dot(h, v) = Σ conj(h_i) v_i: per-block partials of (re, im), the last block sums them in block order (deterministic)
__global__ void __launch_bounds__(256) cdot_kernel(const double *h, const double *v, int64_t n, double *partials, double *out,
                                                   unsigned long long *arrive) {
  __shared__ double sred[2][8];
  __shared__ bool is_last;
  double re = 0.0, im = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const cplx a = cld(h, i), b = cld(v, i);
    re += a.re * b.re + a.im * b.im;
    im += a.re * b.im - a.im * b.re;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  re = warp_sum(re);
  im = warp_sum(im);
  if (lane == 0) {
    sred[0][warp] = re;
    sred[1][warp] = im;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += sred[threadIdx.x][w];
    partials[(size_t)blockIdx.x * 2 + threadIdx.x] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(arrive, 1ULL) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (warp < 2) {
      double s = 0.0;
      for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(&partials[(size_t)b * 2 + warp]);
      s = warp_sum(s);
      if (lane == 0) out[warp] = s;
    }
    if (threadIdx.x == 0) *arrive = 0ULL;
  }
}
static int cew_grid(b2o_ctx *c, int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)c->num_sms * 8)); }
static int check_c128(b2o_ctx *c, const void *a, const void *b, const void *r) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)r) % 16) B2O_FAIL(B2O_EARG, "ComplexF64 vectors must be 16-byte aligned");
  return B2O_OK;
}
template <int OP>
static int cew_launch(b2o_ctx *c, CewArgs &p) {
  if (p.nrow <= 0) return B2O_OK;
  B2O_CUDA(cudaSetDevice(c->device));
  cew_kernel<OP><<<cew_grid(c, p.nrow), 256, 0, c->stream>>>(p);
  c->launches++;
  B2O_CUDA(cudaGetLastError());
  return B2O_OK;
}
// mulSquareOpDiagonal! / mulOpDiagonal! for ComplexF64; conj_d != 0: the ctprod! closure (conj.(d))
extern "C" int b2o_cdiag_apply(b2o_ctx *c, int64_t nrow, int64_t ncol, const void *d, int64_t d_len, int conj_d, void *res,
                               int64_t res_len, const void *v, int64_t v_len, double alpha_re, double alpha_im, double beta_re,
                               double beta_im) {
  B2O_TRY(check_c128(c, d, v, res));
  if (nrow < 0 || ncol < 0) B2O_FAIL(B2O_EARG, "negative size");
  if (v_len != ncol || res_len != nrow) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  const int64_t nmin = std::min(nrow, ncol);
  if (d_len < nmin) B2O_FAIL(B2O_EARG, "diagonal shorter than min(nrow,ncol)");
  if (nrow > 0 && (!res || (nmin > 0 && (!d || !v)))) B2O_FAIL(B2O_EARG, "null vector");
  CewArgs p;
  memset(&p, 0, sizeof(p));
  p.a = (const double *)d; p.v = (const double *)v; p.res = (double *)res;
  p.nmin = nmin; p.nrow = nrow;
  p.alpha = cplx{alpha_re, alpha_im}; p.beta = cplx{beta_re, beta_im};
  p.beta_nz = (beta_re != 0.0 || beta_im != 0.0);
  p.conj_a = conj_d != 0;
  return cew_launch<CEW_DIAG>(c, p);
}
// mulOpEye! for ComplexF64 (rectangular tail: 0 when β == 0, else the scalar β -- quirk Q2)
extern "C" int b2o_ceye_apply(b2o_ctx *c, int64_t nrow, int64_t ncol, void *res, int64_t res_len, const void *v, int64_t v_len,
                              double alpha_re, double alpha_im, double beta_re, double beta_im) {
  B2O_TRY(check_c128(c, nullptr, v, res));
  if (v_len != ncol || res_len != nrow) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  CewArgs p;
  memset(&p, 0, sizeof(p));
  p.v = (const double *)v; p.res = (double *)res;
  p.nmin = std::min(nrow, ncol); p.nrow = nrow;
  p.alpha = cplx{alpha_re, alpha_im}; p.beta = cplx{beta_re, beta_im};
  p.beta_nz = (beta_re != 0.0 || beta_im != 0.0);
  p.tail = p.beta_nz ? p.beta : cplx{0.0, 0.0};
  return cew_launch<CEW_EYE>(c, p);
}
// mulOpZeros! for ComplexF64
extern "C" int b2o_czeros_apply(b2o_ctx *c, int64_t nrow, int64_t ncol, void *res, int64_t res_len, int64_t v_len, double beta_re,
                                double beta_im) {
  B2O_TRY(check_c128(c, nullptr, nullptr, res));
  if (v_len != ncol || res_len != nrow) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  CewArgs p;
  memset(&p, 0, sizeof(p));
  p.res = (double *)res;
  p.nmin = nrow; p.nrow = nrow;
  p.beta = cplx{beta_re, beta_im};
  p.beta_nz = (beta_re != 0.0 || beta_im != 0.0);
  return cew_launch<CEW_ZEROS>(c, p);
}
// conj!(res) (dst == src) / conj.(v): the two halves of the conj-sandwich, src/adjtrans.jl:128-136
extern "C" int b2o_conj(b2o_ctx *c, void *dst, const void *src, int64_t n) {
  B2O_TRY(check_c128(c, nullptr, src, dst));
  if (n > 0 && (!dst || !src)) B2O_FAIL(B2O_EARG, "null vector");
  CewArgs p;
  memset(&p, 0, sizeof(p));
  p.v = (const double *)src; p.res = (double *)dst;
  p.nmin = n; p.nrow = n;
  return cew_launch<CEW_CONJ>(c, p);
}
// mulHouseholder! for ComplexF64: res = α (v - 2 dot(h, v) h) (+ β res), dot conjugating h
extern "C" int b2o_chouseholder_apply(b2o_ctx *c, int64_t n, const void *h, void *res, int64_t res_len, const void *v, int64_t v_len,
                                      double alpha_re, double alpha_im, double beta_re, double beta_im) {
  B2O_TRY(check_c128(c, h, v, res));
  if (v_len != n || res_len != n) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if (n == 0) return B2O_OK;
  if (!h || !res || !v) B2O_FAIL(B2O_EARG, "null vector");
  if (c->nranks > 1) B2O_FAIL(B2O_EUNSUPPORTED, "complex Householder is not row-partitioned");
  B2O_CUDA(cudaSetDevice(c->device));
  const int grid = std::min(cew_grid(c, n), B2O_MAX_GRID);
  cdot_kernel<<<grid, 256, 0, c->stream>>>((const double *)h, (const double *)v, n, c->d_partials, c->d_dots + 336, c->d_bar + 1);
  c->launches++;
  B2O_CUDA(cudaGetLastError());
  CewArgs p;
  memset(&p, 0, sizeof(p));
  p.a = (const double *)h; p.v = (const double *)v; p.res = (double *)res;
  p.nmin = n; p.nrow = n;
  p.alpha = cplx{alpha_re, alpha_im}; p.beta = cplx{beta_re, beta_im};
  p.beta_nz = (beta_re != 0.0 || beta_im != 0.0);
  p.dscal = c->d_dots + 336;
  return cew_launch<CEW_HOUSE>(c, p);
}

// ------------------------------------------------------------------ opRestriction / opExtension
struct b2o_index_s {
  b2o_ctx *ctx;
  int64_t k, ncol;
  int64_t *d_idx0;    // [k] 0-based source rows (gather)
  int64_t *d_wpos;    // [kw] positions in u that win their target (scatter; duplicates: last occurrence wins)
  int64_t kw;
  // gather form of the extension (dense index sets): d_inv[j] = position in u that lands on row j, -1 when none.  Built on
  // the host at create time (index work, exact; last occurrence wins like the scatter).  One streaming pass writes res
  // instead of memset + random 8-byte read-modify-writes.
  int32_t *d_inv;     // [ncol] or nullptr
};

extern "C" int b2o_index_create(b2o_ctx *c, const int64_t *idx1, int64_t k, int64_t ncol, b2o_index **out) {
  if (!c || !out || (k > 0 && !idx1)) B2O_FAIL(B2O_EARG, "null argument");
  if (k < 0 || ncol < 0) B2O_FAIL(B2O_EARG, "negative size");
  for (int64_t i = 0; i < k; ++i)
    if (idx1[i] < 1 || idx1[i] > ncol)
      B2O_FAIL(B2O_EARG, "indices should be between 1 and %lld", (long long)ncol);   // special-operators.jl:188
  B2O_CUDA(cudaSetDevice(c->device));
  std::vector<int64_t> idx0(k), wpos;
  for (int64_t i = 0; i < k; ++i) idx0[i] = idx1[i] - 1;
  // winners of duplicated targets (last occurrence).  Dense index sets go through the inverse map (which the gather form of
  // the extension keeps anyway), sparse ones through a hash set so that nothing of size ncol is built for them.
  std::vector<int32_t> inv;
  const bool dense_set = k > 0 && k >= ncol / 16 && k < 0x7fffffffLL;
  if (dense_set) {
    inv.assign((size_t)ncol, -1);
    for (int64_t i = 0; i < k; ++i) inv[(size_t)idx0[i]] = (int32_t)i;          // later occurrences overwrite earlier ones
    for (int64_t i = 0; i < k; ++i)
      if (inv[(size_t)idx0[i]] == (int32_t)i) wpos.push_back(i);
  } else {
    std::unordered_set<int64_t> seen;
    seen.reserve((size_t)k * 2 + 1);
    for (int64_t i = k - 1; i >= 0; --i)
      if (seen.insert(idx0[i]).second) wpos.push_back(i);
    std::reverse(wpos.begin(), wpos.end());
  }
  b2o_index *ix = new b2o_index_s();
  ix->ctx = c;
  ix->k = k;
  ix->ncol = ncol;
  ix->kw = (int64_t)wpos.size();
  ix->d_idx0 = ix->d_wpos = nullptr;
  ix->d_inv = nullptr;
  cudaError_t e1 = cudaMalloc(&ix->d_idx0, sizeof(int64_t) * std::max<int64_t>(k, 1));
  cudaError_t e2 = cudaMalloc(&ix->d_wpos, sizeof(int64_t) * std::max<int64_t>(ix->kw, 1));
  // the inverse map costs 4 bytes per row of the long vector: worth it when at least one row in 16 is hit (then
  // 4 n + 8 n + 32 kw bytes in one streaming pass beat 8 n of memset + ~88 kw of sector-granular read-modify-writes)
  const bool want_inv = dense_set && ix->kw >= ncol / 16;
  cudaError_t e3 = want_inv ? cudaMalloc(&ix->d_inv, sizeof(int32_t) * (size_t)ncol) : cudaSuccess;
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    cudaGetLastError();
    cudaFree(ix->d_idx0);
    cudaFree(ix->d_wpos);
    cudaFree(ix->d_inv);
    delete ix;
    B2O_FAIL(B2O_ENOMEM, "index allocation failed");
  }
  if (e3 != cudaSuccess) {          // no room for the inverse map: the scatter form still works
    cudaGetLastError();
    ix->d_inv = nullptr;
  }
  if (k > 0) {
    B2O_CUDA(cudaMemcpy(ix->d_idx0, idx0.data(), sizeof(int64_t) * k, cudaMemcpyHostToDevice));
    B2O_CUDA(cudaMemcpy(ix->d_wpos, wpos.data(), sizeof(int64_t) * ix->kw, cudaMemcpyHostToDevice));
  }
  if (ix->d_inv) {
    B2O_CUDA(cudaMemcpy(ix->d_inv, inv.data(), sizeof(int32_t) * (size_t)ncol, cudaMemcpyHostToDevice));
  }
  *out = ix;
  return B2O_OK;
}
extern "C" int b2o_index_destroy(b2o_index *ix) {
  if (!ix) return B2O_OK;
  cudaSetDevice(ix->ctx->device);
  cudaStreamSynchronize(ix->ctx->stream);
  cudaFree(ix->d_idx0);
  cudaFree(ix->d_wpos);
  cudaFree(ix->d_inv);
  delete ix;
  return B2O_OK;
}

template <typename E>
__global__ void __launch_bounds__(256) gather_kernel(E *__restrict__ res, const E *__restrict__ v,
                                                     const int64_t *__restrict__ idx0, int64_t k) {
  const int64_t base = (int64_t)blockIdx.x * (256 * 4) + threadIdx.x;
  int64_t src[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t i = base + u * 256;
    src[u] = (i < k) ? __ldcs(&idx0[i]) : -1;
  }
  E val[4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (src[u] >= 0) val[u] = __ldg(&v[src[u]]);
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int64_t i = base + u * 256;
    if (i < k) __stcs(&res[i], val[u]);
  }
}
template <typename E>
__global__ void __launch_bounds__(256) scatter_kernel(E *__restrict__ res, const E *__restrict__ u,
                                                      const int64_t *__restrict__ idx0,
                                                      const int64_t *__restrict__ wpos, int64_t kw) {
  const int64_t base = (int64_t)blockIdx.x * (256 * 4) + threadIdx.x;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t w = base + q * 256;
    if (w < kw) {
      const int64_t i = __ldcs(&wpos[w]);
      res[__ldcs(&idx0[i])] = __ldcs(&u[i]);
    }
  }
}

// extension in gather form: res[j] = inv[j] >= 0 ? u[inv[j]] : 0 -- one streaming pass over res and inv, gathers from u
template <typename E>
__global__ void __launch_bounds__(256) extend_gather_kernel(E *__restrict__ res, const E *__restrict__ u,
                                                            const int32_t *__restrict__ inv, int64_t ncol) {
  const int64_t base = (int64_t)blockIdx.x * (256 * 4) + threadIdx.x;
  int32_t src[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t j = base + q * 256;
    src[q] = (j < ncol) ? __ldcs(&inv[j]) : -1;
  }
  E val[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) val[q] = (src[q] >= 0) ? __ldg(&u[src[q]]) : (E)0;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int64_t j = base + q * 256;
    if (j < ncol) __stcs(&res[j], val[q]);
  }
}

static int elem_size(int dtype) { return dtype == B2O_F64 ? 8 : dtype == B2O_F32 ? 4 : dtype == B2O_BF16 ? 2 : 0; }

extern "C" int b2o_restrict_apply(b2o_index *ix, int dtype, void *res, int64_t res_len, const void *v, int64_t v_len) {
  if (!ix) B2O_FAIL(B2O_EARG, "null index");
  if (v_len != ix->ncol || res_len != ix->k) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  const int es = elem_size(dtype);
  if (es != 8 && es != 4) B2O_FAIL(B2O_EUNSUPPORTED, "restriction supports 8- and 4-byte elements");
  if (ix->k == 0) return B2O_OK;
  if (!res || !v) B2O_FAIL(B2O_EARG, "null vector");
  b2o_ctx *c = ix->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  const unsigned grid = (unsigned)((ix->k + 1023) / 1024);
  if (es == 8)
    gather_kernel<double><<<grid, 256, 0, c->stream>>>((double *)res, (const double *)v, ix->d_idx0, ix->k);
  else
    gather_kernel<float><<<grid, 256, 0, c->stream>>>((float *)res, (const float *)v, ix->d_idx0, ix->k);
  c->launches++;
  B2O_CUDA(cudaGetLastError());
  return B2O_OK;
}

extern "C" int b2o_extend_apply(b2o_index *ix, int dtype, void *res, int64_t res_len, const void *u, int64_t u_len) {
  if (!ix) B2O_FAIL(B2O_EARG, "null index");
  if (u_len != ix->k || res_len != ix->ncol) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  const int es = elem_size(dtype);
  if (es != 8 && es != 4) B2O_FAIL(B2O_EUNSUPPORTED, "extension supports 8- and 4-byte elements");
  if (ix->ncol == 0) return B2O_OK;
  if (!res || (ix->k > 0 && !u)) B2O_FAIL(B2O_EARG, "null vector");
  b2o_ctx *c = ix->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  if (ix->d_inv && c->extend_form != 1) {                                       // dense index set: one gather-form pass
    const unsigned grid = (unsigned)((ix->ncol + 1023) / 1024);
    if (es == 8)
      extend_gather_kernel<double><<<grid, 256, 0, c->stream>>>((double *)res, (const double *)u, ix->d_inv, ix->ncol);
    else
      extend_gather_kernel<float><<<grid, 256, 0, c->stream>>>((float *)res, (const float *)u, ix->d_inv, ix->ncol);
    c->launches++;
    B2O_CUDA(cudaGetLastError());
    return B2O_OK;
  }
  B2O_CUDA(cudaMemsetAsync(res, 0, (size_t)ix->ncol * es, c->stream));          // res .= 0   special-operators.jl:172
  if (ix->kw > 0) {
    const unsigned grid = (unsigned)((ix->kw + 1023) / 1024);
    if (es == 8)
      scatter_kernel<double><<<grid, 256, 0, c->stream>>>((double *)res, (const double *)u, ix->d_idx0, ix->d_wpos, ix->kw);
    else
      scatter_kernel<float><<<grid, 256, 0, c->stream>>>((float *)res, (const float *)u, ix->d_idx0, ix->d_wpos, ix->kw);
    c->launches++;
    B2O_CUDA(cudaGetLastError());
  }
  return B2O_OK;
}

// ------------------------------------------------------------------ diagonal quasi-Newton updates (src/DiagonalHessianApproximation.jl)
// One pass computes every reduction the four push! variants need: Σs², Σs⁴, Σs·y, Σs²·d, Σ|y|, count(s != 0).
struct DiagQnRedArgs {
  const double *s, *y, *d;
  int64_t n;
  int d_is_scalar;
  double *partials;            // [grid][6]
  double *out;                 // [6]
  unsigned long long *arrive;
};
__global__ void __launch_bounds__(256) diagqn_reduce_kernel(const __grid_constant__ DiagQnRedArgs a) {
  __shared__ double sred[6][8];
  __shared__ bool is_last;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    const double s = a.s[i], y = a.y[i], s2 = s * s;
    acc[0] += s2;
    acc[1] = fma(s2, s2, acc[1]);
    acc[2] = fma(s, y, acc[2]);
    if (!a.d_is_scalar) acc[3] = fma(s2, a.d[i], acc[3]);
    acc[4] += fabs(y);
    acc[5] += (s != 0.0) ? 1.0 : 0.0;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < 6; ++p) {
    double v = warp_sum(acc[p]);
    if (lane == 0) sred[p][warp] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += sred[threadIdx.x][w];
    a.partials[(size_t)blockIdx.x * 6 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(a.arrive, 1ULL) == gridDim.x - 1);
  __syncthreads();
  if (is_last) {
    __threadfence();
    if (warp < 6) {
      double v = 0.0;
      for (int b = lane; b < (int)gridDim.x; b += 32) v += __ldcg(&a.partials[(size_t)b * 6 + warp]);
      v = warp_sum(v);
      if (lane == 0) a.out[warp] = v;
    }
    if (threadIdx.x == 0) *a.arrive = 0ULL;
  }
}
// kind 0 PSB: d += c*s² ; 1 Andrei: d += (c*s² - 1) ; 2 BFGS: d = |y| * c
__global__ void diagqn_update_kernel(double *d, const double *s, const double *y, double c, int kind, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (kind == 0) d[i] = d[i] + c * (s[i] * s[i]);
    else if (kind == 1) d[i] = d[i] + (c * (s[i] * s[i]) - 1.0);
    else d[i] = fabs(y[i]) * c;
  }
}

// push!(B, s, y) for DiagonalPSB (kind 0, :45-64), DiagonalAndrei (1, :120-141), DiagonalBFGS (2, :234-248) and
// SpectralGradient (3, :186-196; d is the 1-element device vector holding σ).  s == 0 -> B2O_ESTATE (the reference errors).
extern "C" int b2o_diagqn_push(b2o_ctx *c, int kind, void *d_, int64_t d_len, const void *s_, const void *y_, int64_t n) {
  if (!c) B2O_FAIL(B2O_EARG, "null context");
  if (kind < 0 || kind > 3) B2O_FAIL(B2O_EARG, "bad diagonal quasi-Newton kind %d", kind);
  if ((kind == 3 && d_len != 1) || (kind != 3 && d_len != n)) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if (n > 0 && (!d_ || !s_ || !y_)) B2O_FAIL(B2O_EARG, "null vector");
  B2O_TRY(check_ptrs8(d_, s_, y_));
  B2O_CUDA(cudaSetDevice(c->device));
  double *d = (double *)d_;
  const double *s = (const double *)s_, *y = (const double *)y_;
  DiagQnRedArgs a;
  a.s = s; a.y = y; a.d = d; a.n = n;
  a.d_is_scalar = kind == 3;
  a.partials = c->d_partials;
  a.out = c->d_dots + 330;
  a.arrive = c->d_bar + 1;
  const int64_t want = (n + 255) / 256;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)c->num_sms * 4));
  diagqn_reduce_kernel<<<grid, 256, 0, c->stream>>>(a);
  c->launches++;
  B2O_CUDA(cudaGetLastError());
  B2O_TRY(b2o_allreduce_sum_f64(c, c->d_dots + 330, 6));
  double h[6];
  B2O_TRY(b2o_read_scalars(c, c->d_dots + 330, 6, h));
  const double ss = h[0], s4 = h[1], sy = h[2], s2d = h[3], sumabsy = h[4], nnz = h[5];
  if (kind == 3) {
    if (nnz == 0) B2O_FAIL(B2O_ESTATE, "Cannot divide by zero and s .= 0");
    const double sigma = sy / ss;                                             // B.d[1] = dot(s,y)/dot(s,s)
    B2O_CUDA(cudaMemcpyAsync(d, &sigma, sizeof(double), cudaMemcpyHostToDevice, c->stream));
    B2O_CUDA(cudaStreamSynchronize(c->stream));
    return B2O_OK;
  }
  const double sNorm = sqrt(ss);
  if (sNorm == 0) B2O_FAIL(B2O_ESTATE, "Cannot update DiagonalQN operator with s=0");
  const double sNorm2 = sNorm * sNorm;
  double coef;
  if (kind == 2) {
    const double sT_y = sy / sNorm2;
    coef = sumabsy / sT_y;                                                    // d .= abs.(y); d .*= sum(d)/sT_y
  } else {
    const double trA2 = s4 / (sNorm2 * sNorm2);
    const double sT_y = sy / sNorm2, sT_B_s = s2d / sNorm2;
    double q = sT_y - sT_B_s;
    if (kind == 1) q += ss / sNorm2;                                          // sT_s = dot(s,s)/sNorm2
    q /= trA2;
    coef = q / sNorm2;
  }
  if (n > 0) {
    diagqn_update_kernel<<<grid, 256, 0, c->stream>>>(d, s, y, coef, kind, n);
    c->launches++;
    B2O_CUDA(cudaGetLastError());
  }
  return B2O_OK;
}
