// b2o_dense_kernels.cuh -- kernels + launch logic of the dense-matrix leaf (see b2o_dense.cu for the description).
// Kept free of CUDA runtime calls so that tests/emu/ can compile the SAME code for the host under a small SIMT emulator
// (B2O_SIMT_EMU: one OS thread per CUDA thread, barriers for __syncthreads / warp shuffles) and check the index logic of
// every path (vectorised / scalar, split / unsplit, ragged edges) on a CPU-only box.
#pragma once
#include <algorithm>

constexpr int DN_THREADS = 256;
constexpr int DN_UNROLL = 8;   // columns in flight per thread (N kernel)
constexpr int DT_CB = 8;       // columns per CTA (T kernel)
constexpr int DENSE_MIN_COLS = 64;      // never split below this many columns (N) ...
constexpr int DENSE_MIN_ROWITERS = 4;   // ... or this many row sweeps of a CTA (T)

struct DenseArgs {
  const void *M, *v;
  void *res;
  int64_t m, n, lda;
  int64_t chunk;   // columns (N) / rows (T) per split
  int nsplit;
  double alpha, beta;
  double *part;    // [nsplit][len(res)] when nsplit > 1
  int tx_log2;     // narrow kernels: log2 of the threads laid along the rows
};

// 16-byte read-only load that does not allocate in L1 (the matrix is streamed; L2 keeps it when it fits)
__device__ __forceinline__ uint4 ldg_nc16(const void *p) {
  uint4 r;
#ifdef B2O_SIMT_EMU
  memcpy(&r, p, 16);
#else
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#endif
  return r;
}
// W consecutive elements of type T as they come out of memory: one 16-byte register quad (W*sizeof(T) == 16) or one scalar.
// Kept raw while in flight (a Float32 quad converted to doubles would occupy twice the registers) and unpacked at use.
template <typename T, int W>
struct Raw {
  uint4 q;
};
template <typename T>
struct Raw<T, 1> {
  T q;
};
template <typename T, int W>
__device__ __forceinline__ Raw<T, W> load_raw(const T *p) {
  Raw<T, W> r;
  if constexpr (W == 1)
    r.q = __ldg(p);
  else
    r.q = ldg_nc16(p);
  return r;
}
template <typename T, int W>
__device__ __forceinline__ void unpack(const Raw<T, W> &r, double (&e)[W]) {
  if constexpr (W == 1) {
    e[0] = (double)r.q;
  } else if constexpr (sizeof(T) == 8) {
    static_assert(sizeof(T) != 8 || W == 2, "two doubles per 16-byte load");
    e[0] = __hiloint2double((int)r.q.y, (int)r.q.x);
    e[1] = __hiloint2double((int)r.q.w, (int)r.q.z);
  } else {
    static_assert(sizeof(T) == 8 || W == 4, "four floats per 16-byte load");
    e[0] = (double)__uint_as_float(r.q.x);
    e[1] = (double)__uint_as_float(r.q.y);
    e[2] = (double)__uint_as_float(r.q.z);
    e[3] = (double)__uint_as_float(r.q.w);
  }
}

template <typename T>
__device__ __forceinline__ void dense_epilogue(const DenseArgs &p, T *res, int64_t i, double acc) {
  double t = p.alpha * acc;
  if (p.beta != 0.0) t += p.beta * (double)res[i];
  res[i] = (T)t;
}

// ------------------------------------------------------------------ res = α M v + β res   (M column-major m x n)
// Register double-buffering: the DN_UNROLL loads of the NEXT column group are issued before the current group is consumed,
// so DN_UNROLL 16-byte loads per thread are in flight whatever the instruction scheduler does inside a group.
template <typename T, bool VEC>
__global__ void __launch_bounds__(DN_THREADS, 2) dense_n_kernel(const __grid_constant__ DenseArgs p) {
  constexpr int W = VEC ? (int)(16 / sizeof(T)) : 1;
  const T *__restrict__ M = (const T *)p.M;
  const T *__restrict__ v = (const T *)p.v;
  const int64_t row0 = ((int64_t)blockIdx.x * DN_THREADS + threadIdx.x) * W;
  if (row0 >= p.m) return;
  const int64_t j0 = (int64_t)blockIdx.y * p.chunk;
  const int64_t j1 = min(p.n, j0 + p.chunk);
  double acc[W];
#pragma unroll
  for (int w = 0; w < W; ++w) acc[w] = 0.0;
  const int nrows = (int)min((int64_t)W, p.m - row0);
  if (nrows == W) {
    const T *col = M + row0 + j0 * p.lda;
    const int64_t ngroups = (j1 - j0) / DN_UNROLL;
    Raw<T, W> cur[DN_UNROLL], nxt[DN_UNROLL];
    T xc[DN_UNROLL], xn[DN_UNROLL];
    if (ngroups > 0) {
#pragma unroll
      for (int u = 0; u < DN_UNROLL; ++u) nxt[u] = load_raw<T, W>(col + (int64_t)u * p.lda);
#pragma unroll
      for (int u = 0; u < DN_UNROLL; ++u) xn[u] = __ldg(v + j0 + u);
    }
    int64_t j = j0;
    for (int64_t g = 0; g < ngroups; ++g) {
#pragma unroll
      for (int u = 0; u < DN_UNROLL; ++u) {
        cur[u] = nxt[u];
        xc[u] = xn[u];
      }
      col += (int64_t)DN_UNROLL * p.lda;
      j += DN_UNROLL;
      if (g + 1 < ngroups) {
#pragma unroll
        for (int u = 0; u < DN_UNROLL; ++u) nxt[u] = load_raw<T, W>(col + (int64_t)u * p.lda);
#pragma unroll
        for (int u = 0; u < DN_UNROLL; ++u) xn[u] = __ldg(v + j + u);
      }
#pragma unroll
      for (int u = 0; u < DN_UNROLL; ++u) {
        double e[W];
        unpack<T, W>(cur[u], e);
        const double x = (double)xc[u];
#pragma unroll
        for (int w = 0; w < W; ++w) acc[w] = fma(e[w], x, acc[w]);
      }
    }
    for (; j < j1; ++j, col += p.lda) {             // remaining (< DN_UNROLL) columns of the split
      double e[W];
      unpack<T, W>(load_raw<T, W>(col), e);
      const double x = (double)__ldg(v + j);
#pragma unroll
      for (int w = 0; w < W; ++w) acc[w] = fma(e[w], x, acc[w]);
    }
  } else {
    // ragged last rows of a vectorised launch (m not a multiple of W)
#pragma unroll
    for (int w = 0; w < W; ++w) {
      if (w < nrows) {
        const T *col = M + row0 + w + j0 * p.lda;
        double s = 0.0;
        for (int64_t j = j0; j < j1; ++j, col += p.lda) s = fma((double)__ldg(col), (double)__ldg(v + j), s);
        acc[w] = s;
      }
    }
  }
#pragma unroll
  for (int w = 0; w < W; ++w) {
    if (w < nrows) {
      if (p.nsplit > 1)
        p.part[(int64_t)blockIdx.y * p.m + row0 + w] = acc[w];
      else
        dense_epilogue<T>(p, (T *)p.res, row0 + w, acc[w]);
    }
  }
}

// ------------------------------------------------------------------ res = α Mᵀ u + β res   (M column-major m x n)
template <typename T, bool VEC>
__global__ void __launch_bounds__(DN_THREADS, 2) dense_t_kernel(const __grid_constant__ DenseArgs p) {
  constexpr int W = VEC ? (int)(16 / sizeof(T)) : 1;
  __shared__ double sred[DN_THREADS / 32][DT_CB];
  const T *__restrict__ M = (const T *)p.M;
  const T *__restrict__ u = (const T *)p.v;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t c0 = (int64_t)blockIdx.x * DT_CB;
  const int nc = (int)min((int64_t)DT_CB, p.n - c0);
  const int64_t i0 = (int64_t)blockIdx.y * p.chunk;          // chunk is a multiple of DN_THREADS * W
  const int64_t i1 = min(p.m, i0 + p.chunk);
  const int64_t iv = i0 + ((i1 - i0) / W) * W;               // end of the vectorised rows
  const int64_t step = (int64_t)DN_THREADS * W;
  const T *colbase = M + c0 * p.lda;
  double acc[DT_CB];
#pragma unroll
  for (int c = 0; c < DT_CB; ++c) acc[c] = 0.0;
  if (nc == DT_CB) {
    // full column group, register double-buffered: the u quad + DT_CB matrix quads of the NEXT row sweep are issued before
    // the current sweep is consumed (9 independent 16-byte loads in flight per thread)
    int64_t i = i0 + (int64_t)threadIdx.x * W;
    Raw<T, W> cur[DT_CB], nxt[DT_CB], xc, xn;
    if (i < iv) {
      xn = load_raw<T, W>(u + i);
#pragma unroll
      for (int c = 0; c < DT_CB; ++c) nxt[c] = load_raw<T, W>(colbase + (int64_t)c * p.lda + i);
    }
    while (i < iv) {
      xc = xn;
#pragma unroll
      for (int c = 0; c < DT_CB; ++c) cur[c] = nxt[c];
      i += step;
      if (i < iv) {
        xn = load_raw<T, W>(u + i);
#pragma unroll
        for (int c = 0; c < DT_CB; ++c) nxt[c] = load_raw<T, W>(colbase + (int64_t)c * p.lda + i);
      }
      double x[W];
      unpack<T, W>(xc, x);
#pragma unroll
      for (int c = 0; c < DT_CB; ++c) {
        double e[W];
        unpack<T, W>(cur[c], e);
#pragma unroll
        for (int w = 0; w < W; ++w) acc[c] = fma(e[w], x[w], acc[c]);
      }
    }
  } else {
    // last, partial column group
    for (int64_t i = i0 + (int64_t)threadIdx.x * W; i < iv; i += step) {
      double x[W];
      unpack<T, W>(load_raw<T, W>(u + i), x);
#pragma unroll
      for (int c = 0; c < DT_CB; ++c) {
        if (c < nc) {
          double e[W];
          unpack<T, W>(load_raw<T, W>(colbase + (int64_t)c * p.lda + i), e);
#pragma unroll
          for (int w = 0; w < W; ++w) acc[c] = fma(e[w], x[w], acc[c]);
        }
      }
    }
  }
  // ragged rows at the end of the last slab (only when m is not a multiple of W)
  for (int64_t i = iv + threadIdx.x; i < i1; i += DN_THREADS) {
    const double x = (double)__ldg(u + i);
#pragma unroll
    for (int c = 0; c < DT_CB; ++c)
      if (c < nc) acc[c] = fma((double)__ldg(colbase + (int64_t)c * p.lda + i), x, acc[c]);
  }
#pragma unroll
  for (int c = 0; c < DT_CB; ++c) {
    const double s = warp_sum(acc[c]);
    if (lane == 0) sred[warp][c] = s;
  }
  __syncthreads();
  if (threadIdx.x < nc) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < DN_THREADS / 32; ++w) s += sred[w][threadIdx.x];
    if (p.nsplit > 1)
      p.part[(int64_t)blockIdx.y * p.n + c0 + threadIdx.x] = s;
    else
      dense_epilogue<T>(p, (T *)p.res, c0 + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------ matrices with few rows (nrow <= 128 quads)
// With a short leading dimension the one-thread-per-row-quad layout above leaves most of a CTA idle.  The narrow kernels lay
// the 256 threads out as TX x TY: TX = 2^tx_log2 threads cover the row quads of one column, TY = 256/TX columns are worked
// on at the same time (for a contiguous matrix a warp then reads one contiguous 512-byte run).
//
// res[m] = α M v + β res: the CTA owns the columns [j0, j1) of its split; thread (tx, ty) accumulates its row quad over the
// columns j0+ty, j0+ty+TY, ...; the TY partial sums per row are added in ty order through shared memory.
template <typename T, bool VEC>
__global__ void __launch_bounds__(DN_THREADS, 2) dense_n_narrow_kernel(const __grid_constant__ DenseArgs p) {
  constexpr int W = VEC ? (int)(16 / sizeof(T)) : 1;
  __shared__ double red[DN_THREADS * W];                     // [ty][TX * W]
  const T *__restrict__ M = (const T *)p.M;
  const T *__restrict__ v = (const T *)p.v;
  const int TX = 1 << p.tx_log2, TY = DN_THREADS >> p.tx_log2;
  const int tx = threadIdx.x & (TX - 1), ty = threadIdx.x >> p.tx_log2;
  const int64_t row0 = (int64_t)tx * W;
  const int nrows = row0 < p.m ? (int)min((int64_t)W, p.m - row0) : 0;
  const int64_t j0 = (int64_t)blockIdx.x * p.chunk;
  const int64_t j1 = min(p.n, j0 + p.chunk);
  const int64_t mine = (j0 + ty < j1) ? (j1 - j0 - ty + TY - 1) / TY : 0;   // columns of this thread
  double acc[W];
#pragma unroll
  for (int w = 0; w < W; ++w) acc[w] = 0.0;
  if (nrows == W) {
    const int64_t cstride = (int64_t)TY * p.lda;
    const T *col = M + row0 + (j0 + ty) * p.lda;
    const T *vj = v + j0 + ty;
    const int64_t ngroups = mine / DN_UNROLL;
    Raw<T, W> cur[DN_UNROLL], nxt[DN_UNROLL];
    T xc[DN_UNROLL], xn[DN_UNROLL];
    if (ngroups > 0) {
#pragma unroll
      for (int u = 0; u < DN_UNROLL; ++u) nxt[u] = load_raw<T, W>(col + (int64_t)u * cstride);
#pragma unroll
      for (int u = 0; u < DN_UNROLL; ++u) xn[u] = __ldg(vj + (int64_t)u * TY);
    }
    int64_t k = 0;
    for (int64_t g = 0; g < ngroups; ++g) {
#pragma unroll
      for (int u = 0; u < DN_UNROLL; ++u) {
        cur[u] = nxt[u];
        xc[u] = xn[u];
      }
      col += (int64_t)DN_UNROLL * cstride;
      vj += (int64_t)DN_UNROLL * TY;
      k += DN_UNROLL;
      if (g + 1 < ngroups) {
#pragma unroll
        for (int u = 0; u < DN_UNROLL; ++u) nxt[u] = load_raw<T, W>(col + (int64_t)u * cstride);
#pragma unroll
        for (int u = 0; u < DN_UNROLL; ++u) xn[u] = __ldg(vj + (int64_t)u * TY);
      }
#pragma unroll
      for (int u = 0; u < DN_UNROLL; ++u) {
        double e[W];
        unpack<T, W>(cur[u], e);
        const double x = (double)xc[u];
#pragma unroll
        for (int w = 0; w < W; ++w) acc[w] = fma(e[w], x, acc[w]);
      }
    }
    for (; k < mine; ++k, col += cstride, vj += TY) {
      double e[W];
      unpack<T, W>(load_raw<T, W>(col), e);
      const double x = (double)__ldg(vj);
#pragma unroll
      for (int w = 0; w < W; ++w) acc[w] = fma(e[w], x, acc[w]);
    }
  } else if (nrows > 0) {
    // ragged last row quad (m not a multiple of W)
#pragma unroll
    for (int w = 0; w < W; ++w) {
      if (w < nrows) {
        double s = 0.0;
        for (int64_t j = j0 + ty; j < j1; j += TY) s = fma((double)__ldg(M + row0 + w + j * p.lda), (double)__ldg(v + j), s);
        acc[w] = s;
      }
    }
  }
#pragma unroll
  for (int w = 0; w < W; ++w) red[(ty * TX + tx) * W + w] = acc[w];
  __syncthreads();
  for (int64_t r = threadIdx.x; r < p.m; r += DN_THREADS) {
    double s = 0.0;
    for (int y = 0; y < TY; ++y) s += red[(int64_t)y * TX * W + r];
    if (p.nsplit > 1)
      p.part[(int64_t)blockIdx.x * p.m + r] = s;
    else
      dense_epilogue<T>(p, (T *)p.res, r, s);
  }
}

// res[n] = α Mᵀ u + β res with at most 32 row quads: the u quad of a thread stays in registers for the whole kernel, thread
// (tx, ty) forms its part of the dot for the columns c0+ty, c0+ty+TY, ..., the TX parts are added with a shuffle butterfly
// (TX <= 32 consecutive lanes) and lane tx == 0 writes the result.  No split: the CTA owns the columns [c0, c1).
template <typename T, bool VEC>
__global__ void __launch_bounds__(DN_THREADS, 2) dense_t_narrow_kernel(const __grid_constant__ DenseArgs p) {
  constexpr int W = VEC ? (int)(16 / sizeof(T)) : 1;
  const T *__restrict__ M = (const T *)p.M;
  const T *__restrict__ u = (const T *)p.v;
  const int TX = 1 << p.tx_log2, TY = DN_THREADS >> p.tx_log2;
  const int tx = threadIdx.x & (TX - 1), ty = threadIdx.x >> p.tx_log2;
  const int64_t row0 = (int64_t)tx * W;
  const int nrows = row0 < p.m ? (int)min((int64_t)W, p.m - row0) : 0;
  const bool full = nrows == W;
  double x[W];
#pragma unroll
  for (int w = 0; w < W; ++w) x[w] = 0.0;
  if (full) {
    unpack<T, W>(load_raw<T, W>(u + row0), x);
  } else {
#pragma unroll
    for (int w = 0; w < W; ++w)
      if (w < nrows) x[w] = (double)__ldg(u + row0 + w);
  }
  const int64_t c0 = (int64_t)blockIdx.x * p.chunk;
  const int64_t c1 = min(p.n, c0 + p.chunk);
  const int64_t blockcols = (int64_t)TY * DN_UNROLL;          // columns one iteration of the CTA covers
  const int64_t nfull = (c1 - c0) / blockcols;
  const int64_t cstride = (int64_t)TY * p.lda;
  const T *col = M + row0 + (c0 + ty) * p.lda;
  Raw<T, W> cur[DN_UNROLL], nxt[DN_UNROLL];
  if (full && nfull > 0) {
#pragma unroll
    for (int k = 0; k < DN_UNROLL; ++k) nxt[k] = load_raw<T, W>(col + (int64_t)k * cstride);
  }
  int64_t j = c0 + ty;
  for (int64_t b = 0; b < nfull; ++b) {
    double s[DN_UNROLL];
    if (full) {
#pragma unroll
      for (int k = 0; k < DN_UNROLL; ++k) cur[k] = nxt[k];
      col += (int64_t)DN_UNROLL * cstride;
      if (b + 1 < nfull) {
#pragma unroll
        for (int k = 0; k < DN_UNROLL; ++k) nxt[k] = load_raw<T, W>(col + (int64_t)k * cstride);
      }
#pragma unroll
      for (int k = 0; k < DN_UNROLL; ++k) {
        double e[W];
        unpack<T, W>(cur[k], e);
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w) t = fma(e[w], x[w], t);
        s[k] = t;
      }
    } else {
#pragma unroll
      for (int k = 0; k < DN_UNROLL; ++k) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < W; ++w)
          if (w < nrows) t = fma((double)__ldg(M + row0 + w + (j + (int64_t)k * TY) * p.lda), x[w], t);
        s[k] = t;
      }
    }
#pragma unroll
    for (int k = 0; k < DN_UNROLL; ++k) {
      double t = s[k];
      for (int o = TX >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (tx == 0) dense_epilogue<T>(p, (T *)p.res, j + (int64_t)k * TY, t);
    }
    j += blockcols;
  }
  // remaining (< TY * DN_UNROLL) columns, TY at a time; the trip count is uniform across the CTA (shuffles stay convergent)
  for (int64_t jb = c0 + nfull * blockcols; jb < c1; jb += TY) {
    const int64_t jj = jb + ty;
    double t = 0.0;
    if (jj < c1) {
      if (full) {
        double e[W];
        unpack<T, W>(load_raw<T, W>(M + row0 + jj * p.lda), e);
#pragma unroll
        for (int w = 0; w < W; ++w) t = fma(e[w], x[w], t);
      } else {
#pragma unroll
        for (int w = 0; w < W; ++w)
          if (w < nrows) t = fma((double)__ldg(M + row0 + w + jj * p.lda), x[w], t);
      }
    }
    for (int o = TX >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (tx == 0 && jj < c1) dense_epilogue<T>(p, (T *)p.res, jj, t);
  }
}

// ------------------------------------------------------------------ fixed-order sum of the splits + epilogue
template <typename T>
__global__ void __launch_bounds__(256) dense_finish_kernel(const __grid_constant__ DenseArgs p, int64_t len) {
  const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= len) return;
  double s = 0.0;
  for (int k = 0; k < p.nsplit; ++k) s += __ldcg(p.part + (int64_t)k * len + i);
  dense_epilogue<T>(p, (T *)p.res, i, s);
}

// ------------------------------------------------------------------ host side
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

struct DensePlan {
  int64_t gx = 0, chunk = 0;
  int nsplit = 1;
  int narrow = 0;    // 1: the narrow kernels (few row quads)
  int tx_log2 = 0;
};
constexpr int64_t DENSE_NARROW_N_QUADS = 128;   // narrow N kernel: at least 2 columns in flight per CTA sweep
constexpr int64_t DENSE_NARROW_T_QUADS = 32;    // narrow T kernel: the row reduction stays inside one warp

// Split so that ~4 CTAs per SM exist whenever the matrix is big enough; never below DENSE_MIN_COLS columns /
// DENSE_MIN_ROWITERS row sweeps per split (keeps the partial-sum traffic under a few % of the matrix bytes).
static DensePlan dense_plan(int num_sms, int trans, int64_t m, int64_t n, int W) {
  DensePlan pl;
  const int64_t target = 4 * (int64_t)num_sms;
  const int64_t quads = ceil_div64(m, W);
  if (quads <= (trans ? DENSE_NARROW_T_QUADS : DENSE_NARROW_N_QUADS)) {
    pl.narrow = 1;
    while (((int64_t)1 << pl.tx_log2) < quads) ++pl.tx_log2;
    const int64_t blockcols = (int64_t)(DN_THREADS >> pl.tx_log2) * DN_UNROLL;   // columns per CTA iteration
    // whole iterations per CTA, at least two of them, ~target CTAs when there are enough columns
    pl.chunk = std::max<int64_t>(2 * blockcols, ceil_div64(ceil_div64(n, target), blockcols) * blockcols);
    if (!trans && ceil_div64(n, pl.chunk) > 1024) pl.chunk = ceil_div64(ceil_div64(n, 1024), blockcols) * blockcols;
    pl.gx = std::max<int64_t>(1, ceil_div64(n, pl.chunk));
    pl.nsplit = trans ? 1 : (int)pl.gx;           // N: every CTA is a column split; T: no reduction across CTAs
    return pl;
  }
  if (!trans) {
    pl.gx = std::max<int64_t>(1, ceil_div64(m, (int64_t)DN_THREADS * W));
    int64_t s = std::max<int64_t>(1, std::min<int64_t>(ceil_div64(target, pl.gx), n / DENSE_MIN_COLS));
    s = std::min<int64_t>(s, 1024);
    pl.chunk = std::max<int64_t>(1, ceil_div64(n, s));
    pl.nsplit = (int)std::max<int64_t>(1, ceil_div64(n, pl.chunk));
  } else {
    pl.gx = std::max<int64_t>(1, ceil_div64(n, DT_CB));
    const int64_t sweep = (int64_t)DN_THREADS * W;
    int64_t s = std::max<int64_t>(1, std::min<int64_t>(ceil_div64(target, pl.gx), m / (sweep * DENSE_MIN_ROWITERS)));
    s = std::min<int64_t>(s, 1024);
    pl.chunk = std::max<int64_t>(sweep, ceil_div64(ceil_div64(m, s), sweep) * sweep);
    pl.nsplit = (int)std::max<int64_t>(1, ceil_div64(m, pl.chunk));
  }
  return pl;
}

// partial-sum workspace (doubles) for the worst case over {N, T} x {vectorised, scalar} products of an m x n matrix
static inline size_t dense_workspace_elems(int num_sms, int Wv, int64_t m, int64_t n) {
  size_t need = 0;
  for (int trans = 0; trans < 2; ++trans)
    for (int W : {1, Wv}) {
      const DensePlan pl = dense_plan(num_sms, trans, m, n, W);
      if (pl.nsplit > 1) need = std::max(need, (size_t)pl.nsplit * (size_t)(trans ? n : m));
    }
  return need;
}

// One product on `stream`: picks the vectorised kernels when alignment allows, plans the split, launches.
// `part` / `part_elems`: the handle's partial-sum workspace.  `launches` counts kernel launches (ctx accounting).
template <typename T>
static int dense_run_impl(int num_sms, B2O_STREAM_T stream, int64_t *launches, const void *M, int64_t m, int64_t n, int64_t lda,
                          double *part, size_t part_elems, int trans, void *res, const void *v, double alpha, double beta,
                          int force_scalar) {
  constexpr int Wv = (int)(16 / sizeof(T));
  const int64_t out_len = trans ? n : m;
  if (out_len == 0) return B2O_OK;
  // 16-byte loads need an aligned matrix, a leading dimension that keeps every column aligned and (T kernel) an aligned u
  bool vec = !force_scalar && ((uintptr_t)M % 16) == 0 && (lda % Wv) == 0;
  if (trans) vec = vec && ((uintptr_t)v % 16) == 0;
  const DensePlan pl = dense_plan(num_sms, trans, m, n, vec ? Wv : 1);
  if (pl.gx > 0x7fffffff || pl.nsplit > 65535) B2O_FAIL(B2O_EARG, "dense: matrix too large");
  if (pl.nsplit > 1 && (size_t)pl.nsplit * (size_t)out_len > part_elems) B2O_FAIL(B2O_ESTATE, "dense: workspace too small");
  DenseArgs a;
  a.M = M;
  a.v = v;
  a.res = res;
  a.m = m;
  a.n = n;
  a.lda = lda;
  a.chunk = pl.chunk;
  a.nsplit = pl.nsplit;
  a.alpha = alpha;
  a.beta = beta;
  a.part = part;
  a.tx_log2 = pl.tx_log2;
  const dim3 grid = pl.narrow ? dim3((unsigned)pl.gx) : dim3((unsigned)pl.gx, (unsigned)pl.nsplit);
  void (*kern)(const DenseArgs);
  if (pl.narrow)
    kern = !trans ? (vec ? dense_n_narrow_kernel<T, true> : dense_n_narrow_kernel<T, false>)
                  : (vec ? dense_t_narrow_kernel<T, true> : dense_t_narrow_kernel<T, false>);
  else
    kern = !trans ? (vec ? dense_n_kernel<T, true> : dense_n_kernel<T, false>)
                  : (vec ? dense_t_kernel<T, true> : dense_t_kernel<T, false>);
  B2O_LAUNCH(kern, grid, dim3(DN_THREADS), 0, stream, a);
  ++*launches;
  B2O_CUDA(cudaGetLastError());
  if (pl.nsplit > 1) {
    void (*fin)(const DenseArgs, int64_t) = dense_finish_kernel<T>;
    B2O_LAUNCH(fin, dim3((unsigned)ceil_div64(out_len, 256)), dim3(256), 0, stream, a, out_len);
    ++*launches;
    B2O_CUDA(cudaGetLastError());
  }
  return B2O_OK;
}
