// b2o_sparse_kernels.cuh -- kernels + launch logic of the sparse-matrix leaf (see b2o_sparse.cu for the description).
// Free of CUDA runtime calls so that tests/emu/ compiles the SAME code for the host under the SIMT emulator.
#pragma once
#include <algorithm>

constexpr int SP_THREADS = 256;
constexpr int SP_CTAS_PER_SM = 5;   // 48 registers per thread (Float64 instantiations) -> 5 resident CTAs of 256 threads

struct SpmvArgs {
  const int64_t *ptr;   // [nrows + 1] offsets into idx / val (0-based)
  const int32_t *idx;   // [nnz] 0-based index into x
  const void *val;      // [nnz]
  const void *x;
  void *y;
  int64_t nrows;
  double alpha, beta;
  int lanes_log2;       // 2^lanes_log2 consecutive lanes share one row (selects the instantiation)
};

// y[r] = α Σ_{k in row r} val[k] x[idx[k]] (+ β y[r]) for a compressed-row structure (CSR of the matrix, or the CSC arrays
// read as the CSR of its transpose).  L = 2^LL consecutive lanes own a row: lane l takes the entries start+l, start+l+L, ...
// (coalesced idx / val reads).  Every trip of the entry loop issues FOUR predicated (idx, val) pairs and their four gathers
// from x before the first multiply -- the row's tail goes through the same 4-wide body, so short rows (the common case) keep
// four independent load chains per lane in flight instead of one.  Four partial sums per lane, then a shuffle butterfly over
// the L lanes in a fixed order -> bit-reproducible.  Rows are dealt to the lane groups warp by warp with a warp-uniform
// trip count, so every lane of a warp reaches every shuffle (rows past the end contribute nothing); the offsets of the
// NEXT row are fetched before the current row is summed (the ptr -> idx -> x chain is three dependent latencies).
template <typename T, int LL>
__global__ void __launch_bounds__(SP_THREADS, SP_CTAS_PER_SM) spmv_rows_kernel(const __grid_constant__ SpmvArgs p) {
  constexpr int L = 1 << LL;
  constexpr int GROUPS_PER_WARP = 32 >> LL;
  const T *__restrict__ val = (const T *)p.val;
  const T *__restrict__ x = (const T *)p.x;
  const int32_t *__restrict__ idx = p.idx;
  T *y = (T *)p.y;
  const int lane = threadIdx.x & (L - 1);
  const int64_t warp_id = ((int64_t)blockIdx.x * SP_THREADS + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * SP_THREADS) >> 5;
  const int group_in_warp = (threadIdx.x & 31) >> LL;
  const int64_t step = nwarps * GROUPS_PER_WARP;
  int64_t r0 = warp_id * GROUPS_PER_WARP;
  int64_t nstart = 0, nend = 0;
  if (r0 + group_in_warp < p.nrows) {
    nstart = __ldg(p.ptr + r0 + group_in_warp);
    nend = __ldg(p.ptr + r0 + group_in_warp + 1);
  }
  for (; r0 < p.nrows; r0 += step) {
    const int64_t r = r0 + group_in_warp;
    const bool valid = r < p.nrows;
    const int64_t start = nstart, end = nend;
    nstart = nend = 0;
    if (r + step < p.nrows) {                 // offsets of this group's next row
      nstart = __ldg(p.ptr + r + step);
      nend = __ldg(p.ptr + r + step + 1);
    }
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int64_t k = start + lane; k < end; k += 4 * (int64_t)L) {
      const int64_t k1 = k + L, k2 = k + 2 * L, k3 = k + 3 * L;
      const bool p1 = k1 < end, p2 = k2 < end, p3 = k3 < end;
      const int32_t i0 = __ldg(idx + k);
      const int32_t i1 = p1 ? __ldg(idx + k1) : 0;
      const int32_t i2 = p2 ? __ldg(idx + k2) : 0;
      const int32_t i3 = p3 ? __ldg(idx + k3) : 0;
      const T a0 = __ldg(val + k);
      const T a1 = p1 ? __ldg(val + k1) : (T)0;
      const T a2 = p2 ? __ldg(val + k2) : (T)0;
      const T a3 = p3 ? __ldg(val + k3) : (T)0;
      const T x0 = __ldg(x + i0);
      const T x1 = p1 ? __ldg(x + i1) : (T)0;      // predicated too: 0 * x[0] would let an Inf / NaN of x[0] leak in
      const T x2 = p2 ? __ldg(x + i2) : (T)0;
      const T x3 = p3 ? __ldg(x + i3) : (T)0;
      s0 = fma((double)a0, (double)x0, s0);
      s1 = fma((double)a1, (double)x1, s1);
      s2 = fma((double)a2, (double)x2, s2);
      s3 = fma((double)a3, (double)x3, s3);
    }
    double s = (s0 + s1) + (s2 + s3);
#pragma unroll
    for (int o = L >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (valid && lane == 0) {
      double t = p.alpha * s;
      if (p.beta != 0.0) t += p.beta * (double)y[r];
      y[r] = (T)t;
    }
  }
}

// dst[k] = src[perm[k]]: refreshes the values of the transposed copy (structure transposition happens once, on the host)
template <typename T>
__global__ void __launch_bounds__(SP_THREADS) perm_gather_kernel(T *dst, const T *src, const int64_t *perm, int64_t nnz) {
  const int64_t stride = (int64_t)gridDim.x * SP_THREADS;
  for (int64_t k = (int64_t)blockIdx.x * SP_THREADS + threadIdx.x; k < nnz; k += stride) dst[k] = __ldg(src + __ldg(perm + k));
}

// ------------------------------------------------------------------ host side
// lanes per row: every lane takes four entries per trip, so the smallest power of two >= mean row length / 4, at most a warp
static inline int spmv_lanes_log2(int64_t nrows, int64_t nnz) {
  const int64_t mean = nrows > 0 ? (nnz + nrows - 1) / nrows : 0;
  const int64_t want = (mean + 3) / 4;
  int l = 0;
  while (l < 5 && ((int64_t)1 << l) < want) ++l;
  return l;
}
static inline int64_t spmv_grid(int num_sms, int64_t nrows, int lanes_log2) {
  const int64_t threads = std::max<int64_t>(1, nrows) << lanes_log2;
  const int64_t want = (threads + SP_THREADS - 1) / SP_THREADS;
  return std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms * SP_CTAS_PER_SM));   // one resident wave, rows grid-strided
}

template <typename T>
static int spmv_run_impl(int num_sms, B2O_STREAM_T stream, int64_t *launches, const int64_t *ptr, const int32_t *idx, const void *val,
                         int64_t nrows, int64_t nnz, void *y, const void *x, double alpha, double beta) {
  if (nrows == 0) return B2O_OK;
  SpmvArgs a;
  a.ptr = ptr;
  a.idx = idx;
  a.val = val;
  a.x = x;
  a.y = y;
  a.nrows = nrows;
  a.alpha = alpha;
  a.beta = beta;
  a.lanes_log2 = spmv_lanes_log2(nrows, nnz);
  void (*kern)(const SpmvArgs) = nullptr;
  switch (a.lanes_log2) {
    case 0: kern = spmv_rows_kernel<T, 0>; break;
    case 1: kern = spmv_rows_kernel<T, 1>; break;
    case 2: kern = spmv_rows_kernel<T, 2>; break;
    case 3: kern = spmv_rows_kernel<T, 3>; break;
    case 4: kern = spmv_rows_kernel<T, 4>; break;
    default: kern = spmv_rows_kernel<T, 5>; break;
  }
  B2O_LAUNCH(kern, dim3((unsigned)spmv_grid(num_sms, nrows, a.lanes_log2)), dim3(SP_THREADS), 0, stream, a);
  ++*launches;
  B2O_CUDA(cudaGetLastError());
  return B2O_OK;
}
