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

// Software-pipelined variant of the row kernel -- THE DEFAULT (sparse_kernel = 0 or 3): the first four (idx, val) pairs of a
// lane's NEXT row are loaded into registers while the gathers of the CURRENT row are in flight, so a pass waits for one
// memory latency (x) instead of two in a row (idx -> x).  Costs registers (one resident CTA fewer per SM) and still wins on
// every pattern measured: 0.150 vs 0.164 ms (24 entries per row, half of them scattered), 0.058 vs 0.062 ms (5-point
// Laplacian), Float32 0.140 vs 0.159-0.170 and 0.046 vs 0.052 ms.  Same lane layout, same summation order as
// spmv_rows_kernel: the two produce identical bits.
constexpr int SP_PIPE_CTAS_PER_SM = 4;
template <typename T, int LL>
__global__ void __launch_bounds__(SP_THREADS, SP_PIPE_CTAS_PER_SM) spmv_rows_pipe_kernel(const __grid_constant__ SpmvArgs p) {
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
  // pipeline registers: offsets two rows ahead, first-trip entries one row ahead
  int64_t start = 0, end = 0, nstart = 0, nend = 0;
  int32_t ci[4] = {0, 0, 0, 0};
  T ca[4] = {(T)0, (T)0, (T)0, (T)0};
  {
    const int64_t r = r0 + group_in_warp;
    if (r < p.nrows) {
      start = __ldg(p.ptr + r);
      end = __ldg(p.ptr + r + 1);
    }
    if (r + step < p.nrows) {
      nstart = __ldg(p.ptr + r + step);
      nend = __ldg(p.ptr + r + step + 1);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t k = start + lane + u * L;
      if (k < end) {
        ci[u] = __ldg(idx + k);
        ca[u] = __ldg(val + k);
      }
    }
  }
  for (; r0 < p.nrows; r0 += step) {
    const int64_t r = r0 + group_in_warp;
    const bool valid = r < p.nrows;
    // offsets two rows ahead, entries of the next row
    int64_t n2start = 0, n2end = 0;
    if (r + 2 * step < p.nrows) {
      n2start = __ldg(p.ptr + r + 2 * step);
      n2end = __ldg(p.ptr + r + 2 * step + 1);
    }
    int32_t ni[4] = {0, 0, 0, 0};
    T na[4] = {(T)0, (T)0, (T)0, (T)0};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t k = nstart + lane + u * L;
      if (k < nend) {
        ni[u] = __ldg(idx + k);
        na[u] = __ldg(val + k);
      }
    }
    // current row: first trip from the pipeline registers
    T cx[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) cx[u] = (start + lane + u * L < end) ? __ldg(x + ci[u]) : (T)0;
    double s0 = fma((double)ca[0], (double)cx[0], 0.0);
    double s1 = fma((double)ca[1], (double)cx[1], 0.0);
    double s2 = fma((double)ca[2], (double)cx[2], 0.0);
    double s3 = fma((double)ca[3], (double)cx[3], 0.0);
    for (int64_t k = start + lane + 4 * (int64_t)L; k < end; k += 4 * (int64_t)L) {      // rows longer than 4 L entries
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
      const T x1 = p1 ? __ldg(x + i1) : (T)0;
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
    start = nstart;
    end = nend;
    nstart = n2start;
    nend = n2end;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ci[u] = ni[u];
      ca[u] = na[u];
    }
  }
}

// ================================================================== TMA-staged tile kernel
// The row kernel above is bound by its dependent load chain (ptr -> idx/val -> x: two DRAM latencies and one L2 latency per
// row batch; measured 47 % DRAM utilisation with the same duration for Float32 and Float64, about 2 us per warp pass).  The
// tile kernel takes the two streaming legs off that chain: the structure is cut ONCE, on the host at create time, into
// tiles of consecutive rows holding <= ST_C entries and <= ST_RT rows; a producer warp stages each tile's idx / val /
// row-offset slices global -> shared with 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx) into a ring, tiles
// ahead of their use.  The consumer warps then run the row kernel's inner step against SHARED memory: 2^LL lanes per row,
// four predicated (idx, val) pairs per lane from the staged slices, their four gathers from x in flight together, fixed-
// order butterfly, α / β, store.  Row batches are dealt round-robin to the warps across tiles and the warps only meet at the
// slot's "empty" barrier, so nobody waits inside a tile.  (A first version gathered tile-wide into shared products with a
// CTA barrier between gather and row sums: 0.28 ms against the row kernel's 0.19 ms -- the barrier exposes the slowest of
// 2048 gathers per tile; profiles/r1_sparse_tilekernel_v1.jsonl.)
// No atomics, fixed order -> bit-reproducible.  Bulk copies need 16-byte aligned sources: a tile's copy starts at its first
// entry rounded down to a multiple of 4 and is clamped to the last whole quad of the array (entries beyond it, at most 3, are
// read straight from global memory); row-offset slices start at an even row.  A row longer than ST_C is a "direct" tile:
// all consumer threads sum it straight from global memory (fixed-order block reduction).
constexpr int ST_C = 2048;                     // entries per tile (including <= 3 leading alignment entries)
constexpr int ST_RT = 1024;                    // rows per tile
constexpr int ST_NCONS = 384;                  // consumer threads (12 warps; 2 CTAs per SM)
constexpr int ST_CONS_WARPS = ST_NCONS / 32;
constexpr int ST_NTHREADS = ST_NCONS + 32;     // + producer warp
constexpr int ST_STAGES = 3;
constexpr int ST_CTAS_PER_SM = 2;
// one ring stage: idx[ST_C] int32 | val[ST_C] (8 bytes reserved per entry) | row offsets [ST_RT + 2] int64
constexpr size_t ST_IDX_OFF = 0;
constexpr size_t ST_VAL_OFF = ST_IDX_OFF + sizeof(int32_t) * ST_C;
constexpr size_t ST_PTR_OFF = ST_VAL_OFF + sizeof(double) * ST_C;
constexpr size_t ST_STAGE_BYTES = ST_PTR_OFF + sizeof(int64_t) * (ST_RT + 2);
constexpr size_t ST_BAR_OFF = ST_STAGE_BYTES * ST_STAGES;
constexpr size_t ST_SMEM_BYTES = ST_BAR_OFF + sizeof(uint64_t) * 2 * ST_STAGES;
static_assert(ST_STAGE_BYTES % 16 == 0 && ST_C % 4 == 0, "tile layout");

struct SpTile {      // 16 bytes; tiles[ntiles] is a sentinel {nnz, nrows, 0}
  int64_t e0;        // first entry staged (multiple of 4, <= ptr[r0]); direct tile: ptr[r0]
  int32_t r0;        // first row; the tile ends where the next one starts
  int32_t ne;        // entries staged (multiple of 4, e0 + ne >= ptr[r1]); -1: direct tile (one row longer than ST_C)
};

struct SpTileArgs {
  const SpTile *tiles;
  int64_t ntiles;
  const int64_t *ptr;   // [nrows + 1], 16-byte aligned, readable up to index nrows + 2
  const int32_t *idx;   // 16-byte aligned
  const void *val;      // 16-byte aligned
  const void *x;
  void *y;
  int64_t nrows;
  int64_t q_tail;       // nnz rounded down to a multiple of 4: no bulk copy reaches beyond it
  double alpha, beta;
};

#ifdef B2O_SIMT_EMU
inline void st_consumers_sync() { emu::named_barrier_sync(1, ST_NCONS); }
#else
__device__ __forceinline__ void st_consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(ST_NCONS) : "memory"); }
#endif

template <typename T, int LL>
__global__ void __launch_bounds__(ST_NTHREADS, ST_CTAS_PER_SM) spmv_tiles_kernel(const __grid_constant__ SpTileArgs p) {
  constexpr int L = 1 << LL;
  constexpr int GROUPS_PER_WARP = 32 >> LL;
#ifdef B2O_SIMT_EMU
  unsigned char *smem_raw = emu::dyn_smem();
#else
  extern __shared__ __align__(128) unsigned char smem_raw[];
#endif
  __shared__ double s_red[ST_CONS_WARPS];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + ST_BAR_OFF);
  uint64_t *empty = full + ST_STAGES;
  const int tid = threadIdx.x, lane32 = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < ST_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], ST_CONS_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // contiguous tile range of this CTA
  const int64_t tb = p.ntiles * (int64_t)blockIdx.x / (int64_t)gridDim.x;
  const int64_t te = p.ntiles * ((int64_t)blockIdx.x + 1) / (int64_t)gridDim.x;
  const T *__restrict__ val = (const T *)p.val;
  uint32_t slot = 0, par = 0;

  if (warp == ST_CONS_WARPS) {
    // ---------------------------------------------------------------- producer: one elected lane issues the bulk copies
    if (lane32 == 0 && tb < te) {
      SpTile cur = p.tiles[tb], nxt = p.tiles[tb + 1];           // descriptors are fetched two tiles ahead (sentinel at ntiles)
      for (int64_t t = tb; t < te; ++t) {
        const SpTile nn = p.tiles[t + 2 <= p.ntiles ? t + 2 : p.ntiles];
        if (cur.ne >= 0) {
          const int64_t r1 = nxt.r0;
          const int64_t rbase = (int64_t)cur.r0 & ~(int64_t)1;
          const uint32_t nptr = (uint32_t)(((r1 - rbase + 1) + 1) & ~(int64_t)1);      // offsets rbase .. r1, even count
          int64_t ncopy = cur.ne;
          if (cur.e0 + ncopy > p.q_tail) ncopy = p.q_tail > cur.e0 ? p.q_tail - cur.e0 : 0;
          unsigned char *stage = smem_raw + (size_t)slot * ST_STAGE_BYTES;
          mbar_wait(&empty[slot], par ^ 1u);
          mbar_expect_tx(&full[slot], (uint32_t)(ncopy * (sizeof(int32_t) + sizeof(T)) + nptr * sizeof(int64_t)));
          bulk_g2s(stage + ST_PTR_OFF, p.ptr + rbase, (uint32_t)(nptr * sizeof(int64_t)), &full[slot]);
          if (ncopy > 0) {
            bulk_g2s(stage + ST_IDX_OFF, p.idx + cur.e0, (uint32_t)(ncopy * sizeof(int32_t)), &full[slot]);
            bulk_g2s(stage + ST_VAL_OFF, val + cur.e0, (uint32_t)(ncopy * sizeof(T)), &full[slot]);
          }
          if (++slot == ST_STAGES) {
            slot = 0;
            par ^= 1u;
          }
        }
        cur = nxt;
        nxt = nn;
      }
    }
    __syncwarp();
    return;
  }

  // ------------------------------------------------------------------ consumers
  const T *__restrict__ x = (const T *)p.x;
  T *y = (T *)p.y;
  const int lane = tid & (L - 1);
  const int group_in_warp = lane32 >> LL;
  if (tb >= te) return;
  SpTile cur = p.tiles[tb], nxt = p.tiles[tb + 1];
  int batch_base = 0;
  for (int64_t t = tb; t < te; ++t) {
    const SpTile nn = p.tiles[t + 2 <= p.ntiles ? t + 2 : p.ntiles];
    const int64_t r0 = cur.r0, r1 = nxt.r0;
    if (cur.ne < 0) {
      // ---- direct tile: one long row summed from global memory by all consumers, fixed-order block reduction
      const int64_t start = __ldg(p.ptr + r0), end = __ldg(p.ptr + r0 + 1);
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      for (int64_t k = start + tid; k < end; k += 4 * ST_NCONS) {
        const int64_t k1 = k + ST_NCONS, k2 = k + 2 * ST_NCONS, k3 = k + 3 * ST_NCONS;
        const bool p1 = k1 < end, p2 = k2 < end, p3 = k3 < end;
        const int32_t i0 = __ldg(p.idx + k);
        const int32_t i1 = p1 ? __ldg(p.idx + k1) : 0;
        const int32_t i2 = p2 ? __ldg(p.idx + k2) : 0;
        const int32_t i3 = p3 ? __ldg(p.idx + k3) : 0;
        const T a0 = __ldg(val + k);
        const T a1 = p1 ? __ldg(val + k1) : (T)0;
        const T a2 = p2 ? __ldg(val + k2) : (T)0;
        const T a3 = p3 ? __ldg(val + k3) : (T)0;
        const T x0 = __ldg(x + i0);
        const T x1 = p1 ? __ldg(x + i1) : (T)0;
        const T x2 = p2 ? __ldg(x + i2) : (T)0;
        const T x3 = p3 ? __ldg(x + i3) : (T)0;
        s0 = fma((double)a0, (double)x0, s0);
        s1 = fma((double)a1, (double)x1, s1);
        s2 = fma((double)a2, (double)x2, s2);
        s3 = fma((double)a3, (double)x3, s3);
      }
      double s = (s0 + s1) + (s2 + s3);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane32 == 0) s_red[warp] = s;
      st_consumers_sync();
      if (tid == 0) {
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < ST_CONS_WARPS; ++w) tot += s_red[w];
        double r = p.alpha * tot;
        if (p.beta != 0.0) r += p.beta * (double)y[r0];
        y[r0] = (T)r;
      }
      st_consumers_sync();                                        // s_red is reused by the next direct tile
      cur = nxt;
      nxt = nn;
      continue;
    }

    unsigned char *stage = smem_raw + (size_t)slot * ST_STAGE_BYTES;
    const int32_t *sidx = reinterpret_cast<const int32_t *>(stage + ST_IDX_OFF);
    const T *sval = reinterpret_cast<const T *>(stage + ST_VAL_OFF);
    const int64_t *sptr = reinterpret_cast<const int64_t *>(stage + ST_PTR_OFF) + (r0 & 1);      // sptr[i] = ptr[r0 + i]
    mbar_wait(&full[slot], par);
    const int64_t e0 = cur.e0;
    int64_t avail = p.q_tail - e0;                                 // entries of this tile the bulk copies may cover
    avail = avail < 0 ? 0 : avail;
    const int staged = (int)(avail < (int64_t)cur.ne ? avail : (int64_t)cur.ne);
    const int nr = (int)(r1 - r0);
    const int nb = (nr + GROUPS_PER_WARP - 1) / GROUPS_PER_WARP;   // row batches of this tile (one batch = one warp pass)
    // batches are dealt round-robin to the consumer warps ACROSS tiles (batch_base counts them modulo the warp count), so
    // the warps stay balanced although a tile rarely holds a multiple of ST_CONS_WARPS batches; warps never wait for each
    // other inside a tile -- a fast warp moves on to the next staged tile
    int b = warp - batch_base;
    b += b < 0 ? ST_CONS_WARPS : 0;
    for (; b < nb; b += ST_CONS_WARPS) {
      const int rr = b * GROUPS_PER_WARP + group_in_warp;
      const bool valid = rr < nr;
      int start = 0, end = 0;
      if (valid) {
        start = (int)(sptr[rr] - e0);
        end = (int)(sptr[rr + 1] - e0);
      }
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      for (int k = start + lane; k < end; k += 4 * L) {
        int32_t ci[4];
        T ca[4], cx[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int jj = k + u * L;
          ci[u] = 0;
          ca[u] = (T)0;
          if (jj < end) {
            if (jj < staged) {
              ci[u] = sidx[jj];
              ca[u] = sval[jj];
            } else {                                               // the <= 3 entries beyond the last whole quad of the arrays
              ci[u] = __ldg(p.idx + e0 + jj);
              ca[u] = __ldg(val + e0 + jj);
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) cx[u] = (k + u * L < end) ? __ldg(x + ci[u]) : (T)0;
        s0 = fma((double)ca[0], (double)cx[0], s0);
        s1 = fma((double)ca[1], (double)cx[1], s1);
        s2 = fma((double)ca[2], (double)cx[2], s2);
        s3 = fma((double)ca[3], (double)cx[3], s3);
      }
      double s = (s0 + s1) + (s2 + s3);
#pragma unroll
      for (int o = L >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (valid && lane == 0) {
        const int64_t r = r0 + rr;
        double tv = p.alpha * s;
        if (p.beta != 0.0) tv += p.beta * (double)y[r];
        y[r] = (T)tv;
      }
    }
    batch_base = (batch_base + nb) % ST_CONS_WARPS;
    // hand the slot back: all shared-memory reads of this warp are done
    __syncwarp();
    if (lane32 == 0) mbar_arrive(&empty[slot]);
    if (++slot == ST_STAGES) {
      slot = 0;
      par ^= 1u;
    }
    cur = nxt;
    nxt = nn;
  }
}

// dst[k] = src[perm[k]]: refreshes the values of the transposed copy (structure transposition happens once, on the host)
template <typename T>
__global__ void __launch_bounds__(SP_THREADS) perm_gather_kernel(T *dst, const T *src, const int64_t *perm, int64_t nnz) {
  const int64_t stride = (int64_t)gridDim.x * SP_THREADS;
  for (int64_t k = (int64_t)blockIdx.x * SP_THREADS + threadIdx.x; k < nnz; k += stride) dst[k] = __ldg(src + __ldg(perm + k));
}

// ------------------------------------------------------------------ host side
// lanes per row: the smallest power of two >= mean row length / 5, at most a warp.  Measured (profiles/r1_sparse.jsonl, lane
// sweep): few lanes with several entries each beat one entry per lane -- 5 entries per row: 1 lane 0.058 ms, 2 lanes 0.078 ms,
// 4 lanes 0.135 ms; 24 entries per row: 4 or 8 lanes 0.150-0.157 ms, 16 lanes 0.231 ms, 32 lanes 0.407 ms.
static inline int spmv_lanes_log2(int64_t nrows, int64_t nnz) {
  const int64_t mean = nrows > 0 ? (nnz + nrows - 1) / nrows : 0;
  const int64_t want = (mean + 4) / 5;
  int l = 0;
  while (l < 5 && ((int64_t)1 << l) < want) ++l;
  return l;
}
static inline int64_t spmv_grid(int num_sms, int64_t nrows, int lanes_log2, int ctas_per_sm) {
  const int64_t threads = std::max<int64_t>(1, nrows) << lanes_log2;
  const int64_t want = (threads + SP_THREADS - 1) / SP_THREADS;
  return std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms * ctas_per_sm));   // one resident wave, rows grid-strided
}

// pipe: the software-pipelined variant; lanes_override >= 0 forces the lane-group width 2^lanes_override (tuning / tests)
template <typename T>
static int spmv_run_impl(int num_sms, B2O_STREAM_T stream, int64_t *launches, const int64_t *ptr, const int32_t *idx, const void *val,
                         int64_t nrows, int64_t nnz, void *y, const void *x, double alpha, double beta, bool pipe = false,
                         int lanes_override = -1) {
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
  a.lanes_log2 = (lanes_override >= 0 && lanes_override <= 5) ? lanes_override : spmv_lanes_log2(nrows, nnz);
  void (*kern)(const SpmvArgs) = nullptr;
  switch (a.lanes_log2) {
    case 0: kern = pipe ? spmv_rows_pipe_kernel<T, 0> : spmv_rows_kernel<T, 0>; break;
    case 1: kern = pipe ? spmv_rows_pipe_kernel<T, 1> : spmv_rows_kernel<T, 1>; break;
    case 2: kern = pipe ? spmv_rows_pipe_kernel<T, 2> : spmv_rows_kernel<T, 2>; break;
    case 3: kern = pipe ? spmv_rows_pipe_kernel<T, 3> : spmv_rows_kernel<T, 3>; break;
    case 4: kern = pipe ? spmv_rows_pipe_kernel<T, 4> : spmv_rows_kernel<T, 4>; break;
    default: kern = pipe ? spmv_rows_pipe_kernel<T, 5> : spmv_rows_kernel<T, 5>; break;
  }
  B2O_LAUNCH(kern, dim3((unsigned)spmv_grid(num_sms, nrows, a.lanes_log2, pipe ? SP_PIPE_CTAS_PER_SM : SP_CTAS_PER_SM)),
             dim3(SP_THREADS), 0, stream, a);
  ++*launches;
  B2O_CUDA(cudaGetLastError());
  return B2O_OK;
}

// ---- tile kernel, host side ----------------------------------------------------------------------------------------
#ifndef B2O_SIMT_EMU
#define B2O_FUNC_SMEM(kern, bytes) \
  B2O_CUDA(cudaFuncSetAttribute((const void *)(kern), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)))
#else
#define B2O_FUNC_SMEM(kern, bytes) \
  do {                             \
  } while (0)
#endif

// cut a compressed-row structure (0-based host offsets) into tiles; returns ntiles, `out` holds ntiles + 1 descriptors
template <typename Vec>
static inline int64_t spmv_build_tiles(const int64_t *ptr, int64_t nrows, int64_t nnz, Vec &out) {
  out.clear();
  int64_t r = 0;
  while (r < nrows) {
    SpTile t;
    const int64_t e0 = ptr[r] & ~(int64_t)3;
    if (ptr[r + 1] - e0 > ST_C) {                      // a row that does not fit a tile: summed straight from global memory
      t.e0 = ptr[r];
      t.r0 = (int32_t)r;
      t.ne = -1;
      out.push_back(t);
      ++r;
      continue;
    }
    int64_t r1 = r + 1;
    while (r1 < nrows && r1 - r < ST_RT && ptr[r1 + 1] - e0 <= ST_C) ++r1;
    t.e0 = e0;
    t.r0 = (int32_t)r;
    t.ne = (int32_t)(((ptr[r1] - e0) + 3) & ~(int64_t)3);
    out.push_back(t);
    r = r1;
  }
  const int64_t ntiles = (int64_t)out.size();
  SpTile sentinel;
  sentinel.e0 = nnz;
  sentinel.r0 = (int32_t)nrows;
  sentinel.ne = 0;
  out.push_back(sentinel);
  return ntiles;
}

template <typename T>
static int spmv_tiles_run_impl(int num_sms, B2O_STREAM_T stream, int64_t *launches, const SpTile *tiles, int64_t ntiles,
                               const int64_t *ptr, const int32_t *idx, const void *val, int64_t nrows, int64_t nnz, void *y,
                               const void *x, double alpha, double beta, int lanes_override = -1) {
  if (nrows == 0 || ntiles == 0) return B2O_OK;
  SpTileArgs a;
  a.tiles = tiles;
  a.ntiles = ntiles;
  a.ptr = ptr;
  a.idx = idx;
  a.val = val;
  a.x = x;
  a.y = y;
  a.nrows = nrows;
  a.q_tail = nnz & ~(int64_t)3;
  a.alpha = alpha;
  a.beta = beta;
  void (*kern)(const SpTileArgs) = nullptr;
  switch ((lanes_override >= 0 && lanes_override <= 5) ? lanes_override : spmv_lanes_log2(nrows, nnz)) {
    case 0: kern = spmv_tiles_kernel<T, 0>; break;
    case 1: kern = spmv_tiles_kernel<T, 1>; break;
    case 2: kern = spmv_tiles_kernel<T, 2>; break;
    case 3: kern = spmv_tiles_kernel<T, 3>; break;
    case 4: kern = spmv_tiles_kernel<T, 4>; break;
    default: kern = spmv_tiles_kernel<T, 5>; break;
  }
  B2O_FUNC_SMEM(kern, ST_SMEM_BYTES);
  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)num_sms * ST_CTAS_PER_SM));
  B2O_LAUNCH(kern, dim3((unsigned)grid), dim3(ST_NTHREADS), ST_SMEM_BYTES, stream, a);
  ++*launches;
  B2O_CUDA(cudaGetLastError());
  return B2O_OK;
}
