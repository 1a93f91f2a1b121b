// b2o_stream.cuh -- the streaming substrate shared by the persistent quasi-Newton kernels.
//
// One CTA per SM, 8 consumer warps + 1 producer warp.  Library-owned column data (16-byte aligned,
// pitch padded with zeros to a whole number of tiles) is staged global -> shared by 1-D TMA bulk
// copies (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) into a ring of `stages` chunks of R
// rows; consumers wait on the chunk's "full" mbarrier, read it with conflict-free LDS.128 and hand
// the slot back through its "empty" mbarrier.  Caller-owned vectors (x, res: any 8-byte alignment,
// any length) are read/written straight from registers with a one-tile-ahead prefetch.
//
// Row <-> thread mapping inside a tile: pair e = j*256 + tid (j < EPT/2) holds rows 2e, 2e+1.
#pragma once
#include "b2o_internal.cuh"

constexpr int B2O_NCONS = 256;                     // consumer threads
constexpr int B2O_NTHREADS = B2O_NCONS + 32;       // + producer warp
constexpr int B2O_CONS_WARPS = B2O_NCONS / 32;

struct Ring {
  double *buf;       // [stages][R]
  uint64_t *full;    // [stages]
  uint64_t *empty;   // [stages]
  int stages;
};

// item counter -> (slot, parity)
struct RingPos {
  uint32_t slot = 0, par = 0;
  int stages;
  __device__ __forceinline__ explicit RingPos(int s) : stages(s) {}
  __device__ __forceinline__ void advance() {
    if (++slot == (uint32_t)stages) {
      slot = 0;
      par ^= 1u;
    }
  }
};

__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(B2O_NCONS) : "memory"); }

template <int R>
__device__ __forceinline__ void producer_push(const Ring &rg, RingPos &pos, const double *src) {
  mbar_wait(&rg.empty[pos.slot], pos.par ^ 1u);
  mbar_expect_tx(&rg.full[pos.slot], (uint32_t)(R * sizeof(double)));
  bulk_g2s(rg.buf + (size_t)pos.slot * R, src, (uint32_t)(R * sizeof(double)), &rg.full[pos.slot]);
  pos.advance();
}

__device__ __forceinline__ void consumer_release(const Ring &rg, uint32_t slot) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&rg.empty[slot]);
}

// caller-owned vector tile -> registers (zero beyond n)
template <int R>
__device__ __forceinline__ void load_user_tile(const double *__restrict__ p, int64_t row0, int64_t n, bool al16,
                                               double (&out)[R / B2O_NCONS]) {
  constexpr int EPT = R / B2O_NCONS;
#pragma unroll
  for (int j = 0; j < EPT / 2; ++j) {
    int64_t r = row0 + 2 * ((int64_t)j * B2O_NCONS + threadIdx.x);
    if (al16 && r + 1 < n) {
      double2 v = ldg_stream2(p + r);
      out[2 * j] = v.x;
      out[2 * j + 1] = v.y;
    } else {
      out[2 * j] = (r < n) ? p[r] : 0.0;
      out[2 * j + 1] = (r + 1 < n) ? p[r + 1] : 0.0;
    }
  }
}
template <int R>
__device__ __forceinline__ void store_user_tile(double *__restrict__ p, int64_t row0, int64_t n, bool al16,
                                                const double (&v)[R / B2O_NCONS]) {
  constexpr int EPT = R / B2O_NCONS;
#pragma unroll
  for (int j = 0; j < EPT / 2; ++j) {
    int64_t r = row0 + 2 * ((int64_t)j * B2O_NCONS + threadIdx.x);
    if (al16 && r + 1 < n) {
      stg_stream2(p + r, make_double2(v[2 * j], v[2 * j + 1]));
    } else {
      if (r < n) p[r] = v[2 * j];
      if (r + 1 < n) p[r + 1] = v[2 * j + 1];
    }
  }
}

// carve dynamic shared memory: ring | accs | coef | barriers
struct SmemLayout {
  size_t ring_off, accs_off, coef_off, bar_off, total;
};
static inline SmemLayout smem_layout(int R, int stages, int acc_cols) {
  SmemLayout L;
  L.ring_off = 0;
  L.accs_off = L.ring_off + (size_t)stages * R * sizeof(double);
  L.coef_off = L.accs_off + (size_t)acc_cols * B2O_NCONS * sizeof(double);
  L.bar_off = L.coef_off + (size_t)(B2O_MAX_COLS + 8) * sizeof(double);
  L.total = L.bar_off + (size_t)2 * stages * sizeof(uint64_t);
  return L;
}
