// b2o_stream.cuh -- the streaming substrate shared by the persistent quasi-Newton kernels.
//
// One CTA per SM, 8 consumer warps + 1 producer warp.  Library-owned column data (16-byte aligned,
// pitch padded with zeros to a whole number of tiles) is staged global -> shared by 1-D TMA bulk
// copies (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) into a ring of `stages` chunks of R
// rows; consumers wait on the chunk's "full" mbarrier, read it with conflict-free LDS.128 and hand
// the slot back through its "empty" mbarrier.  Caller-owned vectors (x, res: any 8-byte alignment,
// any length) are read/written straight from registers with a one-tile-ahead prefetch.
//
// Row <-> thread mapping inside a tile: 16-byte vector e = j*256 + tid (j < EPT/VN) holds rows VN*e .. VN*e + VN-1
// (VN = 2 for Float64, 4 for Float32).  The kernels are templates on the element type T; inner products are always
// accumulated in double, elementwise statements run in T with the reference's rounding (no contraction).
#pragma once
#ifdef B2O_SIMT_EMU
#include "b2o_shared_defs.h"   // host SIMT emulator (tests/emu): the emulator header supplies mbarrier / bulk-copy / barrier stand-ins
#else
#include "b2o_internal.cuh"
#endif

template <typename T>
struct Vec16;
template <>
struct Vec16<double> {
  using type = double2;
  static constexpr int N = 2;
};
template <>
struct Vec16<float> {
  using type = float4;
  static constexpr int N = 4;
};
__device__ __forceinline__ void vec_unpack(const double2 &v, double *o) {
  o[0] = v.x;
  o[1] = v.y;
}
__device__ __forceinline__ void vec_unpack(const float4 &v, float *o) {
  o[0] = v.x;
  o[1] = v.y;
  o[2] = v.z;
  o[3] = v.w;
}
__device__ __forceinline__ double2 vec_pack(const double *o) { return make_double2(o[0], o[1]); }
__device__ __forceinline__ float4 vec_pack(const float *o) { return make_float4(o[0], o[1], o[2], o[3]); }
#ifdef B2O_SIMT_EMU
inline float4 ldg_stream16(const float *p) { return make_float4(p[0], p[1], p[2], p[3]); }
inline double2 ldg_stream16(const double *p) { return make_double2(p[0], p[1]); }
inline void stg_stream16(float *p, float4 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w; }
inline void stg_stream16(double *p, double2 v) { p[0] = v.x; p[1] = v.y; }
#else
__device__ __forceinline__ float4 ldg_stream16(const float *p) {
  float4 r;
  asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ldg_stream16(const double *p) { return ldg_stream2(p); }
__device__ __forceinline__ void stg_stream16(float *p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream16(double *p, double2 v) { stg_stream2(p, v); }
#endif

constexpr int B2O_NCONS = 256;                     // consumer threads
constexpr int B2O_NTHREADS = B2O_NCONS + 32;       // + producer warp
constexpr int B2O_CONS_WARPS = B2O_NCONS / 32;

struct Ring {
  unsigned char *buf;  // [stages][R * sizeof(T)]
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

#ifdef B2O_SIMT_EMU
inline void consumers_sync() { emu::named_barrier_sync(1, B2O_NCONS); }
#else
__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(B2O_NCONS) : "memory"); }
#endif

template <int R, typename T>
__device__ __forceinline__ void producer_push(const Ring &rg, RingPos &pos, const T *src) {
  mbar_wait(&rg.empty[pos.slot], pos.par ^ 1u);
  mbar_expect_tx(&rg.full[pos.slot], (uint32_t)(R * sizeof(T)));
  bulk_g2s(rg.buf + (size_t)pos.slot * R * sizeof(T), src, (uint32_t)(R * sizeof(T)), &rg.full[pos.slot]);
  pos.advance();
}
// the thread's EPT elements of the staged tile in ring slot `slot` (conflict-free LDS.128)
template <int R, typename T>
__device__ __forceinline__ void tile_from_ring(const Ring &rg, uint32_t slot, T (&out)[R / B2O_NCONS]) {
  using V = typename Vec16<T>::type;
  constexpr int VN = Vec16<T>::N, EPT = R / B2O_NCONS;
  const V *b = reinterpret_cast<const V *>(rg.buf + (size_t)slot * R * sizeof(T));
#pragma unroll
  for (int j = 0; j < EPT / VN; ++j) vec_unpack(b[j * B2O_NCONS + threadIdx.x], &out[VN * j]);
}
// library-owned, padded column: the thread's EPT elements of tile row0 .. row0 + R go back with 16-byte stores
template <int R, typename T>
__device__ __forceinline__ void tile_to_owned(T *col_tile, const T (&v)[R / B2O_NCONS]) {
  using V = typename Vec16<T>::type;
  constexpr int VN = Vec16<T>::N, EPT = R / B2O_NCONS;
  V *g = reinterpret_cast<V *>(col_tile);
#pragma unroll
  for (int j = 0; j < EPT / VN; ++j) g[j * B2O_NCONS + threadIdx.x] = vec_pack(&v[VN * j]);
}
// thread-local inner product of two register tiles, accumulated in double: VN independent chains (one per vector lane,
// the Float64 kernel's s0 / s1), folded pairwise
template <int EPT, typename T>
__device__ __forceinline__ double tile_dot(const T (&a)[EPT], const T (&b)[EPT]) {
  constexpr int VN = Vec16<T>::N;
  double s[VN];
#pragma unroll
  for (int e = 0; e < VN; ++e) s[e] = 0.0;
#pragma unroll
  for (int j = 0; j < EPT / VN; ++j)
#pragma unroll
    for (int e = 0; e < VN; ++e) s[e] = fma((double)a[VN * j + e], (double)b[VN * j + e], s[e]);
  if (VN == 2) return s[0] + s[1];
  return (s[0] + s[1]) + (s[VN - 2] + s[VN - 1]);
}

__device__ __forceinline__ void consumer_release(const Ring &rg, uint32_t slot) {
  __syncwarp();
  if ((threadIdx.x & 31) == 0) mbar_arrive(&rg.empty[slot]);
}

// caller-owned vector tile -> registers (zero beyond n)
template <int R, typename T>
__device__ __forceinline__ void load_user_tile(const T *__restrict__ p, int64_t row0, int64_t n, bool al16, T (&out)[R / B2O_NCONS]) {
  constexpr int VN = Vec16<T>::N, EPT = R / B2O_NCONS;
#pragma unroll
  for (int j = 0; j < EPT / VN; ++j) {
    int64_t r = row0 + VN * ((int64_t)j * B2O_NCONS + threadIdx.x);
    if (al16 && r + VN - 1 < n) {
      vec_unpack(ldg_stream16(p + r), &out[VN * j]);
    } else {
#pragma unroll
      for (int e = 0; e < VN; ++e) out[VN * j + e] = (r + e < n) ? p[r + e] : (T)0;
    }
  }
}
template <int R, typename T>
__device__ __forceinline__ void store_user_tile(T *__restrict__ p, int64_t row0, int64_t n, bool al16, const T (&v)[R / B2O_NCONS]) {
  constexpr int VN = Vec16<T>::N, EPT = R / B2O_NCONS;
#pragma unroll
  for (int j = 0; j < EPT / VN; ++j) {
    int64_t r = row0 + VN * ((int64_t)j * B2O_NCONS + threadIdx.x);
    if (al16 && r + VN - 1 < n) {
      stg_stream16(p + r, vec_pack(&v[VN * j]));
    } else {
#pragma unroll
      for (int e = 0; e < VN; ++e)
        if (r + e < n) p[r + e] = v[VN * j + e];
    }
  }
}

// carve dynamic shared memory: ring | accs | coef | barriers
struct SmemLayout {
  size_t ring_off, accs_off, coef_off, bar_off, total;
};
static inline SmemLayout smem_layout(int R, int stages, int acc_cols, size_t esize = sizeof(double)) {
  SmemLayout L;
  L.ring_off = 0;
  L.accs_off = L.ring_off + (size_t)stages * R * esize;
  L.coef_off = L.accs_off + (size_t)acc_cols * B2O_NCONS * sizeof(double);
  L.bar_off = L.coef_off + (size_t)(B2O_MAX_COLS + 8) * sizeof(double);
  L.total = L.bar_off + (size_t)2 * stages * sizeof(uint64_t);
  return L;
}
