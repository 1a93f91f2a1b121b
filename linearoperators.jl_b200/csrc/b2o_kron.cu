// b2o_kron.cu -- kron(A,B)*vec as a TMA-fed tcgen05 GEMM pair (the only tensor-core path).
//
// Reference: src/kron.jl:14-40.  prod!:  X = reshape(x, q, n);  res = α·vec(B·X·Aᵀ) + β·res   (A m×n, B p×q, column-major)
//            tprod!/ctprod!:  X = reshape(x, p, m);  res = α·vec(Bᵀ·X·A) + β·res
// The reference materialises Matrix(B*X*transpose(A)) through n operator applies (3 GEMVs each).  Here it is exactly two
// GEMMs on the 5th-generation tensor cores, in ONE cooperative launch:
//   phase 0   Y[M×N1]  = A1[M×K1] · X'[N1×K1]ᵀ      (fp32 accumulate in TMEM; stored as a bf16 hi/lo pair so that the
//                                                   intermediate costs no accuracy; 2*M*N1*2 bytes, stays in L2)
//   phase 1   Z[M×N2]  = [Yhi|Ylo][M×2N1] · [B2|B2]ᵀ    epilogue: res[j*M+i] = α·Z[i,j] (+ β·res), bf16, column-major
// with every operand K-major (K contiguous) so one code path serves both directions:
//   prod :  M=p K1=q N1=n N2=m   A1 = B row-major (transposed copy made at create), X' = reshape(x,q,n)ᵀ = x as stored,
//           B2 = A row-major (transposed copy made at create)
//   tprod:  M=q K1=p N1=m N2=n   A1 = Bᵀ row-major = B as stored (column-major), X' = x as stored, B2 = Aᵀ row-major = A as stored
// Per CTA: warp 0 = TMA producer (cp.async.bulk.tensor.2d, 128-byte swizzle, 4-stage mbarrier ring), warp 1 = tcgen05.mma
// issuer (one elected thread; UMMA 128×BN×16, accumulator in TMEM) + TMEM alloc/dealloc, warps 2-5 = epilogue
// (tcgen05.ld 32x32b → registers → global).  `nb` right-hand sides are batched by stacking them along N (phase 0) / M (phase 1).
#include "b2o_internal.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <dlfcn.h>
#include <algorithm>

constexpr int KR_BM = 128, KR_BK = 64, KR_MAX_STAGES = 12, KR_THREADS = 192;
// ring depth: the per-SM stream is latency-bound (Little: bytes in flight / L2 latency), so use what shared memory allows
__host__ __device__ constexpr int kr_stages(int BN) { return (200 * 1024) / (KR_BM * KR_BK * 2 + BN * KR_BK * 2) > KR_MAX_STAGES ? KR_MAX_STAGES : (200 * 1024) / (KR_BM * KR_BK * 2 + BN * KR_BK * 2); }

struct KronArgs {
  int M, K1, N1, N2, nb;        // see header comment; phase-1 K = N1
  int ldy;                      // N1 rounded up to 64; a Y row holds [hi: ldy | lo: ldy] elements
  __nv_bfloat16 *Y;             // [(nb*M) × 2*ldy], padding columns stay zero
  void *res;                    // nb × (M*N2), each column-major M×N2; bf16 or (out_f32) fp32
  int out_f32;
  float alpha, beta;
  unsigned long long *bar;
  unsigned long long bar_target;
  unsigned long long *dbg;      // optional timeline of CTA 0 (%globaltimer, ns): [0] start [1] setup done [2+4*ph] first stage landed
                                // [3+4*ph] accumulator complete [4+4*ph] epilogue done [5+4*ph] phase end/barrier passed [10] exit
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define KR_STAMP(idx)                                                     \
  do {                                                                   \
    if (p.dbg && blockIdx.x == 0) p.dbg[idx] = gtimer();                  \
  } while (0)

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 B apart (SBO), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_sw128(const void *smem) {
  uint64_t d = (uint64_t)((smem_u32(smem) >> 4) & 0x3FFFu);
  d |= (uint64_t)(1024u >> 4) << 32;   // stride byte offset
  d |= (uint64_t)1 << 46;              // descriptor version
  d |= (uint64_t)2 << 61;              // SWIZZLE_128B
  return d;
}

// the kron launch is latency-bound (tens of CTAs): poll the barrier without sleeping
__device__ __forceinline__ void grid_barrier_tight(unsigned long long *ctr, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1ULL);
    while (ld_acquire_u64(ctr) < target) {}
    __threadfence();
  }
  __syncthreads();
}

// epilogue helpers: the per-element branches (result type, beta, bounds) are hoisted out of the 32-column loops
template <bool F32, bool BETA>
__device__ __forceinline__ void kron_store_cols(void *res, size_t off, int M, float alpha, float beta, const uint32_t (&v)[32], int nvalid) {
  if (nvalid >= 32) {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      float z = alpha * __uint_as_float(v[e]);
      if (F32) {
        float *dst = reinterpret_cast<float *>(res) + off + (size_t)e * M;
        if (BETA) z += beta * *dst;
        *dst = z;
      } else {
        __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(res) + off + (size_t)e * M;
        if (BETA) z += beta * __bfloat162float(*dst);
        *dst = __float2bfloat16_rn(z);
      }
    }
  } else {
#pragma unroll
    for (int e = 0; e < 32; ++e) {
      if (e >= nvalid) continue;   // static register indices: v[] must stay in registers
      float z = alpha * __uint_as_float(v[e]);
      if (F32) {
        float *dst = reinterpret_cast<float *>(res) + off + (size_t)e * M;
        if (BETA) z += beta * *dst;
        *dst = z;
      } else {
        __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(res) + off + (size_t)e * M;
        if (BETA) z += beta * __bfloat162float(*dst);
        *dst = __float2bfloat16_rn(z);
      }
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(KR_THREADS, 1)
kron_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmX,
                      const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmB2,
                      const __grid_constant__ KronArgs p) {
  constexpr uint32_t A_BYTES = KR_BM * KR_BK * 2, B_BYTES = BN * KR_BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int KR_STAGES = kr_stages(BN);
  // instruction descriptor: D=F32, A=B=BF16, both K-major, N=BN, M=128
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(KR_BM >> 4) << 24);
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SW128 needs 1024 B
  __shared__ __align__(8) uint64_t full[KR_STAGES], empty[KR_STAGES], tmem_full, tmem_empty;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) KR_STAMP(0);
  if (threadIdx.x == 32) {   // hide the descriptor fetches behind the setup
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmY) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < KR_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&tmem_full, 1);
    mbar_init(&tmem_empty, 4);
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"((uint32_t)BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  if (threadIdx.x == 0) KR_STAMP(1);

  uint32_t stage = 0, sphase = 0;   // smem ring position (producer and MMA warp each keep their own copy)
  uint32_t tphase = 0;              // accumulator hand-over parity (MMA warp and epilogue warps)
  unsigned long long bar_target = p.bar_target;

  for (int ph = 0; ph < 2; ++ph) {
    const CUtensorMap *tmA = ph == 0 ? &tmA1 : &tmY;
    const CUtensorMap *tmB = ph == 0 ? &tmX : &tmB2;
    const int Mrows = ph == 0 ? p.M : p.nb * p.M;
    const int Ncols = ph == 0 ? p.nb * p.N1 : p.N2;
    const int Kdim = ph == 0 ? p.K1 : p.N1;
    const int mt = (Mrows + KR_BM - 1) / KR_BM, nt = (Ncols + BN - 1) / BN;
    const int khalf = (Kdim + KR_BK - 1) / KR_BK;
    const int ntiles = mt * nt, kblocks = ph == 0 ? khalf : 2 * khalf;   // phase 1 runs over the hi and the lo half of Y

    if (warp == 0) {
      // ===== TMA producer
      if (lane == 0) {
        if (ph == 1) asm volatile("fence.proxy.async;" ::: "memory");   // Y was written with generic-proxy stores
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
          const int m0 = (t / nt) * KR_BM, n0 = (t % nt) * BN;
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(&empty[stage], sphase ^ 1u);
            mbar_expect_tx(&full[stage], STAGE_BYTES);
            unsigned char *sa = smem + (size_t)stage * STAGE_BYTES;
            const int kk = kb < khalf ? kb : kb - khalf;
            const int ka = (ph == 1 && kb >= khalf) ? p.ldy + kk * KR_BK : kk * KR_BK;
            tma_load_2d(sa, tmA, ka, m0, &full[stage]);
            tma_load_2d(sa + A_BYTES, tmB, kk * KR_BK, n0, &full[stage]);
            if (++stage == KR_STAGES) { stage = 0; sphase ^= 1u; }
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ===== MMA issuer
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(&tmem_empty, tphase ^ 1u);      // epilogue has drained the accumulator of the previous tile
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], sphase);
          tc_fence_after();
          if (lane == 0 && kb == 0 && t == (int)blockIdx.x) KR_STAMP(2 + 4 * ph);
          if (lane == 0) {
            unsigned char *sa = smem + (size_t)stage * STAGE_BYTES;
            const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + A_BYTES);
#pragma unroll
            for (int k = 0; k < KR_BK / 16; ++k)   // UMMA_K = 16 bf16 = 32 B: advance the start address inside the swizzle atom
              tc_mma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (kb | k) ? 1u : 0u);
            tc_commit(&empty[stage]);              // frees the smem stage when these MMAs have read it
            if (kb == kblocks - 1) tc_commit(&tmem_full);
          }
          __syncwarp();
          if (++stage == KR_STAGES) { stage = 0; sphase ^= 1u; }
        }
        tphase ^= 1u;
      }
    } else {
      // ===== epilogue: warp w owns TMEM lanes [32*(w%4), +32) = rows of the tile
      const int quarter = warp & 3;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int m0 = (t / nt) * KR_BM, n0 = (t % nt) * BN;
        mbar_wait(&tmem_full, tphase);
        tc_fence_after();
        if (warp == 2 && lane == 0 && t == (int)blockIdx.x) KR_STAMP(3 + 4 * ph);
        const int r = m0 + quarter * 32 + lane;    // global row of this thread
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32];
          tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
          if (ph == 1 && warp == 2 && lane == 0 && t == (int)blockIdx.x) KR_STAMP(11 + 2 * (c0 / 32));
          if (r < Mrows) {
            if (ph == 0) {
              // Y[(b*M + r) * 2*ldy + j] (hi) / + ldy (lo), global column jj = b*N1 + j
              const int jj0 = n0 + c0;
              const int b0 = jj0 / p.N1, j0 = jj0 - b0 * p.N1;
              if (jj0 + 32 <= Ncols && j0 + 32 <= p.N1) {
                // fast path: the 32 columns belong to one right-hand side; 16-byte stores (j0 is a multiple of 8, ldy of 64)
                __nv_bfloat16 *dst = p.Y + ((size_t)b0 * p.M + r) * (2 * (size_t)p.ldy) + j0;
#pragma unroll
                for (int g = 0; g < 32; g += 8) {
                  uint32_t hi[4], lo[4];
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float y0 = __uint_as_float(v[g + 2 * e]), y1 = __uint_as_float(v[g + 2 * e + 1]);
                    const __nv_bfloat162 h = __floats2bfloat162_rn(y0, y1);
                    const __nv_bfloat162 l = __floats2bfloat162_rn(y0 - __low2float(h), y1 - __high2float(h));
                    hi[e] = *reinterpret_cast<const uint32_t *>(&h);
                    lo[e] = *reinterpret_cast<const uint32_t *>(&l);
                  }
                  *reinterpret_cast<uint4 *>(dst + g) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                  *reinterpret_cast<uint4 *>(dst + g + p.ldy) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
              } else {
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                  const int jje = jj0 + e;
                  if (jje < Ncols) {
                    const int be = jje / p.N1, je = jje - be * p.N1;
                    const float y = __uint_as_float(v[e]);
                    const __nv_bfloat16 h = __float2bfloat16_rn(y);
                    __nv_bfloat16 *d1 = p.Y + ((size_t)be * p.M + r) * (2 * (size_t)p.ldy) + je;
                    d1[0] = h;
                    d1[p.ldy] = __float2bfloat16_rn(y - __bfloat162float(h));
                  }
                }
              }
            } else {
              // res_b[j*M + i] = α·Z (+ β·res), rows r = b*M + i; lanes of a warp write consecutive i: coalesced
              const int b = r / p.M, i = r - b * p.M;
              const size_t off = (size_t)b * p.M * p.N2 + i + (size_t)(n0 + c0) * p.M;
              const int nvalid = Ncols - (n0 + c0);
              if (nvalid > 0) {
                if (p.out_f32) {
                  if (p.beta != 0.f) kron_store_cols<true, true>(p.res, off, p.M, p.alpha, p.beta, v, nvalid);
                  else kron_store_cols<true, false>(p.res, off, p.M, p.alpha, p.beta, v, nvalid);
                } else {
                  if (p.beta != 0.f) kron_store_cols<false, true>(p.res, off, p.M, p.alpha, p.beta, v, nvalid);
                  else kron_store_cols<false, false>(p.res, off, p.M, p.alpha, p.beta, v, nvalid);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (ph == 1 && warp == 2 && lane == 0 && t == (int)blockIdx.x) KR_STAMP(15);
        if (lane == 0) mbar_arrive(&tmem_empty);
        if (warp == 2 && lane == 0 && t == (int)blockIdx.x) KR_STAMP(4 + 4 * ph);
        tphase ^= 1u;
      }
      if (ph == 0) asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (ph == 0) {
      grid_barrier_tight(p.bar, bar_target);   // all of Y is written before any CTA starts streaming it
      bar_target += gridDim.x;
      if (threadIdx.x == 0) KR_STAMP(5);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) KR_STAMP(10);
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
}

// column-major (rows×cols, ld=rows) -> row-major copy with pitch `ldo`
__global__ void kron_transpose_kernel(__nv_bfloat16 *out, const __nv_bfloat16 *in, int rows, int cols, int ldo) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;   // bx: row block, by: col block
  for (int k = threadIdx.y; k < 32; k += 8) {
    int r = bx + threadIdx.x, c = by + k;
    if (r < rows && c < cols) tile[k][threadIdx.x] = in[(size_t)c * rows + r];
  }
  __syncthreads();
  for (int k = threadIdx.y; k < 32; k += 8) {
    int r = bx + k, c = by + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)r * ldo + c] = tile[threadIdx.x][k];
  }
}

// ------------------------------------------------------------------ host
typedef CUresult (*fn_cuTensorMapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static fn_cuTensorMapEncodeTiled g_encode = nullptr;
static int load_encode() {
  if (g_encode) return B2O_OK;
  void *d = dlopen("libcuda.so.1", RTLD_NOW);
  if (!d) B2O_FAIL(B2O_ECUDA, "cannot load libcuda.so.1 (needed for cuTensorMapEncodeTiled)");
  g_encode = (fn_cuTensorMapEncodeTiled)dlsym(d, "cuTensorMapEncodeTiled");
  if (!g_encode) B2O_FAIL(B2O_ECUDA, "cuTensorMapEncodeTiled not found in libcuda");
  return B2O_OK;
}
// bf16 matrix [rows][cols], cols contiguous, pitch `ld` elements; box = 64 cols × box_rows rows, 128-byte swizzle, zero OOB fill
static int make_tmap(CUtensorMap *tm, const void *ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  B2O_TRY(load_encode());
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)KR_BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) B2O_FAIL(B2O_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu", (int)r,
                                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
  return B2O_OK;
}

struct b2o_kron_s {
  b2o_ctx *ctx;
  int m, n, p, q, max_batch;
  const __nv_bfloat16 *A, *B;          // caller's column-major matrices (aliased, like the reference's closures)
  __nv_bfloat16 *Arm = nullptr, *Brm = nullptr;   // row-major copies, pitch padded to 8
  int ldArm, ldBrm;
  __nv_bfloat16 *Y[2] = {nullptr, nullptr};   // [0] prod, [1] tprod workspaces ([hi|lo] rows, zero padded)
  size_t y_elems[2] = {0, 0};
  CUtensorMap tmA1[2], tmB2[2][3], tmY[2][2];   // [direction][BN index: 128, 64, 32]; fixed operands are encoded once at create
  int y_rows[2] = {0, 0};              // rows (nb*M) the cached Y map was encoded for
  CUtensorMap tmX[2];                  // last x map per direction (re-encoded only when x / nb / BN change)
  const void *x_last[2] = {nullptr, nullptr};
  int x_nb[2] = {0, 0}, x_bn[2] = {0, 0};
};

static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }

extern "C" int b2o_kron_create(b2o_ctx *ctx, int dtype, const void *A, int64_t m, int64_t n, const void *B, int64_t p, int64_t q,
                               int max_batch, b2o_kron **out) {
  if (!ctx || !out || !A || !B) B2O_FAIL(B2O_EARG, "null argument");
  if (dtype != B2O_BF16) B2O_FAIL(B2O_EUNSUPPORTED, "kron: the tcgen05 path is bf16 (fp32 accumulate)");
  if (m < 1 || n < 1 || p < 1 || q < 1 || max_batch < 1) B2O_FAIL(B2O_EARG, "bad size");
  if ((m | n | p | q) % 8) B2O_FAIL(B2O_EUNSUPPORTED, "kron: every dimension must be a multiple of 8 (16-byte TMA row pitch)");
  if (((uintptr_t)A | (uintptr_t)B) % 16) B2O_FAIL(B2O_EARG, "kron: matrices must be 16-byte aligned");
  if (m > 1 << 20 || n > 1 << 20 || p > 1 << 20 || q > 1 << 20) B2O_FAIL(B2O_EARG, "kron: dimension too large");
  B2O_CUDA(cudaSetDevice(ctx->device));
  b2o_kron *k = new b2o_kron_s();
  k->ctx = ctx;
  k->m = (int)m; k->n = (int)n; k->p = (int)p; k->q = (int)q;
  k->max_batch = max_batch;
  k->A = (const __nv_bfloat16 *)A;
  k->B = (const __nv_bfloat16 *)B;
  k->ldArm = (int)n;
  k->ldBrm = (int)q;
  k->y_elems[0] = (size_t)max_batch * p * 2 * round_up((int)n, 64);
  k->y_elems[1] = (size_t)max_batch * q * 2 * round_up((int)m, 64);
  cudaError_t e1 = cudaMalloc(&k->Arm, sizeof(__nv_bfloat16) * m * n), e2 = cudaMalloc(&k->Brm, sizeof(__nv_bfloat16) * p * q),
              e3 = cudaMalloc(&k->Y[0], sizeof(__nv_bfloat16) * k->y_elems[0]),
              e4 = cudaMalloc(&k->Y[1], sizeof(__nv_bfloat16) * k->y_elems[1]);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess) {
    cudaGetLastError();
    cudaFree(k->Arm); cudaFree(k->Brm); cudaFree(k->Y[0]); cudaFree(k->Y[1]);
    delete k;
    B2O_FAIL(B2O_ENOMEM, "kron: allocation failed");
  }
  B2O_CUDA(cudaMemsetAsync(k->Y[0], 0, sizeof(__nv_bfloat16) * k->y_elems[0], ctx->stream));
  B2O_CUDA(cudaMemsetAsync(k->Y[1], 0, sizeof(__nv_bfloat16) * k->y_elems[1], ctx->stream));
  dim3 tb(32, 8);
  kron_transpose_kernel<<<dim3((m + 31) / 32, (n + 31) / 32), tb, 0, ctx->stream>>>(k->Arm, k->A, (int)m, (int)n, k->ldArm);
  kron_transpose_kernel<<<dim3((p + 31) / 32, (q + 31) / 32), tb, 0, ctx->stream>>>(k->Brm, k->B, (int)p, (int)q, k->ldBrm);
  ctx->launches += 2;
  B2O_CUDA(cudaGetLastError());
  // prod : A1 = B row-major [p × q], B2 = A row-major [m × n];   tprod: A1 = Bᵀ = [q × p] (B as stored), B2 = Aᵀ = [n × m]
  int st = make_tmap(&k->tmA1[0], k->Brm, p, q, k->ldBrm, KR_BM);
  if (st == B2O_OK) st = make_tmap(&k->tmA1[1], k->B, q, p, p, KR_BM);
  for (int w = 0; w < 3 && st == B2O_OK; ++w) {
    const int bn = 128 >> w;
    st = make_tmap(&k->tmB2[0][w], k->Arm, m, n, k->ldArm, bn);                       // prod : B2 = A row-major [m × n]
    if (st == B2O_OK) st = make_tmap(&k->tmB2[1][w], k->A, n, m, m, bn);              // tprod: B2 = Aᵀ = [n × m], A as stored
  }
  if (st == B2O_OK) {                                                                 // Y maps for the full batch capacity
    const int ldy0 = round_up((int)n, 64), ldy1 = round_up((int)m, 64);
    st = make_tmap(&k->tmY[0][0], k->Y[0], (uint64_t)max_batch * p, 2 * (uint64_t)ldy0, 2 * (uint64_t)ldy0, KR_BM);
    if (st == B2O_OK) st = make_tmap(&k->tmY[1][0], k->Y[1], (uint64_t)max_batch * q, 2 * (uint64_t)ldy1, 2 * (uint64_t)ldy1, KR_BM);
  }
  if (st != B2O_OK) {
    b2o_kron_destroy(k);
    return st;
  }
  *out = k;
  return B2O_OK;
}

extern "C" int b2o_kron_destroy(b2o_kron *k) {
  if (!k) return B2O_OK;
  cudaSetDevice(k->ctx->device);
  cudaStreamSynchronize(k->ctx->stream);
  cudaFree(k->Arm);
  cudaFree(k->Brm);
  cudaFree(k->Y[0]);
  cudaFree(k->Y[1]);
  delete k;
  return B2O_OK;
}

template <int BN>
static int kron_launch(b2o_ctx *c, const CUtensorMap &tA1, const CUtensorMap &tX, const CUtensorMap &tY, const CUtensorMap &tB2,
                       KronArgs &a, int grid) {
  const size_t smem = (size_t)kr_stages(BN) * (KR_BM * KR_BK * 2 + BN * KR_BK * 2) + 1024;
  static thread_local bool configured = false;
  if (!configured) {
    B2O_CUDA(cudaFuncSetAttribute(kron_gemm_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  void *kargs[] = {(void *)&tA1, (void *)&tX, (void *)&tY, (void *)&tB2, (void *)&a};
  if (c->time_kernels) B2O_CUDA(cudaEventRecord(c->ev0, c->stream));
  B2O_CUDA(cudaLaunchCooperativeKernel((const void *)kron_gemm_pair_kernel<BN>, dim3(grid), dim3(KR_THREADS), kargs, smem, c->stream));
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

// trans: 0 prod!, 1 tprod! (== ctprod! for real element types).  x: nb vectors back to back, res likewise.
extern "C" int b2o_kron_apply(b2o_kron *k, int trans, void *res, int res_dtype, int64_t res_len, const void *x, int64_t x_len,
                              int nb, double alpha, double beta) {
  if (!k) B2O_FAIL(B2O_EARG, "null operator");
  if (nb < 1 || nb > k->max_batch) B2O_FAIL(B2O_EARG, "batch %d outside [1, %d]", nb, k->max_batch);
  const int M = trans ? k->q : k->p, K1 = trans ? k->p : k->q, N1 = trans ? k->m : k->n, N2 = trans ? k->n : k->m;
  if (x_len != (int64_t)K1 * N1 || res_len != (int64_t)M * N2) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if (!res || !x) B2O_FAIL(B2O_EARG, "null vector");
  if (res_dtype != B2O_BF16 && res_dtype != B2O_F32) B2O_FAIL(B2O_EUNSUPPORTED, "kron: result dtype must be bf16 or f32");
  if (((uintptr_t)x % 16) || ((uintptr_t)res % 4)) B2O_FAIL(B2O_EARG, "kron: x must be 16-byte aligned");
  b2o_ctx *c = k->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  KronArgs a;
  a.M = M; a.K1 = K1; a.N1 = N1; a.N2 = N2; a.nb = nb;
  a.ldy = round_up(N1, 64);
  a.Y = k->Y[trans ? 1 : 0];
  a.res = res;
  a.out_f32 = res_dtype == B2O_F32;
  a.alpha = (float)alpha;
  a.beta = (float)beta;
  a.bar = c->d_bar;
  a.dbg = c->kron_debug ? (unsigned long long *)(c->d_dots + 448) : nullptr;
  const int t0 = ((M + KR_BM - 1) / KR_BM), t1rows = ((nb * M + KR_BM - 1) / KR_BM);
  // narrower N tiles when the problem has few tiles (latency-bound sizes like 512^3): more CTAs, fewer bytes per CTA
  int w = 0;
  while (w < 2 && (int64_t)t0 * (((int64_t)nb * N1 + (128 >> w) - 1) / (128 >> w)) < c->num_sms / 2) ++w;
  const int BN = 128 >> w;
  const int tiles0 = t0 * ((nb * N1 + BN - 1) / BN), tiles1 = t1rows * ((N2 + BN - 1) / BN);
  int grid = std::max(1, std::min(c->num_sms, std::max(tiles0, tiles1)));
  a.bar_target = c->bar_base + (unsigned long long)grid;
  const int d = trans ? 1 : 0;
  if (k->x_last[d] != x || k->x_nb[d] != nb || k->x_bn[d] != BN) {   // the only per-call descriptor: x (host-side encode, ~1 us)
    B2O_TRY(make_tmap(&k->tmX[d], x, (uint64_t)nb * N1, K1, K1, BN));
    k->x_last[d] = x;
    k->x_nb[d] = nb;
    k->x_bn[d] = BN;
  }
  // Y rows beyond nb*M hold stale (finite) data from larger batches; phase 1 masks its rows with Mrows = nb*M
  const CUtensorMap &tY = k->tmY[d][0], &tB2 = k->tmB2[d][w];
  int st = w == 2   ? kron_launch<32>(c, k->tmA1[d], k->tmX[d], tY, tB2, a, grid)
           : w == 1 ? kron_launch<64>(c, k->tmA1[d], k->tmX[d], tY, tB2, a, grid)
                    : kron_launch<128>(c, k->tmA1[d], k->tmX[d], tY, tB2, a, grid);
  if (st == B2O_OK) c->bar_base += (unsigned long long)grid;
  return st;
}

extern "C" int b2o_kron_flops(b2o_kron *k, int nb, double *flops) {
  if (!k || !flops) B2O_FAIL(B2O_EARG, "null argument");
  *flops = (double)nb * (2.0 * k->p * k->q * k->n + 2.0 * k->p * k->n * k->m);
  return B2O_OK;
}
