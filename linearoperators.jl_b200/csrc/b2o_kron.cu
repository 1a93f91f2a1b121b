// b2o_kron.cu -- kron(A,B)*vec as a TMA-fed tcgen05 GEMM pair (the only tensor-core path).
//
// Reference: src/kron.jl:14-40.  prod!:  X = reshape(x, q, n);  res = α·vec(B·X·Aᵀ) + β·res   (A m×n, B p×q, column-major)
//            tprod!/ctprod!:  X = reshape(x, p, m);  res = α·vec(Bᵀ·X·A) + β·res
// The reference materialises Matrix(B*X*transpose(A)) through n operator applies (3 GEMVs each).  Here it is exactly two
// GEMMs on the 5th-generation tensor cores, in ONE plain (non-cooperative) clustered launch:
//   phase 0   Y[M×N1]  = A1[M×K1] · X'[N1×K1]ᵀ      (fp32 accumulate in TMEM; stored as a bf16 hi/lo pair so that the
//                                                   intermediate costs no accuracy; stays in L2)
//   phase 1   Z[M×N2]  = [Yhi|Ylo][M×2N1] · [B2|B2]ᵀ    epilogue: res[j*M+i] = α·Z[i,j] (+ β·res), bf16 or fp32, column-major
// with every operand K-major (K contiguous) so one code path serves both directions:
//   prod :  M=p K1=q N1=n N2=m   A1 = B row-major (transposed copy made at create), X' = reshape(x,q,n)ᵀ = x as stored,
//           B2 = A row-major (transposed copy made at create)
//   tprod:  M=q K1=p N1=m N2=n   A1 = Bᵀ row-major = B as stored (column-major), X' = x as stored, B2 = Aᵀ row-major = A as stored
//
// Work decomposition (round 2): rows of Z depend only on the same rows of Y, so a 128-row block of one right-hand side is a
// self-contained UNIT handled by one thread-block CLUSTER of C CTAs -- there is no grid-wide dependency, hence no cooperative
// launch and no grid barrier.  CTA c of the cluster owns the N tiles c, c+C, ... of both phases.  The A operand of a phase
// (the 128×K row block of A1, then of [Yhi|Ylo]) is needed by every CTA of the cluster: each CTA fetches 1/C of every k-block
// and TMA-MULTICASTS it into all C shared memories (the per-row-block operand leaves L2 once per cluster, not once per CTA);
// smem slots are handed back cluster-wide by tcgen05.commit ... multicast::cluster.  The intermediate is written with TMA
// stores (UTMASTG) from a swizzled staging tile and published to the cluster through a release/acquire mbarrier in every
// CTA (y_ready); the B operand of phase 1 is prefetched while that hand-over is in flight.  The result leaves through a TMA
// store as well (β = 0; the β ≠ 0 read-modify-write keeps direct stores).
// Per CTA: warp 0 = TMA producer, warp 1 = tcgen05.mma issuer (one elected thread; UMMA 128×BN×16, accumulator in TMEM)
// + TMEM alloc/dealloc, warps 2-5 = epilogue (tcgen05.ld 32x32b → registers → staging → TMA store).
#include "b2o_internal.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <dlfcn.h>
#include <algorithm>

constexpr int KR_BM = 128, KR_BK = 64, KR_THREADS = 192;
constexpr uint32_t KR_A_BYTES = KR_BM * KR_BK * 2;                                   // one 128×64 bf16 k-block of the A operand
constexpr uint32_t KR_SMEM_LIMIT = 227 * 1024;
__host__ __device__ constexpr uint32_t kr_b_bytes(int BN) { return (uint32_t)BN * KR_BK * 2; }
// a ring stage holds TWO (A, B) k-blocks (phase 0: k-blocks 2i, 2i+1; phase 1: Yhi, Ylo k-block i sharing one B2 k-block)
__host__ __device__ constexpr uint32_t kr_stage_bytes(int BN) { return 2 * KR_A_BYTES + 2 * kr_b_bytes(BN); }
// epilogue staging: max(Yhi + Ylo tiles, one fp32 result tile) = 512·BN bytes
__host__ __device__ constexpr uint32_t kr_staging_bytes(int BN) { return 512u * (uint32_t)BN; }
__host__ __device__ constexpr int kr_stages(int BN) {
  return (int)((KR_SMEM_LIMIT - 2048 - kr_staging_bytes(BN)) / kr_stage_bytes(BN)) > 6 ? 6 : (int)((KR_SMEM_LIMIT - 2048 - kr_staging_bytes(BN)) / kr_stage_bytes(BN));
}

struct KronArgs {
  int M, K1, N1, N2, nb;        // see header comment; phase-1 K = N1
  int ldy;                      // N1 rounded up to 128; a Y row holds [hi: ldy | lo: ldy] elements
  int rblocks, units;           // 128-row blocks per right-hand side; units = nb * rblocks
  void *res;                    // nb × (M*N2), each column-major M×N2; bf16 or (out_f32) fp32
  int out_f32, store_tma;       // store_tma: the result tile leaves through a TMA store (beta == 0, 16-byte aligned res)
  float alpha, beta;
  unsigned long long *dbg;      // optional timeline of CTA 0 (%globaltimer, ns): [0] start [1] setup done [2+4*ph] first stage landed
                                // [3+4*ph] accumulator complete [4+4*ph] epilogue done [5] Y published cluster-wide [10] exit
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

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t nclusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (release, cluster scope) on the mbarrier at the same shared-memory offset in CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *smem_dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
               : "memory");
}
// multicast: the box lands at the same CTA-relative offset in every CTA of `mask`, completing bytes on each one's mbarrier
__device__ __forceinline__ void tma_load_2d_mc(void *smem_dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
          smem_u32(smem_dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void *smem_dst, const CUtensorMap *tm, int c0, int c1, int c2, uint64_t *bar, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, %4}], [%5], %6;" ::"r"(
          smem_u32(smem_dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, const void *smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tm), "r"(smem_u32(smem_src)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrives on the mbarrier at this offset in every CTA of `mask` when the MMAs issued so far have read their operands
__device__ __forceinline__ void tc_commit_mc(uint64_t *bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
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

// epilogue helpers: the per-element branches (result type, beta, bounds) are hoisted out of the 32-column loops
template <bool F32, bool BETA>
__device__ __forceinline__ void kron_store_cols(void *res, size_t off, int M, float alpha, float beta, const uint32_t (&v)[32], int nvalid) {
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

template <int BN>
__global__ void __launch_bounds__(KR_THREADS, 1)
kron_cluster_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmX,
                    const __grid_constant__ CUtensorMap tmYld, const __grid_constant__ CUtensorMap tmB2,
                    const __grid_constant__ CUtensorMap tmYhi, const __grid_constant__ CUtensorMap tmYlo,
                    const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ KronArgs p) {
  constexpr uint32_t A_BYTES = KR_A_BYTES, B_BYTES = kr_b_bytes(BN), STAGE_BYTES = kr_stage_bytes(BN);
  constexpr int ST = kr_stages(BN);
  constexpr int YB = BN < 64 ? BN : 64;                  // columns per Y staging box (TMA swizzle span: 64 or 128 bytes)
  constexpr int NYB = BN / YB;                           // Y staging boxes per half
  static_assert(ST >= 2, "ring too shallow");
  // instruction descriptor: D=F32, A=B=BF16, both K-major, N=BN, M=128
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(KR_BM >> 4) << 24);
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);  // SW128 needs 1024 B
  unsigned char *staging = smem + (size_t)ST * STAGE_BYTES;
  __shared__ __align__(8) uint64_t full[ST], empty[ST], tmem_full, tmem_empty, y_ready;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t C = cluster_nctarank(), crank = cluster_ctarank();
  const uint16_t mc_mask = (uint16_t)((1u << C) - 1u);
  const int slice_rows = KR_BM / (int)C;                 // rows of every A k-block this CTA fetches for the whole cluster
  const uint32_t slice_bytes = A_BYTES / C;
  if (threadIdx.x == 0) KR_STAMP(0);
  if (threadIdx.x == 32) {   // hide the descriptor fetches behind the setup
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYld) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmRes) : "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < ST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], C);       // one multicast commit from every CTA of the cluster
    }
    mbar_init(&tmem_full, 1);
    mbar_init(&tmem_empty, 4);
    mbar_init(&y_ready, C);          // one remote arrive from every CTA's Y store
    mbar_fence_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"((uint32_t)(BN < 32 ? 32 : BN)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                // every CTA's barriers exist before any peer multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  if (threadIdx.x == 0) KR_STAMP(1);

  uint32_t stage = 0, sphase = 0;    // smem ring position (producer and MMA warp each keep their own copy)
  uint32_t tphase = 0;               // accumulator hand-over parity (MMA warp and epilogue warps)
  uint32_t yphase = 0;               // y_ready parity (one completion per unit)
  const int kb0 = (p.K1 + KR_BK - 1) / KR_BK, ks0 = (kb0 + 1) / 2;   // phase 0: two k-blocks per stage
  const int ks1 = (p.N1 + KR_BK - 1) / KR_BK;                        // phase 1: one (hi, lo) k-block pair per stage
  const int nt0 = (p.N1 + BN - 1) / BN, nt1 = (p.N2 + BN - 1) / BN;
  const int it0 = (nt0 + (int)C - 1) / (int)C, it1 = (nt1 + (int)C - 1) / (int)C;

  for (int u = (int)cluster_id_x(); u < p.units; u += (int)nclusters_x()) {
    const int b = u / p.rblocks, m0 = (u - b * p.rblocks) * KR_BM;
    const bool first_unit = u == (int)cluster_id_x();
    if (warp == 0) {
      // ===== TMA producer
      if (lane == 0) {
        for (int it = 0; it < it0; ++it) {
          const int tile = it * (int)C + (int)crank;
          const bool valid = tile < nt0;
          for (int ks = 0; ks < ks0; ++ks) {
            const int npair = min(2, kb0 - 2 * ks);
            mbar_wait(&empty[stage], sphase ^ 1u);
            mbar_expect_tx(&full[stage], (uint32_t)npair * (A_BYTES + (valid ? B_BYTES : 0u)));
            unsigned char *sa = smem + (size_t)stage * STAGE_BYTES;
            for (int pr = 0; pr < npair; ++pr) {
              const int kk = (2 * ks + pr) * KR_BK;
              if (valid) tma_load_3d(sa + 2 * A_BYTES + pr * B_BYTES, &tmX, kk, tile * BN, b, &full[stage]);
              if (C > 1) tma_load_2d_mc(sa + pr * A_BYTES + crank * slice_bytes, &tmA1, kk, m0 + (int)crank * slice_rows, &full[stage], mc_mask);
              else tma_load_2d(sa + pr * A_BYTES, &tmA1, kk, m0, &full[stage]);
            }
            if (++stage == ST) { stage = 0; sphase ^= 1u; }
          }
        }
        bool y_waited = false;
        for (int it = 0; it < it1; ++it) {
          const int tile = it * (int)C + (int)crank;
          const bool valid = tile < nt1;
          for (int ks = 0; ks < ks1; ++ks) {
            mbar_wait(&empty[stage], sphase ^ 1u);
            mbar_expect_tx(&full[stage], 2 * A_BYTES + (valid ? B_BYTES : 0u));
            unsigned char *sa = smem + (size_t)stage * STAGE_BYTES;
            if (valid) tma_load_2d(sa + 2 * A_BYTES, &tmB2, ks * KR_BK, tile * BN, &full[stage]);   // does not depend on Y: prefetched
            if (!y_waited) {
              mbar_wait_cluster(&y_ready, yphase);                     // every CTA of the cluster has stored its Y tiles
              asm volatile("fence.proxy.async;" ::: "memory");          // ... and the bulk reads below must observe them
              y_waited = true;
              if (first_unit) KR_STAMP(5);
            }
            const int row = m0 + (int)crank * slice_rows;
            if (C > 1) {
              tma_load_3d_mc(sa + crank * slice_bytes, &tmYld, ks * KR_BK, row, b, &full[stage], mc_mask);
              tma_load_3d_mc(sa + A_BYTES + crank * slice_bytes, &tmYld, p.ldy + ks * KR_BK, row, b, &full[stage], mc_mask);
            } else {
              tma_load_3d(sa, &tmYld, ks * KR_BK, row, b, &full[stage]);
              tma_load_3d(sa + A_BYTES, &tmYld, p.ldy + ks * KR_BK, row, b, &full[stage]);
            }
            if (++stage == ST) { stage = 0; sphase ^= 1u; }
          }
        }
      }
      __syncwarp();
    } else if (warp == 1) {
      // ===== MMA issuer
      for (int ph = 0; ph < 2; ++ph) {
        const int iters = ph == 0 ? it0 : it1, nt = ph == 0 ? nt0 : nt1, ksteps = ph == 0 ? ks0 : ks1;
        for (int it = 0; it < iters; ++it) {
          const bool valid = it * (int)C + (int)crank < nt;
          if (valid) {
            mbar_wait(&tmem_empty, tphase ^ 1u);      // epilogue has drained the accumulator of the previous tile
            tc_fence_after();
          }
          for (int ks = 0; ks < ksteps; ++ks) {
            mbar_wait(&full[stage], sphase);
            tc_fence_after();
            if (lane == 0 && ks == 0 && it == 0 && first_unit) KR_STAMP(2 + 4 * ph);
            if (lane == 0) {
              unsigned char *sa = smem + (size_t)stage * STAGE_BYTES;
              if (valid) {
                const int npair = ph == 0 ? min(2, kb0 - 2 * ks) : 2;
                for (int pr = 0; pr < npair; ++pr) {
                  const uint64_t adesc = umma_desc_sw128(sa + pr * A_BYTES);
                  const uint64_t bdesc = umma_desc_sw128(sa + 2 * A_BYTES + (ph == 0 ? pr * B_BYTES : 0u));
#pragma unroll
                  for (int k = 0; k < KR_BK / 16; ++k)   // UMMA_K = 16 bf16 = 32 B: advance the start address inside the swizzle atom
                    tc_mma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (ks | pr | k) ? 1u : 0u);
                }
              }
              // hand the slot back to EVERY producer of the cluster (their multicasts write into this CTA's copy too)
              if (C > 1) tc_commit_mc(&empty[stage], mc_mask);
              else tc_commit(&empty[stage]);
              if (valid && ks == ksteps - 1) tc_commit(&tmem_full);
            }
            __syncwarp();
            if (++stage == ST) { stage = 0; sphase ^= 1u; }
          }
          if (valid) tphase ^= 1u;
        }
      }
    } else {
      // ===== epilogue: warp w owns TMEM lanes [32*(w%4), +32) = rows of the tile
      const int quarter = warp & 3;
      const int rloc = quarter * 32 + lane;      // row inside the 128-row block
      const bool issuer = threadIdx.x == 64;
      // ---- phase 0: Y tile -> bf16 hi/lo -> swizzled staging -> TMA store
      for (int it = 0; it < it0; ++it) {
        const int tile = it * (int)C + (int)crank;
        if (tile < nt0) {
          mbar_wait(&tmem_full, tphase);
          tc_fence_after();
          if (warp == 2 && lane == 0 && it == 0 && first_unit) KR_STAMP(3);
          if (issuer) tma_store_wait_read();      // the staging tile of the previous store has been read
          epi_sync();
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
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
              // staging box = [128 rows][YB cols] bf16, rows of YB*2 bytes, TMA 128-/64-byte swizzle on the 16-byte chunk index
              const int col = c0 + g, box = col / YB, chunk = (col % YB) / 8;
              const int sw = (YB == 64) ? (chunk ^ (rloc & 7)) : (chunk ^ ((rloc >> 1) & 3));
              unsigned char *dst = staging + (size_t)box * (KR_BM * YB * 2) + (size_t)rloc * (YB * 2) + sw * 16;
              *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4 *>(dst + (size_t)NYB * (KR_BM * YB * 2)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty);
          fence_proxy_async_smem();               // staging writes -> visible to the bulk store
          epi_sync();
          if (issuer) {
#pragma unroll
            for (int bx = 0; bx < NYB; ++bx) {
              tma_store_3d(&tmYhi, staging + (size_t)bx * (KR_BM * YB * 2), tile * BN + bx * YB, m0, b);
              tma_store_3d(&tmYlo, staging + (size_t)(NYB + bx) * (KR_BM * YB * 2), tile * BN + bx * YB, m0, b);
            }
            tma_store_commit();
          }
          tphase ^= 1u;
        }
      }
      if (issuer) {
        // publish: all Y tiles of this CTA are in global memory -> release-arrive on y_ready of every CTA of the cluster
        tma_store_wait_all();
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence();
        for (uint32_t r = 0; r < C; ++r) mbar_arrive_remote(&y_ready, r);
        if (first_unit) KR_STAMP(4);
      }
      // ---- phase 1: result tile
      for (int it = 0; it < it1; ++it) {
        const int tile = it * (int)C + (int)crank;
        if (tile < nt1) {
          const int n0 = tile * BN;
          mbar_wait(&tmem_full, tphase);
          tc_fence_after();
          if (warp == 2 && lane == 0 && it == 0 && first_unit) KR_STAMP(7);
          if (p.store_tma) {
            if (issuer) tma_store_wait_read();
            epi_sync();
          }
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tc_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            if (p.store_tma) {
              // staging tile [BN cols (j)][128 rows (i)]: lanes write consecutive i -> conflict-free, no swizzle needed
              if (p.out_f32) {
                float *st = reinterpret_cast<float *>(staging) + (size_t)c0 * KR_BM + rloc;
#pragma unroll
                for (int e = 0; e < 32; ++e) st[(size_t)e * KR_BM] = p.alpha * __uint_as_float(v[e]);
              } else {
                __nv_bfloat16 *st = reinterpret_cast<__nv_bfloat16 *>(staging) + (size_t)c0 * KR_BM + rloc;
#pragma unroll
                for (int e = 0; e < 32; ++e) st[(size_t)e * KR_BM] = __float2bfloat16_rn(p.alpha * __uint_as_float(v[e]));
              }
            } else if (m0 + rloc < p.M) {
              // res_b[j*M + i] = α·Z (+ β·res); lanes of a warp write consecutive i: coalesced
              const size_t off = (size_t)b * p.M * p.N2 + (size_t)(m0 + rloc) + (size_t)(n0 + c0) * p.M;
              const int nvalid = p.N2 - (n0 + c0);
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
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty);
          if (p.store_tma) {
            fence_proxy_async_smem();
            epi_sync();
            if (issuer) {
              tma_store_3d(&tmRes, staging, m0, n0, b);
              tma_store_commit();
            }
          }
          if (warp == 2 && lane == 0 && it == 0 && first_unit) KR_STAMP(8);
          tphase ^= 1u;
        }
      }
    }
    yphase ^= 1u;
  }
  if (threadIdx.x == 64) tma_store_wait_all();      // bulk stores must have left shared memory before the CTA exits
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // no peer may still multicast into / arrive on this CTA after it exits
  if (threadIdx.x == 0) KR_STAMP(10);
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(BN < 32 ? 32 : BN)) : "memory");
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
