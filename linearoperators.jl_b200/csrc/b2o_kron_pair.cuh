// b2o_kron_pair.cuh -- kron(A,B)*vec for MANY units (batched right-hand sides): the same GEMM pair as kron_cluster_kernel
// (src/kron.jl:14-40, see b2o_kron.cu), issued as `tcgen05.mma.cta_group::2` by a CTA PAIR (the two SMs of one TPC).
//
// Why a second kernel: an SM takes operand bytes in at ~60 B/clk (profiles/r2_kron_timeline.md).  A 128x128 single-CTA tile
// needs (128 + 128) x 2 B of operands per 2*128*128 flop = 125 B/clk at tensor-pipe speed -> the single-CTA kernel is
// intake-bound at ~0.4 of the pipe.  A 256x256 pair tile needs the same (128 + 128) rows per CTA for TWICE the flops:
// each CTA loads its own 128 rows of the A operand and its own 128 of the 256 rows of the B operand, the pair's tensor cores
// read the B operand halves from both shared memories -> 62 B/clk.
//
//   unit     = 256 rows of Y / Z of ONE right-hand side; cluster = 1 pair = 1 unit at a time, persistent over units
//   CTA c    owns rows [256u + 128c, +128) of the unit: it stores exactly the Y rows it later reads back as the A operand of
//              the second GEMM -> no cross-CTA hand-over of Y (a local mbarrier after the bulk stores completed)
//   ring     = 12 slots of 16 KB (one 128-row x 64-col bf16 block each, 128-byte swizzle); item of GEMM 1 = (A1, X') blocks
//              = 2 slots, item of GEMM 2 = (B2, Yhi, Ylo) blocks = 3 slots, per-slot full/empty mbarriers.  Both CTAs' loads
//              of a slot complete on the LEADER's full barrier (cp.async.bulk.tensor ... .cta_group::2), the leader's MMA
//              warp hands a slot back to both producers with tcgen05.commit.cta_group::2 ... multicast::cluster.
//   TMEM     = two 256-column fp32 accumulators (all 512 columns): the MMAs of a tile overlap the epilogue of the previous
//   epilogue = 32-column chunks through two 16 KB staging buffers -> TMA stores (Y as a bf16 hi/lo pair, the result as bf16 /
//              fp32); the peer's epilogue warps release an accumulator with a remote arrive on the leader's barrier.
// Units of a cluster are software-pipelined like in kron_cluster_kernel: GEMM 1 of unit i+1 runs before GEMM 2 of unit i.
// Measured (profiles/r2_kron_pair.jsonl, r2_kron_pair_timeline_v*.jsonl): 64 right-hand sides 49.5 us = 694 TFLOP/s algorithmic, 1041
// issued (single-CTA kernel: 80 us; first version of this kernel 54 us, tensor pipe 55 % of active cycles); 296: 769 TFLOP/s.  What bounds it now: with all 148 SMs streaming operands the chip
// sits at the L2 -> SM ceiling (a GEMM 1 tile takes 3.7 us for 256 KB per CTA = 10.2 TB/s chip-wide, a GEMM 2 tile 5.2 us for
// 384 KB = 10.9 TB/s; ~6300 B/clk x 1.7 GHz), i.e. ~19.5 us of MMA issue per unit against 12.9 us at the nominal tensor rate, and
// the epilogue of a tile (8 chunks x ~0.55 us of TMEM load, conversion, proxy fence, barrier, store issue) is as long as the MMAs
// of the next tile, so the two-accumulator hand-over adds ~7 us per unit.  Tried and kept as evidence: plain stores instead of
// TMA stores (option pair_tma_stores = 0: 76 us), two epilogue warp groups draining alternate chunks (56 us: the per-chunk latency
// rises as much as the concurrency gains -- profiles/r2_kron_pair_two_epilogue_groups*.jsonl).  Next lever: multicast the shared
// B operand across the two pairs of one right-hand side (-20 % L2 reads).
// Every wait carries a %globaltimer watchdog (trap after 4 s): a protocol error fails loudly instead of hanging the GPU.
#pragma once

constexpr int KP_SLOT = 16384, KP_SLOTS = 12, KP_STG = 16384, KP_ROWS = 128, KP_TN = 256;
constexpr uint32_t KP_TCOLS = 256;
constexpr size_t KP_SMEM = 1024 + (size_t)KP_SLOTS * KP_SLOT + 2 * (size_t)KP_STG;

__device__ __forceinline__ uint32_t kp_mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
// CLUSTER = true only where the arrivals come from the peer CTA (tmem_empty): an acquire at cluster scope makes ptxas put a
// CCTL.IVALL (L1 invalidate) into every polling iteration -- 10 % of the kernel's warp samples when every wait had it
// (ncu source page, profiles/r2_ncu_summary.md)
template <bool CLUSTER = false>
__device__ __forceinline__ void kp_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done = 0, it = 0;
  unsigned long long t0 = 0;
  for (;;) {
    if (CLUSTER)
      asm volatile(
          "{\n"
          ".reg .pred P1;\n"
          "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n"
          "selp.u32 %0, 1, 0, P1;\n"
          "}\n"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity)
          : "memory");
    else
      asm volatile(
          "{\n"
          ".reg .pred P1;\n"
          "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
          "selp.u32 %0, 1, 0, P1;\n"
          "}\n"
          : "=r"(done)
          : "r"(smem_u32(bar)), "r"(parity)
          : "memory");
    if (done) break;
    if ((++it & 1023u) == 0) {
      const unsigned long long t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 4000000000ULL) asm volatile("trap;");
    }
  }
}
// both CTAs of the pair load into their OWN shared memory and complete the bytes on the barrier at cluster address `bar_cl`
__device__ __forceinline__ void kp_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint32_t bar_cl) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(bar_cl)
               : "memory");
}
__device__ __forceinline__ void kp_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, uint32_t bar_cl) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                   smem_u32(dst)),
               "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar_cl)
               : "memory");
}
__device__ __forceinline__ void kp_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrives on the barrier at this offset in BOTH CTAs of the pair once the MMAs issued so far have completed
__device__ __forceinline__ void kp_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// The arrive that hands an accumulator back orders nothing in memory (the TMEM reads are complete: tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync precede it, the MMA warp issues tcgen05.fence::after_thread_sync behind its wait): relaxed.
// A release at cluster scope costs a MEMBAR + ERRBAR per warp and tile (10 % of the warp samples).
__device__ __forceinline__ void kp_arrive_cl(uint32_t bar_cl) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cl) : "memory");
}
__device__ __forceinline__ void kp_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

#define KP_ADVANCE(n)                                \
  do {                                               \
    slot += (n);                                     \
    if (slot >= (uint32_t)KP_SLOTS) {                \
      slot -= (uint32_t)KP_SLOTS;                    \
      sphase ^= 1u;                                  \
    }                                                \
  } while (0)

// DBG: the %globaltimer stamps of tools/kron_pair_dbg.py are compiled in only for the debug instantiation (their predicated-off
// LDC / CS2R / STG still took 15 % of the epilogue's issue slots in the production kernel)
template <bool DBG>
__global__ void __launch_bounds__(KR_THREADS, 1)
kron_pair_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmYld, const __grid_constant__ CUtensorMap tmB2,
                 const __grid_constant__ CUtensorMap tmYhi, const __grid_constant__ CUtensorMap tmYlo,
                 const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ KronArgs p) {
  // instruction descriptor: D = F32, A = B = BF16, both K-major, N = 256, M = 256 (128 rows per CTA)
  constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP_TN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  extern __shared__ unsigned char smem_raw[];
  unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char *staging = smem + (size_t)KP_SLOTS * KP_SLOT;
  __shared__ __align__(8) uint64_t full[KP_SLOTS], empty[KP_SLOTS], tmem_full[2], tmem_empty[2], y_ready[2];
  __shared__ uint32_t s_tmem;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  const uint32_t crank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const int cid = __shfl_sync(0xffffffffu, (int)cluster_id_x(), 0), ncl = __shfl_sync(0xffffffffu, (int)nclusters_x(), 0);
  const bool leader = crank == 0;
  if (DBG && threadIdx.x == 0 && p.dbg && blockIdx.x == 0) p.dbg[0] = gtimer();
  if (threadIdx.x == 32) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYld) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYhi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmYlo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmRes) : "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < KP_SLOTS; ++s) {
      mbar_init(&full[s], 1);        // the leader's arrive.expect_tx; the bytes of both CTAs' loads complete here (leader's copy only)
      mbar_init(&empty[s], 1);       // one multicast commit from the leader's MMA warp
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);   // multicast commit
      mbar_init(&tmem_empty[a], 8);  // 4 epilogue warps of each CTA (leader's copy only)
      mbar_init(&y_ready[a], 1);     // this CTA's own Y rows are in global memory
    }
    mbar_fence_init();
  }
  __syncthreads();
  cluster_sync_all();                // both CTAs' barriers exist before the peer completes bytes / arrives on them
  if (warp == 1) {                   // the same warp of both CTAs allocates collectively
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(2u * KP_TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                // the peer's columns are allocated before the leader's first MMA writes them
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  uint32_t slot = 0, sphase = 0;     // ring position (producer and MMA warp keep their own copies)
  uint32_t tcount = 0;               // accumulator tiles handed over so far (MMA warp and epilogue warps count alike)
  uint32_t gcount = 0;               // epilogue: staging chunks stored so far (buffer gcount & 1)
  const int kb0 = (p.K1 + KR_BK - 1) / KR_BK;                       // GEMM 1: k-blocks
  const int ks1 = (p.N1 + KR_BK - 1) / KR_BK;                       // GEMM 2: (hi, lo) k-block pairs
  const int nt0 = (p.N1 + KP_TN - 1) / KP_TN, nt1 = (p.N2 + KP_TN - 1) / KP_TN;
  const int nu = cid < p.units ? (p.units - 1 - cid) / ncl + 1 : 0;

  for (int i = 0; i <= nu; ++i) {
    const bool do0 = i < nu, do1 = i >= 1;
    const int u0 = cid + i * ncl, u1 = cid + (i - 1) * ncl;
    const int b0 = do0 ? u0 / p.rblocks : 0, row0 = do0 ? (u0 - b0 * p.rblocks) * 256 + (int)crank * KP_ROWS : 0;   // this CTA's rows, GEMM 1
    const int b1 = do1 ? u1 / p.rblocks : 0, row1 = do1 ? (u1 - b1 * p.rblocks) * 256 + (int)crank * KP_ROWS : 0;   // this CTA's rows, GEMM 2
    uint64_t *yr1 = &y_ready[(i - 1) & 1];
    const uint32_t yphase1 = (uint32_t)((i - 1) >> 1) & 1u;
    if (warp == 0) {
      // ===== TMA producer (both CTAs): one elected lane per slot; the leader also arms the slot's full barrier for both CTAs' bytes
      if (do0) {
        for (int t = 0; t < nt0; ++t) {
          for (int kb = 0; kb < kb0; ++kb) {
            kp_wait(&empty[slot], sphase ^ 1u);
            if (elect_one()) {
              if (leader) mbar_expect_tx(&full[slot], 2u * KP_SLOT);
              kp_load_2d(smem + (size_t)slot * KP_SLOT, &tmA1, kb * KR_BK, row0, kp_mapa(smem_u32(&full[slot]), 0));
            }
            __syncwarp();
            KP_ADVANCE(1);
            kp_wait(&empty[slot], sphase ^ 1u);
            if (elect_one()) {
              if (leader) mbar_expect_tx(&full[slot], 2u * KP_SLOT);
              kp_load_3d(smem + (size_t)slot * KP_SLOT, &tmX, kb * KR_BK, t * KP_TN + (int)crank * KP_ROWS, b0, kp_mapa(smem_u32(&full[slot]), 0));
            }
            __syncwarp();
            KP_ADVANCE(1);
          }
        }
      }
      if (do1) {
        bool y_waited = false;
        for (int t = 0; t < nt1; ++t) {
          for (int ks = 0; ks < ks1; ++ks) {
            kp_wait(&empty[slot], sphase ^ 1u);
            if (elect_one()) {                       // rows of B2 do not depend on Y: in flight while the Y stores drain
              if (leader) mbar_expect_tx(&full[slot], 2u * KP_SLOT);
              kp_load_2d(smem + (size_t)slot * KP_SLOT, &tmB2, ks * KR_BK, t * KP_TN + (int)crank * KP_ROWS, kp_mapa(smem_u32(&full[slot]), 0));
            }
            __syncwarp();
            KP_ADVANCE(1);
            if (!y_waited) {
              kp_wait(yr1, yphase1);                                    // this CTA's Y rows of the unit are stored
              asm volatile("fence.proxy.async;" ::: "memory");
              y_waited = true;
            }
            kp_wait(&empty[slot], sphase ^ 1u);
            if (elect_one()) {
              if (leader) mbar_expect_tx(&full[slot], 2u * KP_SLOT);
              kp_load_3d(smem + (size_t)slot * KP_SLOT, &tmYld, ks * KR_BK, row1, b1, kp_mapa(smem_u32(&full[slot]), 0));
            }
            __syncwarp();
            KP_ADVANCE(1);
            kp_wait(&empty[slot], sphase ^ 1u);
            if (elect_one()) {
              if (leader) mbar_expect_tx(&full[slot], 2u * KP_SLOT);
              kp_load_3d(smem + (size_t)slot * KP_SLOT, &tmYld, p.ldy + ks * KR_BK, row1, b1, kp_mapa(smem_u32(&full[slot]), 0));
            }
            __syncwarp();
            KP_ADVANCE(1);
          }
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer: the leader CTA issues for the pair
      if (leader) {
        for (int ph = 0; ph < 2; ++ph) {
          if (ph == 0 ? !do0 : !do1) continue;
          const int nt = ph == 0 ? nt0 : nt1, ksteps = ph == 0 ? kb0 : ks1;
          for (int t = 0; t < nt; ++t) {
            const uint32_t acc = tcount & 1u, tpar = (tcount >> 1) & 1u;
            const uint32_t tmem_acc = tmem_base + acc * KP_TCOLS;
            kp_wait<true>(&tmem_empty[acc], tpar ^ 1u); // both CTAs' epilogues have drained this accumulator (remote arrives)
            tc_fence_after();
            if (DBG && p.dbg && p.dbg_tile < 0 && blockIdx.x == 0 && lane == 0 && tcount < 8) p.dbg[16 + 2 * tcount] = gtimer();
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint32_t s0 = slot, p0 = sphase;
              KP_ADVANCE(1);
              const uint32_t s1 = slot, p1 = sphase;
              KP_ADVANCE(1);
              uint32_t s2 = 0, p2 = 0;
              if (ph == 1) {
                s2 = slot;
                p2 = sphase;
                KP_ADVANCE(1);
              }
              kp_wait(&full[s0], p0);
              kp_wait(&full[s1], p1);
              if (ph == 1) kp_wait(&full[s2], p2);
              tc_fence_after();
              if (elect_one()) {
                if (ph == 0) {
                  const uint64_t adesc = umma_desc_sw128(smem + (size_t)s0 * KP_SLOT), bdesc = umma_desc_sw128(smem + (size_t)s1 * KP_SLOT);
#pragma unroll
                  for (int k = 0; k < KR_BK / 16; ++k) kp_mma(tmem_acc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (ks | k) ? 1u : 0u);
                  kp_commit(&empty[s0]);
                  kp_commit(&empty[s1]);
                } else {
                  const uint64_t bdesc = umma_desc_sw128(smem + (size_t)s0 * KP_SLOT);
                  const uint64_t hdesc = umma_desc_sw128(smem + (size_t)s1 * KP_SLOT), ldesc = umma_desc_sw128(smem + (size_t)s2 * KP_SLOT);
#pragma unroll
                  for (int k = 0; k < KR_BK / 16; ++k) kp_mma(tmem_acc, hdesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (ks | k) ? 1u : 0u);
#pragma unroll
                  for (int k = 0; k < KR_BK / 16; ++k) kp_mma(tmem_acc, ldesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, 1u);
                  kp_commit(&empty[s0]);
                  kp_commit(&empty[s1]);
                  kp_commit(&empty[s2]);
                }
                if (ks == ksteps - 1) kp_commit(&tmem_full[acc]);
              }
              __syncwarp();
            }
            if (DBG && p.dbg && p.dbg_tile < 0 && blockIdx.x == 0 && lane == 0 && tcount < 8) p.dbg[17 + 2 * tcount] = gtimer();
            ++tcount;
          }
        }
      }
    } else {
      // ===== epilogue (both CTAs): warp w owns TMEM lanes [32*(w%4), +32) = rows of this CTA's half of the unit
      const int quarter = warp & 3;
      const int rloc = quarter * 32 + lane;
      const bool issuer = threadIdx.x == 64;
      const uint32_t te_leader = kp_mapa(smem_u32(&tmem_empty[0]), 0);
      // ---- GEMM 1: Y tile -> bf16 hi/lo -> swizzled staging -> TMA store, 32 columns at a time
      if (do0) {
        for (int t = 0; t < nt0; ++t) {
          const uint32_t acc = tcount & 1u, tpar = (tcount >> 1) & 1u;
          const uint32_t tmem_acc = tmem_base + acc * KP_TCOLS;
          kp_wait(&tmem_full[acc], tpar);
          tc_fence_after();
          if (DBG && p.dbg && p.dbg_tile < 0 && blockIdx.x == 0 && threadIdx.x == 64 && tcount < 8) p.dbg[48 + 2 * tcount] = gtimer();
#pragma unroll 1
          for (int c0 = 0; c0 < KP_TN; c0 += 32) {
            const int col0 = t * KP_TN + c0;
            if (col0 >= p.N1) break;
            uint32_t v[32];
            const bool stamp = DBG && p.dbg && p.dbg_tile == (int)tcount && blockIdx.x == 0 && threadIdx.x == 64;
            if (stamp) p.dbg[16 + 5 * (c0 >> 5)] = gtimer();
            tc_ld32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            if (stamp) p.dbg[17 + 5 * (c0 >> 5)] = gtimer();
            unsigned char *stg = staging + (size_t)(gcount & 1u) * KP_STG;
            if (p.y_tma) {
              if (issuer) kp_store_wait_read1();     // the store that used this buffer two chunks ago has read it
              epi_sync();
            }
            if (stamp) p.dbg[18 + 5 * (c0 >> 5)] = gtimer();
            // (manual path: the copy-out of the chunk two back was finished by every thread before it passed the epi_sync of
            //  the previous chunk, so this buffer is free)
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
              // box = [128 rows][32 cols] bf16: rows of 64 bytes, 16-byte chunk index xor-ed with (row / 2) % 4 (= TMA's 64-byte
              // swizzle): the 32 lanes of a warp write 32 rows conflict-free
              const int sw = (g >> 3) ^ ((rloc >> 1) & 3);
              unsigned char *dst = stg + (size_t)rloc * 64 + sw * 16;
              *reinterpret_cast<uint4 *>(dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4 *>(dst + KP_ROWS * 64) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            if (stamp) p.dbg[19 + 5 * (c0 >> 5)] = gtimer();
            if (p.y_tma) {
              fence_proxy_async_smem();
              epi_sync();
              if (issuer) {
                if (row0 < p.M) {
                  tma_store_3d(&tmYhi, stg, col0, row0, b0);
                  tma_store_3d(&tmYlo, stg + KP_ROWS * 64, col0, row0, b0);
                }
                tma_store_commit();
              }
            } else {
              // copy-out with plain 16-byte stores: 4 consecutive threads cover the 64 bytes of one row of the box (option
              // pair_tma_stores = 0; measured slower than the TMA stores: 76 vs 54 us at 64 right-hand sides)
              epi_sync();
              const int te = (int)threadIdx.x - 64;
              __nv_bfloat16 *Yb = reinterpret_cast<__nv_bfloat16 *>(p.y) + (size_t)b0 * p.M * 2 * p.ldy;
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int u = te + 128 * k, half = u >> 9, w = u & 511, row = w >> 2, pc = w & 3;
                const int gcol = col0 + ((pc ^ ((row >> 1) & 3)) << 3), grow = row0 + row;
                if (grow < p.M && gcol < p.N1)
                  *reinterpret_cast<uint4 *>(Yb + (size_t)grow * 2 * p.ldy + (size_t)half * p.ldy + gcol) = *reinterpret_cast<const uint4 *>(stg + (size_t)u * 16);
              }
            }
            if (stamp) p.dbg[20 + 5 * (c0 >> 5)] = gtimer();
            ++gcount;
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) kp_arrive_cl(te_leader + acc * 8u);
          if (DBG && p.dbg && p.dbg_tile < 0 && blockIdx.x == 0 && threadIdx.x == 64 && tcount < 8) p.dbg[49 + 2 * tcount] = gtimer();
          ++tcount;
        }
        if (p.y_tma) {
          if (issuer) {
            tma_store_wait_all();                       // this CTA's Y rows of the unit are in global memory
            asm volatile("fence.proxy.async;" ::: "memory");
            mbar_arrive(&y_ready[i & 1]);
          }
        } else {
          __threadfence();                              // every thread's Y stores are performed ...
          asm volatile("fence.proxy.async;" ::: "memory");   // ... and ordered before the bulk (async-proxy) loads that follow the hand-over
          epi_sync();
          if (issuer) mbar_arrive(&y_ready[i & 1]);
        }
      }
      // ---- GEMM 2: result tile, 32 columns (j) at a time
      if (do1) {
        for (int t = 0; t < nt1; ++t) {
          const uint32_t acc = tcount & 1u, tpar = (tcount >> 1) & 1u;
          const uint32_t tmem_acc = tmem_base + acc * KP_TCOLS;
          kp_wait(&tmem_full[acc], tpar);
          tc_fence_after();
          if (DBG && p.dbg && p.dbg_tile < 0 && blockIdx.x == 0 && threadIdx.x == 64 && tcount < 8) p.dbg[48 + 2 * tcount] = gtimer();
#pragma unroll 1
          for (int c0 = 0; c0 < KP_TN; c0 += 32) {
            const int n0 = t * KP_TN + c0;
            if (n0 >= p.N2) break;
            uint32_t v[32];
            const bool stamp = DBG && p.dbg && p.dbg_tile == (int)tcount && blockIdx.x == 0 && threadIdx.x == 64;
            if (stamp) p.dbg[16 + 5 * (c0 >> 5)] = gtimer();
            tc_ld32(tmem_acc + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
            if (stamp) p.dbg[17 + 5 * (c0 >> 5)] = gtimer();
            if (p.store_tma) {
              unsigned char *stg = staging + (size_t)(gcount & 1u) * KP_STG;
              if (issuer) kp_store_wait_read1();
              epi_sync();
              if (stamp) p.dbg[18 + 5 * (c0 >> 5)] = gtimer();
              // staging tile [32 cols (j)][128 rows (i)]: lanes write consecutive i -> conflict-free
              if (p.out_f32) {
                float *st = reinterpret_cast<float *>(stg) + rloc;
#pragma unroll
                for (int e = 0; e < 32; ++e) st[(size_t)e * KP_ROWS] = p.alpha * __uint_as_float(v[e]);
              } else {
                __nv_bfloat16 *st = reinterpret_cast<__nv_bfloat16 *>(stg) + rloc;
#pragma unroll
                for (int e = 0; e < 32; ++e) st[(size_t)e * KP_ROWS] = __float2bfloat16_rn(p.alpha * __uint_as_float(v[e]));
              }
              if (stamp) p.dbg[19 + 5 * (c0 >> 5)] = gtimer();
              fence_proxy_async_smem();
              epi_sync();
              if (issuer) {
                if (row1 < p.M) tma_store_3d(&tmRes, stg, row1, n0, b1);
                tma_store_commit();
              }
              if (stamp) p.dbg[20 + 5 * (c0 >> 5)] = gtimer();
              ++gcount;
            } else if (row1 + rloc < p.M) {
              const size_t off = (size_t)b1 * p.M * p.N2 + (size_t)(row1 + rloc) + (size_t)n0 * p.M;
              const int nvalid = p.N2 - n0;
              if (p.out_f32) {
                if (p.beta != 0.f) kron_store_cols<true, true>(p.res, off, p.M, p.alpha, p.beta, v, nvalid);
                else kron_store_cols<true, false>(p.res, off, p.M, p.alpha, p.beta, v, nvalid);
              } else {
                if (p.beta != 0.f) kron_store_cols<false, true>(p.res, off, p.M, p.alpha, p.beta, v, nvalid);
                else kron_store_cols<false, false>(p.res, off, p.M, p.alpha, p.beta, v, nvalid);
              }
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) kp_arrive_cl(te_leader + acc * 8u);
          if (DBG && p.dbg && p.dbg_tile < 0 && blockIdx.x == 0 && threadIdx.x == 64 && tcount < 8) p.dbg[49 + 2 * tcount] = gtimer();
          ++tcount;
        }
      }
    }
  }
  if (threadIdx.x == 64) tma_store_wait_all();      // bulk stores must have left shared memory before the CTA exits
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                               // the leader's MMAs into the peer's TMEM / arrives on the peer are all done
  if (DBG && threadIdx.x == 0 && p.dbg && blockIdx.x == 0) p.dbg[10] = gtimer();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * KP_TCOLS) : "memory");
}
#undef KP_ADVANCE
