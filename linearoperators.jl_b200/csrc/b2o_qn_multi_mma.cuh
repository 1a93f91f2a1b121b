// b2o_qn_multi_mma.cuh -- block apply for 5..8 right-hand sides on the FP64 tensor cores (mma.sync m8n8k4, SASS DMMA).
//
// Same operator and the same two streaming phases as qn_multi_kernel<8, OP> (mul!(Res::Matrix, op, X::Matrix, α, β),
// src/operations.jl:34-36, per column src/lbfgs.jl:173-202 / src/lsr1.jl:89-107), but both phases are written as the small
// GEMMs they are, G = colsᵀ X (ncols x 8, contraction over the rows) and Res = base(X) + cols * coef (rows x 8, contraction
// over the columns):
//   phase 1: a warp owns 128 rows of the 1024-row tile.  Its X fragments (32 k-steps of 4 rows x 8 right-hand sides, one double
//            per lane each) are loaded once per tile and stay in registers; for every group of 8 columns the A fragment is one
//            LDS.64 per lane straight from the TMA ring (8 columns x 4 rows) and ONE DMMA accumulates 256 products into the
//            lane's two accumulators -- kept in registers for the whole kernel.  No shuffles, no shared-memory read-modify-write
//            per column tile (the SIMT kernel spent more issue slots folding partials than multiplying: 14 SHFL + 7 DADD + an
//            smem RMW per 32 DFMA; profiles/r1_ncu_multi_summary.md).
//   phase 2: 16 row groups of 8 rows per warp, accumulators C[16][2] in registers initialised with the base term, one LDS.64 +
//            one DMMA per row group and 4 columns, coefficients as the B fragment.
// DMMA runs at the FP64 pipe's full rate on B200 (37 TFLOP/s measured, tools/micro/dmma_rate.cu); the kernel needs ~25 % of it.
// STATUS: opt-in (ctx option "multi_mma" = 1).  Measured at n = 1e8, m = 10: 8 right-hand sides 11.3 ms against 10.5 ms of the SIMT
// kernel (first version, 1024-row tiles, one dependent chain of 32 DMMAs per group, no X prefetch: 14.8 ms) -- removing the
// shuffle folds did not remove the stall, so the SIMT kernel's limiter is not its instruction count; kept as the measured negative
// result and as the correctness-tested DMMA path (tests/test_gpu_parity.py::test_block_apply_dmma_kernel).
// NOTE (cost one GPU lease): `mma.sync.aligned` after an mbarrier polling loop needs an explicit __syncwarp() -- lanes leave the
// loop at different iterations and an aligned MMA issued by a diverged warp hangs.
// Ring: 16 slots of one 1024-row column tile; slot s is skewed by 32·(s mod 4) bytes so that the 8 (phase 1) / 4 (phase 2)
// tiles a fragment load touches -- consecutive slots -- fall into different bank groups (2 wavefronts per LDS.64, the minimum).
// Rounding: like qn_multi_kernel the arithmetic is contracted (tensor-core FMAs, fixed order: rows of a warp in k-step order,
// warps and CTAs in index order -> deterministic); per-column results differ from the vector kernel by a few ulp.
#pragma once
#include "b2o_qn_multi.cuh"

// 512-row tiles: 64 rows = 16 k-steps (phase 1) / 8 row groups (phase 2) per warp keep the X fragments of the current AND the
// prefetched next tile in registers (the kernel is capped at 168 registers: 9 warps, 3 on one scheduler).
constexpr int MM_R = 512, MM_STAGES = 32, MM_SLOT = MM_R * 8 + 128, MM_MAXG = 4;   // up to 32 columns (B2O_MULTI_MAXV / 8)
constexpr int MM_WROWS = MM_R / 8, MM_KS = MM_WROWS / 4, MM_RG = MM_WROWS / 8;

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
// mbarrier wait with a %globaltimer watchdog: a protocol error reports where it is stuck and traps instead of hanging the GPU
__device__ __forceinline__ void mm_wait(uint64_t *bar, uint32_t parity, int code) {
  uint32_t done = 0, it = 0;
  unsigned long long t0 = 0;
  for (;;) {
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
    if ((++it & 63u) == 0) {
      const unsigned long long t = globaltimer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > 2000000000ULL) {
        printf("qn_multi_mma_kernel: wait %d stuck (cta %d thread %d parity %u)\n", code, (int)blockIdx.x, (int)threadIdx.x, parity);
        asm volatile("trap;");
      }
    }
  }
}
__device__ __forceinline__ unsigned char *mm_slot_ptr(unsigned char *ring, uint32_t slot) { return ring + (size_t)slot * MM_SLOT + 32u * (slot & 3u); }

template <int OP>
__global__ void __launch_bounds__(B2O_NTHREADS, 1) qn_multi_mma_kernel(const __grid_constant__ MultiArgs p) {
  constexpr int NR = 8, S = MM_STAGES;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char *ring = smem_raw;
  double *wacc = reinterpret_cast<double *>(smem_raw + p.wacc_off);   // [8 warps][ncols][8]
  double *coef = reinterpret_cast<double *>(smem_raw + p.coef_off);   // [ncols][8]
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + p.bar_off);
  uint64_t *empty = full + S;
  unsigned *s_landed = reinterpret_cast<unsigned *>(smem_raw + p.landed_off);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_producer = warp == B2O_CONS_WARPS;
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], B2O_CONS_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  uint32_t slot = 0, par = 0;   // ring position (producer: next slot to fill; consumers: next slot to read)
  const int64_t grid = gridDim.x;
  const int64_t my_tiles = (p.ntiles > (int64_t)blockIdx.x) ? (p.ntiles - 1 - blockIdx.x) / grid + 1 : 0;
  const int ncols = p.ncols, nrhs = p.nrhs;
  const int nv = ncols * NR;
  const int lk = lane & 3, lm = lane >> 2;

  auto push = [&](const double *src) {
    mm_wait(&empty[slot], par ^ 1u, 100 + (int)slot);
    mbar_expect_tx(&full[slot], (uint32_t)(MM_R * sizeof(double)));
    bulk_g2s(mm_slot_ptr(ring, slot), src, (uint32_t)(MM_R * sizeof(double)), &full[slot]);
    if (++slot == (uint32_t)S) {
      slot = 0;
      par ^= 1u;
    }
  };

  // ------------------------------------------------------------------ phase 1: G = colsᵀ X
  if (is_producer) {
    if (lane == 0) {
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int64_t t = blockIdx.x + i * grid;
        for (int c = 0; c < ncols; ++c) push(p.cols[c] + t * MM_R);
      }
    }
    __syncwarp();
  } else {
    // two accumulator pairs per column group (even / odd k-steps): dependent DMMA chains of 16 instead of 32
    double C[MM_MAXG][2][2];
#pragma unroll
    for (int g = 0; g < MM_MAXG; ++g) C[g][0][0] = C[g][0][1] = C[g][1][0] = C[g][1][1] = 0.0;
    // B fragments of the warp's 32 k-steps: lane (k = lane & 3, n = lane >> 2) holds X[row_w + 4 ks + k][n]; the fragments of the
    // NEXT tile are loaded while the current tile's DMMAs run (one DRAM latency per tile would otherwise stall all 8 warps at once)
    double xf[MM_KS], xn[MM_KS];
    const bool rhs_ok = lm < nrhs;
    auto load_x = [&](int64_t tile, double (&dst)[MM_KS]) {
      const int64_t row_w = tile * MM_R + warp * MM_WROWS;
      const double *xp = p.x + (int64_t)lm * p.ldx + row_w + lk;
#pragma unroll
      for (int ks = 0; ks < MM_KS; ++ks) dst[ks] = (rhs_ok && row_w + 4 * ks + lk < p.n) ? __ldg(xp + 4 * ks) : 0.0;
    };
    if (my_tiles > 0) load_x(blockIdx.x, xn);
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int64_t t = blockIdx.x + i * grid;
#pragma unroll
      for (int ks = 0; ks < MM_KS; ++ks) xf[ks] = xn[ks];
      if (i + 1 < my_tiles) load_x(t + grid, xn);
#pragma unroll
      for (int g = 0; g < MM_MAXG; ++g) {
        if (8 * g >= ncols) break;
        const int nc = min(8, ncols - 8 * g);
        // the group's column tiles sit in nc consecutive ring slots
        const uint32_t s0 = slot, p0 = par;
        for (int j = 0; j < nc; ++j) {
          mm_wait(&full[slot], par, 200 + (int)slot);
          if (++slot == (uint32_t)S) {
            slot = 0;
            par ^= 1u;
          }
        }
        (void)p0;
        __syncwarp();   // mma.sync.aligned needs the warp CONVERGED: lanes leave the polling loops at different iterations
        const bool col_ok = lm < nc;
        uint32_t myslot = s0 + (uint32_t)(col_ok ? lm : 0);
        if (myslot >= (uint32_t)S) myslot -= (uint32_t)S;
        const double *ap = reinterpret_cast<const double *>(mm_slot_ptr(ring, myslot)) + warp * MM_WROWS + lk;   // A[m = column lm][k]
#pragma unroll
        for (int ks = 0; ks < MM_KS; ++ks) {
          const double av = ap[4 * ks];             // (lanes of missing columns read slot s0 and are zeroed by the select)
          dmma884(C[g][ks & 1], col_ok ? av : 0.0, xf[ks]);
        }
        // hand the slots back once the DMMAs that consumed the loads have executed (data-dependent store, see smem_reads_landed)
        smem_reads_landed(&s_landed[tid], (unsigned)__double2hiint(C[g][0][0]) ^ (unsigned)__double2hiint(C[g][1][0]));
        __syncwarp();
        if (lane < nc) {
          uint32_t rs = s0 + (uint32_t)lane;
          if (rs >= (uint32_t)S) rs -= (uint32_t)S;
          mbar_arrive(&empty[rs]);
        }
      }
    }
    // per-CTA G: lane holds G[c = 8g + lm][r = 2 lk + {0,1}] of its warp's rows; warps are added in index order
#pragma unroll
    for (int g = 0; g < MM_MAXG; ++g) {
      const int c = 8 * g + lm;
      if (c < ncols) {
        wacc[((size_t)warp * ncols + c) * NR + 2 * lk] = C[g][0][0] + C[g][1][0];
        wacc[((size_t)warp * ncols + c) * NR + 2 * lk + 1] = C[g][0][1] + C[g][1][1];
      }
    }
    consumers_sync();
    for (int i = tid; i < nv; i += B2O_NCONS) {
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < B2O_CONS_WARPS; ++w) sum += wacc[(size_t)w * nv + i];
      p.partials[(size_t)blockIdx.x * nv + i] = sum;
    }
  }

  // ------------------------------------------------------------------ reduce G over the grid (and over the ranks)
  unsigned long long bar_target = p.bar_target;
  if (p.mbox.nranks > 1) {
    const int nep = (nv + MBOX_MAXV - 1) / MBOX_MAXV;
    const unsigned long long epoch_last = p.mbox.epoch_base + nep;
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      atomicAdd(p.bar, 1ULL);
    }
    if (blockIdx.x == 0) {
      if (tid == 0) {
        while (ld_acquire_u64(p.bar) < bar_target) { __nanosleep(32); }
        __threadfence();
      }
      __syncthreads();
      if (!is_producer) {
        for (int c = warp; c < nv; c += B2O_CONS_WARPS) {
          double s = 0.0;
          for (int b = lane; b < (int)grid; b += 32) s += __ldcg(&p.partials[(size_t)b * nv + c]);
          s = warp_sum(s);
          if (lane == 0) coef[c] = s;
        }
      }
      __syncthreads();
      if (warp == 0)
        for (int e = 0; e < nep; ++e)
          mbox_allreduce_warp(p.mbox, p.mbox.epoch_base + 1 + e, coef + e * MBOX_MAXV, min(MBOX_MAXV, nv - e * MBOX_MAXV));
      __syncthreads();
      for (int c = tid; c < nv; c += B2O_NTHREADS) p.dots[c] = coef[c];
      __threadfence();
      __syncthreads();
      if (tid == 0) st_release_gpu_u64(p.mbox.ready, epoch_last);
    } else {
      if (tid == 0)
        while (ld_acquire_u64(p.mbox.ready) < epoch_last) { __nanosleep(32); }
      __syncthreads();
      for (int c = tid; c < nv; c += B2O_NTHREADS) coef[c] = __ldcg(&p.dots[c]);
      __syncthreads();
    }
  } else {
    grid_barrier(p.bar, bar_target);
    if (!is_producer) {
      for (int c = warp; c < nv; c += B2O_CONS_WARPS) {
        double s = 0.0;
        for (int b = lane; b < (int)grid; b += 32) s += __ldcg(&p.partials[(size_t)b * nv + c]);
        s = warp_sum(s);
        if (lane == 0) coef[c] = s;
      }
    }
    __syncthreads();
  }

  if (OP == OP_INV_COMPACT) {
    // coefficients(:, r) = W * G(:, r); every CTA repeats the tiny product in the same order
    for (int idx = tid; idx < nv; idx += B2O_NTHREADS) {
      const int j = idx / NR, r = idx % NR;
      double s = 0.0;
      for (int k = 0; k < ncols; ++k) s = fma(__ldcg(&p.W[(size_t)j * ncols + k]), coef[k * NR + r], s);
      wacc[idx] = s;
    }
    __syncthreads();
    for (int idx = tid; idx < nv; idx += B2O_NTHREADS) coef[idx] = wacc[idx];
    __syncthreads();
  }
  // signed / scaled coefficient of column c for right-hand side r, as the combine uses it
  for (int idx = tid; idx < nv; idx += B2O_NTHREADS) {
    const int c = idx / NR;
    double v = coef[idx];
    if (OP == OP_LBFGS_FWD) v = (c & 1) ? v : -v;                      // q += bx b_k - ax a_k   (columns a_k, b_k alternate)  src/lbfgs.jl:194
    else if (OP == OP_LSR1) v = (p.alpha * v) / p.cdiv[c];              // ax = α dot(a_k, x) / as_k                            src/lsr1.jl:101
    wacc[idx] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < nv; idx += B2O_NTHREADS) coef[idx] = wacc[idx];
  __syncthreads();

  // ------------------------------------------------------------------ phase 2: Res = base(X) + cols * coef
  if (is_producer) {
    if (lane == 0) {
      for (int64_t i = my_tiles - 1; i >= 0; --i) {   // reverse order: the tail of phase 1 is still in L2
        const int64_t t = blockIdx.x + i * grid;
        for (int c = 0; c < ncols; ++c) push(p.cols[c] + t * MM_R);
      }
    }
    __syncwarp();
  } else {
    const double alpha = p.alpha, beta = p.beta, gamma = p.gamma, inv_gamma = 1.0 / p.gamma;
    const int n0 = 2 * lk;                               // the lane's two right-hand sides
    const bool ok0 = n0 < nrhs, ok1 = n0 + 1 < nrhs;
    double xb[MM_RG][2];                                   // X values of the NEXT tile to process (prefetched during the current one)
    auto load_xb = [&](int64_t tile) {
      const int64_t row_w = tile * MM_R + warp * MM_WROWS + lm;
#pragma unroll
      for (int rg = 0; rg < MM_RG; ++rg) {
        const int64_t row = row_w + 8 * rg;
        xb[rg][0] = (row < p.n && ok0) ? __ldg(p.x + (int64_t)n0 * p.ldx + row) : 0.0;
        xb[rg][1] = (row < p.n && ok1) ? __ldg(p.x + (int64_t)(n0 + 1) * p.ldx + row) : 0.0;
      }
    };
    if (my_tiles > 0) load_xb(blockIdx.x + (my_tiles - 1) * grid);
    for (int64_t i = my_tiles - 1; i >= 0; --i) {
      const int64_t t = blockIdx.x + i * grid;
      const int64_t row_w = t * MM_R + warp * MM_WROWS + lm;  // the lane's row inside row group 0
      double C[MM_RG][2];
#pragma unroll
      for (int rg = 0; rg < MM_RG; ++rg) {
        const int64_t row = row_w + 8 * rg;
        const bool rok = row < p.n;
        const double x0 = xb[rg][0], x1 = xb[rg][1];
        if (OP == OP_LBFGS_FWD) {
          C[rg][0] = p.scaling ? x0 * inv_gamma : x0;                                              // src/lbfgs.jl:183-186 (one reciprocal)
          C[rg][1] = p.scaling ? x1 * inv_gamma : x1;
        } else if (OP == OP_INV_COMPACT) {
          C[rg][0] = !p.scaling ? x0 : (p.base_div ? x0 * inv_gamma : x0 * gamma);
          C[rg][1] = !p.scaling ? x1 : (p.base_div ? x1 * inv_gamma : x1 * gamma);
        } else {
          C[rg][0] = (alpha * x0) * inv_gamma;                                                      // src/lsr1.jl:92-96
          C[rg][1] = (alpha * x1) * inv_gamma;
          if (beta != 0.0) {
            if (rok && ok0) C[rg][0] += beta * p.res[(int64_t)n0 * p.ldr + row];
            if (rok && ok1) C[rg][1] += beta * p.res[(int64_t)(n0 + 1) * p.ldr + row];
          }
        }
      }
      if (i > 0) load_xb(t - grid);
      for (int c0 = 0; c0 < ncols; c0 += 4) {
        const int nc = min(4, ncols - c0);
        const uint32_t s0 = slot;
        for (int j = 0; j < nc; ++j) {
          mm_wait(&full[slot], par, 300 + (int)slot);
          if (++slot == (uint32_t)S) {
            slot = 0;
            par ^= 1u;
          }
        }
        __syncwarp();   // converge before the aligned DMMAs
        const bool col_ok = lk < nc;
        uint32_t myslot = s0 + (uint32_t)(col_ok ? lk : 0);
        if (myslot >= (uint32_t)S) myslot -= (uint32_t)S;
        const double *ap = reinterpret_cast<const double *>(mm_slot_ptr(ring, myslot)) + warp * MM_WROWS + lm;   // A[m = row lm][k = column lk]
        const double bf = col_ok ? coef[(c0 + lk) * NR + lm] : 0.0;                                          // B[k = column lk][n = rhs lm]
        unsigned fold = 0;
#pragma unroll
        for (int rg = 0; rg < MM_RG; ++rg) {
          const double av = ap[8 * rg];
          dmma884(C[rg], col_ok ? av : 0.0, bf);
          fold ^= (unsigned)__double2hiint(C[rg][0]);
        }
        smem_reads_landed(&s_landed[tid], fold);
        __syncwarp();
        if (lane < nc) {
          uint32_t rs = s0 + (uint32_t)lane;
          if (rs >= (uint32_t)S) rs -= (uint32_t)S;
          mbar_arrive(&empty[rs]);
        }
      }
#pragma unroll
      for (int rg = 0; rg < MM_RG; ++rg) {
        const int64_t row = row_w + 8 * rg;
        if (row < p.n) {
          double *r0 = p.res + (int64_t)n0 * p.ldr + row, *r1 = p.res + (int64_t)(n0 + 1) * p.ldr + row;
          if (OP == OP_LSR1) {
            if (ok0) *r0 = C[rg][0];
            if (ok1) *r1 = C[rg][1];
          } else if (beta != 0.0) {
            if (ok0) *r0 = alpha * C[rg][0] + beta * *r0;                                           // src/lbfgs.jl:197-201
            if (ok1) *r1 = alpha * C[rg][1] + beta * *r1;
          } else {
            if (ok0) *r0 = alpha * C[rg][0];
            if (ok1) *r1 = alpha * C[rg][1];
          }
        }
      }
    }
  }
}
