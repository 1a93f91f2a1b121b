// b2o_qn_multi.cuh -- block (multi right-hand-side) variant of qn_compact_kernel: mul!(Res::Matrix, op, X::Matrix, α, β)
// for the forward LBFGSOperator, LSR1Operator and the compact forms (the reference hands the matrices to
// the closure, src/operations.jl:34-36 -- its quasi-Newton closures are vector code, so this is an extension with the
// per-column semantics of src/lbfgs.jl:173-202 / src/lsr1.jl:89-107).
//
// Same two streaming phases as the vector kernel, but every TMA-staged column tile is used for NR right-hand sides while it
// is in shared memory:  algorithmic DRAM bytes per launch (2*ncols + 3*nrhs)*8*n  instead of nrhs*(2*ncols+3)*8*n.
//   phase 1: G[c][r] = col_c · x_r.  The thread's x rows stay in registers (NR x EPT doubles, one tile prefetched ahead);
//            per column tile the NR thread partials are folded to one value per lane by a transposing butterfly
//            (NR/2 + NR/4 + ... exchange shuffles) and accumulated in the lane's shared-memory cell [warp][c][lane]
//            -> fixed order, deterministic.  With 8 right-hand sides the kernel sits near the FP64 pipe's balance point
//            (64 DFMA/clk/SM against 25 B/clk/SM of HBM), so the arithmetic is contracted to FMAs here (the vector kernel
//            keeps the reference's unfused statement rounding); per-column results differ from it by a few ulp.
//   phase 2: res_r = α (x_r/γ + Σ_c coef[c][r] col_c) + β res_r with NR x EPT accumulators in registers.
#pragma once
#include "b2o_qn_kernels.cuh"

constexpr int B2O_MULTI_MAXV = 256;   // ncols * NR values reduced per launch (fits d_dots and two mailbox epochs)

template <typename T>
struct MultiArgsT {
  const T *cols[B2O_MAX_COLS];
  double cdiv[B2O_MAX_COLS];
  int ncols;
  const T *x;        // column-major n x nrhs, leading dimension ldx
  T *res;            // column-major n x nrhs, leading dimension ldr
  int64_t ldx, ldr;
  int nrhs;          // <= NR
  int64_t n, ntiles;
  double alpha, beta, gamma;
  int scaling;
  int x_al16, res_al16;
  double *partials;                  // [grid][ncols*NR]
  double *dots;                      // [ncols*NR] (mailbox publication)
  unsigned long long *bar;
  unsigned long long bar_target;
  int stages;
  uint32_t wacc_off, coef_off, bar_off, landed_off;
  MboxDev mbox;
  const double *W;
  int base_div;
};
using MultiArgs = MultiArgsT<double>;

// Transposing butterfly: log2(NR) exchange stages fold the NR per-thread partials into ONE value per lane -- the sum over the
// NR lanes that differ in the top log2(NR) lane bits -- for right-hand side r = lane / (32/NR).  The remaining 32/NR lanes of a
// group are NOT folded here: each lane accumulates its value over all tiles in its own shared-memory cell and the group is
// summed once at the end (keeps the dependent shuffle chain per column tile short).  Fixed pattern -> deterministic.
template <int NR>
__device__ __forceinline__ double transpose_fold(double (&s)[NR], int lane);

template <>
__device__ __forceinline__ double transpose_fold<8>(double (&s)[8], int lane) {
  const unsigned F = 0xffffffffu;
  double t[4], u[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double send = h16 ? s[i] : s[i + 4];
    const double keep = h16 ? s[i + 4] : s[i];
    t[i] = keep + __shfl_xor_sync(F, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = h8 ? t[i] : t[i + 2];
    const double keep = h8 ? t[i + 2] : t[i];
    u[i] = keep + __shfl_xor_sync(F, send, 8);
  }
  const double send = h4 ? u[0] : u[1];
  const double keep = h4 ? u[1] : u[0];
  return keep + __shfl_xor_sync(F, send, 4);   // r = lane >> 2 (bit4 -> 4, bit3 -> 2, bit2 -> 1)
}

template <>
__device__ __forceinline__ double transpose_fold<4>(double (&s)[4], int lane) {
  const unsigned F = 0xffffffffu;
  double t[2];
  const bool h16 = lane & 16, h8 = lane & 8;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = h16 ? s[i] : s[i + 2];
    const double keep = h16 ? s[i + 2] : s[i];
    t[i] = keep + __shfl_xor_sync(F, send, 16);
  }
  const double send = h8 ? t[0] : t[1];
  const double keep = h8 ? t[1] : t[0];
  return keep + __shfl_xor_sync(F, send, 8);   // r = lane >> 3
}

template <>
__device__ __forceinline__ double transpose_fold<2>(double (&s)[2], int lane) {
  const bool h16 = lane & 16;
  const double send = h16 ? s[0] : s[1];
  const double keep = h16 ? s[1] : s[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 16);   // r = lane >> 4
}

// A ring slot is handed back to the TMA producer right after its tile was copied to registers, BEFORE the arithmetic (the
// shuffle chain is long).  ptxas then schedules the mbarrier arrive a few instructions behind LDS that are still in flight, and
// with that schedule the block apply showed intermittent 1e-8 errors at n = 1e8 (a tile refilled under the read).  The
// hand-back is therefore made data-dependent on every LDS of the tile: one 4-byte shared-memory store whose operand is
// folded from a register of each LDS.128 -- it cannot issue before the loads have returned, and the arrive is ordered after it.
template <int EPT>
__device__ __forceinline__ unsigned fold_loaded(const double (&a)[EPT]) {
  unsigned v = 0;
#pragma unroll
  for (int j = 0; j < EPT; j += 2) v ^= (unsigned)__double2hiint(a[j]);
  return v;
}
template <int EPT>
__device__ __forceinline__ unsigned fold_loaded(const float (&a)[EPT]) {   // one register of every LDS.128 (four floats)
  unsigned v = 0;
#pragma unroll
  for (int j = 0; j < EPT; j += 4) v ^= __float_as_uint(a[j]);
  return v;
}
#ifdef B2O_SIMT_EMU
inline void smem_reads_landed(unsigned *cell, unsigned v) { *cell = v; }
#else
__device__ __forceinline__ void smem_reads_landed(unsigned *cell, unsigned v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_u32(cell)), "r"(v) : "memory");
}
#endif

// rows per tile: 8 KB (8 right-hand sides) or 16 KB of column data per ring stage for either element type
template <int NR, typename T = double>
struct MultiTile {
  static constexpr int R = ((NR == 8) ? 1024 : 2048) * (int)(sizeof(double) / sizeof(T));
};

// T = float: the Float32 operators' matrix right-hand sides (columns, X and Res in Float32; the inner products and their
// reduction stay in double, the combine runs in Float32 FMAs)
template <int NR, int OP, typename T = double>
__global__ void __launch_bounds__(B2O_NTHREADS, 1) qn_multi_kernel(const __grid_constant__ MultiArgsT<T> p) {
  constexpr int R = MultiTile<NR, T>::R;
  constexpr int EPT = R / B2O_NCONS;
  constexpr int LPG = 32 / NR;   // lanes per right-hand-side group after the transposing reduce
#ifdef B2O_SIMT_EMU
  unsigned char *smem_raw = emu::dyn_smem();
#else
  extern __shared__ __align__(128) unsigned char smem_raw[];
#endif
  Ring rg;
  rg.buf = smem_raw;
  double *wacc = reinterpret_cast<double *>(smem_raw + p.wacc_off);   // [8 warps][ncols][32 lanes]
  double *coef = reinterpret_cast<double *>(smem_raw + p.coef_off);   // [ncols*NR]
  rg.full = reinterpret_cast<uint64_t *>(smem_raw + p.bar_off);
  rg.empty = rg.full + p.stages;
  rg.stages = p.stages;
  unsigned *s_landed = reinterpret_cast<unsigned *>(smem_raw + p.landed_off);   // [256] one cell per consumer thread

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_producer = warp == B2O_CONS_WARPS;
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&rg.full[s], 1);
      mbar_init(&rg.empty[s], B2O_CONS_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  RingPos pos(p.stages);
  const int64_t grid = gridDim.x;
  const int64_t my_tiles = (p.ntiles > (int64_t)blockIdx.x) ? (p.ntiles - 1 - blockIdx.x) / grid + 1 : 0;
  const int ncols = p.ncols, nrhs = p.nrhs;
  const int nv = ncols * NR;

  // ------------------------------------------------------------------ phase 1: G = colsᵀ X
  if (is_producer) {
    if (lane == 0) {
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int64_t t = blockIdx.x + i * grid;
        for (int c = 0; c < ncols; ++c) producer_push<R>(rg, pos, p.cols[c] + t * R);
      }
    }
    __syncwarp();
  } else {
    for (int i = tid; i < B2O_CONS_WARPS * ncols * 32; i += B2O_NCONS) wacc[i] = 0.0;
    consumers_sync();
    T xr[NR][EPT], xn[NR][EPT];
    if (my_tiles > 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        if (r < nrhs) load_user_tile<R>(p.x + (int64_t)r * p.ldx, (int64_t)blockIdx.x * R, p.n, p.x_al16, xn[r]);
        else {
#pragma unroll
          for (int j = 0; j < EPT; ++j) xn[r][j] = (T)0;
        }
      }
    }
    double *my_acc = wacc + (size_t)warp * ncols * 32 + lane;    // cell [warp][c][lane]
    auto take = [&](uint32_t slot, T (&a)[EPT]) { tile_from_ring<R>(rg, slot, a); };
    auto dots = [&](const T (&a)[EPT], double (&sv)[NR]) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        double acc = (double)a[0] * (double)xr[r][0];
#pragma unroll
        for (int j = 1; j < EPT; ++j) acc = fma((double)a[j], (double)xr[r][j], acc);
        sv[r] = acc;
      }
    };
    for (int64_t i = 0; i < my_tiles; ++i) {
      const int64_t t = blockIdx.x + i * grid;
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int j = 0; j < EPT; ++j) xr[r][j] = xn[r][j];
      if (i + 1 < my_tiles) {
#pragma unroll
        for (int r = 0; r < NR; ++r)
          if (r < nrhs) load_user_tile<R>(p.x + (int64_t)r * p.ldx, (t + grid) * R, p.n, p.x_al16, xn[r]);
      }
      int c = 0;
      for (; c + 1 < ncols; c += 2) {     // two columns per iteration: two independent shuffle chains in flight
        const uint32_t s0 = pos.slot, p0 = pos.par;
        pos.advance();
        const uint32_t s1 = pos.slot, p1 = pos.par;
        pos.advance();
        T a0[EPT], a1[EPT];
        double v0[NR], v1[NR];
        mbar_wait(&rg.full[s0], p0);
        take(s0, a0);
        mbar_wait(&rg.full[s1], p1);
        take(s1, a1);
        smem_reads_landed(&s_landed[tid], fold_loaded(a0) ^ fold_loaded(a1));
        consumer_release(rg, s0);         // the tiles are in registers: hand the slots back before the arithmetic
        consumer_release(rg, s1);
        dots(a0, v0);
        dots(a1, v1);
        const double f0 = transpose_fold<NR>(v0, lane), f1 = transpose_fold<NR>(v1, lane);
        my_acc[c * 32] += f0;
        my_acc[(c + 1) * 32] += f1;
      }
      if (c < ncols) {
        T a0[EPT];
        double v0[NR];
        mbar_wait(&rg.full[pos.slot], pos.par);
        take(pos.slot, a0);
        smem_reads_landed(&s_landed[tid], fold_loaded(a0));
        consumer_release(rg, pos.slot);
        pos.advance();
        dots(a0, v0);
        my_acc[c * 32] += transpose_fold<NR>(v0, lane);
      }
    }
    consumers_sync();
    // G[c][r] of this CTA: the LPG lanes of group r, then the 8 warps, in a fixed order
    for (int i = tid; i < nv; i += B2O_NCONS) {
      const int c = i / NR, r = i % NR;
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < B2O_CONS_WARPS; ++w) {
        double sw = 0.0;
#pragma unroll
        for (int l = 0; l < LPG; ++l) sw += wacc[((size_t)w * ncols + c) * 32 + r * LPG + l];
        sum += sw;
      }
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

  // ------------------------------------------------------------------ phase 2: combine and write Res
  if (is_producer) {
    if (lane == 0) {
      for (int64_t i = my_tiles - 1; i >= 0; --i) {
        const int64_t t = blockIdx.x + i * grid;
        for (int c = 0; c < ncols; ++c) producer_push<R>(rg, pos, p.cols[c] + t * R);
      }
    }
    __syncwarp();
  } else {
    const T alpha = (T)p.alpha, beta = (T)p.beta, gamma = (T)p.gamma;
    // base term with ONE reciprocal instead of NR x EPT divisions per tile: a double division is ~30 instructions, and with 8
    // right-hand sides the divisions of `q ./= γ` outweighed the combine's FMAs (ncu source page: 17 % of the warp samples,
    // profiles/r2_ncu_summary.md).  x * (1/γ) differs from x / γ by at most one ulp -- the block kernels are contracted anyway.
    const T inv_gamma = (T)1 / gamma;
    T xn[NR][EPT], q[NR][EPT];
    if (my_tiles > 0) {
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        if (r < nrhs) load_user_tile<R>(p.x + (int64_t)r * p.ldx, (blockIdx.x + (my_tiles - 1) * grid) * R, p.n, p.x_al16, xn[r]);
        else {
#pragma unroll
          for (int j = 0; j < EPT; ++j) xn[r][j] = (T)0;
        }
      }
    }
    for (int64_t i = my_tiles - 1; i >= 0; --i) {
      const int64_t t = blockIdx.x + i * grid;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        if (OP == OP_LBFGS_FWD) {
#pragma unroll
          for (int j = 0; j < EPT; ++j) q[r][j] = p.scaling ? xn[r][j] * inv_gamma : xn[r][j];      // src/lbfgs.jl:183-186
        } else if (OP == OP_INV_COMPACT) {
#pragma unroll
          for (int j = 0; j < EPT; ++j) q[r][j] = !p.scaling ? xn[r][j] : (p.base_div ? xn[r][j] * inv_gamma : xn[r][j] * gamma);
        } else {
          T rold[EPT];
          if (beta != (T)0 && r < nrhs) load_user_tile<R>(p.res + (int64_t)r * p.ldr, t * R, p.n, p.res_al16, rold);
#pragma unroll
          for (int j = 0; j < EPT; ++j) {                                                           // src/lsr1.jl:92-96
            const T v = (alpha * xn[r][j]) * inv_gamma;
            q[r][j] = (beta != (T)0 && r < nrhs) ? v + beta * rold[j] : v;
          }
        }
      }
      if (i > 0) {
#pragma unroll
        for (int r = 0; r < NR; ++r)
          if (r < nrhs) load_user_tile<R>(p.x + (int64_t)r * p.ldx, (t - grid) * R, p.n, p.x_al16, xn[r]);
      }
      if (OP == OP_LBFGS_FWD) {
        for (int c = 0; c < ncols; c += 2) {
          const uint32_t sa = pos.slot, pa = pos.par;
          pos.advance();
          const uint32_t sb = pos.slot, pb = pos.par;
          pos.advance();
          mbar_wait(&rg.full[sa], pa);
          mbar_wait(&rg.full[sb], pb);
          T a[EPT], b[EPT];
          tile_from_ring<R>(rg, sa, a);
          tile_from_ring<R>(rg, sb, b);
          smem_reads_landed(&s_landed[tid], fold_loaded(a) ^ fold_loaded(b));
          consumer_release(rg, sa);
          consumer_release(rg, sb);
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            const T ax = (T)coef[c * NR + r], bx = (T)coef[(c + 1) * NR + r];
#pragma unroll
            for (int j = 0; j < EPT; ++j) q[r][j] = fma(bx, b[j], fma(-ax, a[j], q[r][j]));          // src/lbfgs.jl:194, contracted
          }
        }
      } else {
        for (int c = 0; c < ncols; ++c) {
          mbar_wait(&rg.full[pos.slot], pos.par);
          T a[EPT];
          tile_from_ring<R>(rg, pos.slot, a);
          smem_reads_landed(&s_landed[tid], fold_loaded(a));
          consumer_release(rg, pos.slot);
          pos.advance();
          const T cd = (T)p.cdiv[c];
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            const T ax = (OP == OP_INV_COMPACT) ? (T)coef[c * NR + r] : (alpha * (T)coef[c * NR + r]) / cd;   // src/lsr1.jl:101
#pragma unroll
            for (int j = 0; j < EPT; ++j) q[r][j] = fma(ax, a[j], q[r][j]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        if (r < nrhs) {
          if (OP != OP_LSR1) {
            if (beta != (T)0) {
              T rold[EPT];
              load_user_tile<R>(p.res + (int64_t)r * p.ldr, t * R, p.n, p.res_al16, rold);
#pragma unroll
              for (int j = 0; j < EPT; ++j) q[r][j] = alpha * q[r][j] + beta * rold[j];             // src/lbfgs.jl:197-201
            } else {
#pragma unroll
              for (int j = 0; j < EPT; ++j) q[r][j] = alpha * q[r][j];
            }
          }
          store_user_tile<R>(p.res + (int64_t)r * p.ldr, t * R, p.n, p.res_al16, q[r]);
        }
      }
    }
  }
}

// =====================================================================================================
// Block two-loop recursion: mul!(Res::Matrix, H, X::Matrix, α, β) for the two-loop InverseLBFGSOperator
// (src/operations.jl:34-36 hands the matrices to the closure; per column this is src/lbfgs.jl:117-154).
//
// Same 2A+1 fused sweeps as qn_twoloop_kernel, for NR right-hand sides at once: the update column v1 (y_k or s_k) and the
// column of the next inner product v2 are staged ONCE per sweep and used for all NR work vectors q_r while they sit in
// registers.  Algorithmic DRAM bytes per sweep (2·nrhs + 2)·8·n instead of nrhs·4·8·n (q_r must still make its round trip:
// n doubles do not fit on chip).  Every elementwise statement and every reduction order is that of the vector kernel
// (per-thread s0/s1 pairs -> warp butterfly -> warps in order -> CTAs in lane-strided order), so column r of the result is
// BIT-IDENTICAL to the vector kernel applied to column r with the same tile size and grid (tests check ==).
// =====================================================================================================
struct TwoLoopMultiArgs {
  const double *s[B2O_MAX_MEM];  // active slots, newest -> oldest
  const double *y[B2O_MAX_MEM];
  double ys[B2O_MAX_MEM];
  int nact;
  const double *x;               // column-major n x nrhs
  double *res;
  int64_t ldx, ldr;
  int nrhs;                      // <= NR
  double *q;                     // library-owned work vectors [NR][qpitch], zero padded
  int64_t qpitch;
  int64_t n, ntiles;
  double alpha, beta, gamma;
  int scaling;
  int x_al16, res_al16;
  double *partials;              // [grid][NR]
  unsigned long long *bar;
  unsigned long long bar_target;
  int stages;
  uint32_t scal_off, bar_off;    // scal: alphas [B2O_MAX_MEM][NR] | sred [8][NR] | s_dot [NR]
  uint32_t landed_off;           // [256] one cell per consumer thread (smem_reads_landed)
};

template <int R, int NR>
__global__ void __launch_bounds__(B2O_NTHREADS, 1) qn_twoloop_multi_kernel(const __grid_constant__ TwoLoopMultiArgs p) {
  constexpr int EPT = R / B2O_NCONS;
#ifdef B2O_SIMT_EMU
  unsigned char *smem_raw = emu::dyn_smem();
#else
  extern __shared__ __align__(128) unsigned char smem_raw[];
#endif
  Ring rg;
  rg.buf = smem_raw;
  double *alphas = reinterpret_cast<double *>(smem_raw + p.scal_off);   // [B2O_MAX_MEM][NR]
  double *sred = alphas + B2O_MAX_MEM * NR;                             // [8 warps][NR]
  double *s_dot = sred + B2O_CONS_WARPS * NR;                           // [NR]
  rg.full = reinterpret_cast<uint64_t *>(smem_raw + p.bar_off);
  rg.empty = rg.full + p.stages;
  rg.stages = p.stages;
  unsigned *s_landed = reinterpret_cast<unsigned *>(smem_raw + p.landed_off);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_producer = warp == B2O_CONS_WARPS;
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&rg.full[s], 1);
      mbar_init(&rg.empty[s], B2O_CONS_WARPS);
    }
    mbar_fence_init();
  }
  if (tid < NR) s_dot[tid] = 0.0;
  __syncthreads();

  RingPos pos(p.stages);
  const int A = p.nact, nrhs = p.nrhs;
  const int64_t grid = gridDim.x;
  const int64_t my_tiles = (p.ntiles > (int64_t)blockIdx.x) ? (p.ntiles - 1 - blockIdx.x) / grid + 1 : 0;
  unsigned long long bar_target = p.bar_target;

  for (int w = 0; w <= 2 * A; ++w) {
    // sweep w, as in qn_twoloop_kernel; c1[r] is the step's coefficient for right-hand side r
    const bool dot_only = (w == 0);
    const bool loop1 = (w >= 1 && w <= A);
    const bool last = (w == 2 * A);
    const double *v1 = nullptr, *v2 = nullptr;
    double c1[NR];
    bool qin_is_x = false, apply_gamma = false;
#pragma unroll
    for (int r = 0; r < NR; ++r) c1[r] = 0.0;
    if (dot_only) {
      v2 = p.s[0];
      qin_is_x = true;
    } else if (loop1) {
      const int i = w;
#pragma unroll
      for (int r = 0; r < NR; ++r) c1[r] = s_dot[r] / p.ys[i - 1];                        // αk = dot(s[k], q) / ys[k]      :133
      if (tid < NR) alphas[(i - 1) * NR + tid] = s_dot[tid] / p.ys[i - 1];
      v1 = p.y[i - 1];
      qin_is_x = (i == 1);
      apply_gamma = (i == A) && p.scaling;
      v2 = (i < A) ? p.s[i] : p.y[A - 1];
    } else {
      const int i = w - A, o = A - i;
#pragma unroll
      for (int r = 0; r < NR; ++r) c1[r] = alphas[o * NR + r] - s_dot[r] / p.ys[o];       // β = αk - dot(y[k], q) / ys[k]  :144-145
      v1 = p.s[o];
      v2 = last ? nullptr : p.y[o - 1];
    }
    __syncthreads();

    if (is_producer) {
      if (lane == 0) {
        fence_proxy_async();
        for (int64_t i = 0; i < my_tiles; ++i) {
          const int64_t t = blockIdx.x + i * grid;
          if (v1) producer_push<R>(rg, pos, v1 + t * R);
          if (v2) producer_push<R>(rg, pos, v2 + t * R);
          if (!qin_is_x)
            for (int r = 0; r < nrhs; ++r) producer_push<R>(rg, pos, p.q + (int64_t)r * p.qpitch + t * R);
        }
      }
      __syncwarp();
    } else {
      double acc[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[r] = 0.0;
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int64_t t = blockIdx.x + i * grid;
        double a1[EPT], a2[EPT];
        uint32_t slot1 = 0xffffffffu, slot2 = 0xffffffffu;
        if (v1) {
          mbar_wait(&rg.full[pos.slot], pos.par);
          const double2 *V = reinterpret_cast<const double2 *>(rg.buf + (size_t)pos.slot * R * sizeof(double));
#pragma unroll
          for (int j = 0; j < EPT / 2; ++j) {
            double2 v = V[j * B2O_NCONS + tid];
            a1[2 * j] = v.x;
            a1[2 * j + 1] = v.y;
          }
          slot1 = pos.slot;
          pos.advance();
        }
        if (v2) {
          mbar_wait(&rg.full[pos.slot], pos.par);
          const double2 *V = reinterpret_cast<const double2 *>(rg.buf + (size_t)pos.slot * R * sizeof(double));
#pragma unroll
          for (int j = 0; j < EPT / 2; ++j) {
            double2 v = V[j * B2O_NCONS + tid];
            a2[2 * j] = v.x;
            a2[2 * j + 1] = v.y;
          }
          slot2 = pos.slot;
          pos.advance();
        }
        // v1 / v2 live in registers from here on: their slots go back BEFORE the q_r tiles are awaited (a consumer that held
        // them across the r loop would deadlock a ring with fewer than nrhs + 2 stages).  The hand-back is made data-dependent
        // on every LDS of the two tiles (see smem_reads_landed above).
        if (slot1 != 0xffffffffu || slot2 != 0xffffffffu) {
          smem_reads_landed(&s_landed[tid], (slot1 != 0xffffffffu ? fold_loaded(a1) : 0u) ^ (slot2 != 0xffffffffu ? fold_loaded(a2) : 0u));
          if (slot1 != 0xffffffffu) consumer_release(rg, slot1);
          if (slot2 != 0xffffffffu) consumer_release(rg, slot2);
        }
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          if (r >= nrhs) continue;
          double q[EPT], rold[EPT];
          uint32_t q_slot = 0xffffffffu;
          if (qin_is_x) {
            load_user_tile<R>(p.x + (int64_t)r * p.ldx, t * R, p.n, p.x_al16, q);          // q .= x   :127-128
          } else {
            mbar_wait(&rg.full[pos.slot], pos.par);
            const double2 *Q = reinterpret_cast<const double2 *>(rg.buf + (size_t)pos.slot * R * sizeof(double));
#pragma unroll
            for (int j = 0; j < EPT / 2; ++j) {
              double2 v = Q[j * B2O_NCONS + tid];
              q[2 * j] = v.x;
              q[2 * j + 1] = v.y;
            }
            q_slot = pos.slot;
            pos.advance();
            smem_reads_landed(&s_landed[tid], fold_loaded(q));
            consumer_release(rg, q_slot);
          }
          if (last && p.beta != 0.0) load_user_tile<R>(p.res + (int64_t)r * p.ldr, t * R, p.n, p.res_al16, rold);
          if (v1) {
#pragma unroll
            for (int j = 0; j < EPT; ++j) q[j] = loop1 ? q[j] - c1[r] * a1[j] : q[j] + c1[r] * a1[j];   // :135 / :146
            if (apply_gamma) {
#pragma unroll
              for (int j = 0; j < EPT; ++j) q[j] = q[j] * p.gamma;                          // q .*= γ            :139
            }
            if (last) {
#pragma unroll
              for (int j = 0; j < EPT; ++j) q[j] = (p.beta != 0.0) ? p.alpha * q[j] + p.beta * rold[j] : p.alpha * q[j];   // :149-153
              store_user_tile<R>(p.res + (int64_t)r * p.ldr, t * R, p.n, p.res_al16, q);
            } else {
              double2 *Qg = reinterpret_cast<double2 *>(p.q + (int64_t)r * p.qpitch + t * R);
#pragma unroll
              for (int j = 0; j < EPT / 2; ++j) Qg[j * B2O_NCONS + tid] = make_double2(q[2 * j], q[2 * j + 1]);
            }
          }
          if (v2) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int j = 0; j < EPT / 2; ++j) {
              s0 = fma(a2[2 * j], q[2 * j], s0);
              s1 = fma(a2[2 * j + 1], q[2 * j + 1], s1);
            }
            acc[r] += s0 + s1;
          }
        }
      }
      if (v2) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const double s = warp_sum(acc[r]);
          if (lane == 0) sred[warp * NR + r] = s;
        }
      }
      fence_proxy_async();  // our generic-proxy stores to q precede the next sweep's TMA reads of q
    }
    if (!v2) break;  // the last sweep wrote Res
    __syncthreads();
    if (tid < NR) {
      double s = 0.0;
      for (int wv = 0; wv < B2O_CONS_WARPS; ++wv) s += sred[wv * NR + tid];
      p.partials[(size_t)blockIdx.x * NR + tid] = s;
    }
    grid_barrier(p.bar, bar_target);
    bar_target += gridDim.x;
    if (warp < NR) {
      double s = 0.0;
      for (int b = lane; b < (int)grid; b += 32) s += __ldcg(&p.partials[(size_t)b * NR + warp]);
      s = warp_sum(s);
      if (lane == 0) s_dot[warp] = s;
    }
    __syncthreads();
  }
}
