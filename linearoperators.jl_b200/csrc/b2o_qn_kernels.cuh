// b2o_qn_kernels.cuh -- persistent kernels for the quasi-Newton applies.
//
//  qn_compact_kernel<R,OP> : forward LBFGSOperator (src/lbfgs.jl:173-202) and LSR1Operator
//                            (src/lsr1.jl:89-107) applies.  Phase 1 streams every active column once and
//                            takes all dots against x; grid barrier; phase 2 re-streams the columns and
//                            writes res.  Algorithmic DRAM bytes (2*ncols+3)*8*n (+8n if beta != 0).
//  qn_twoloop_kernel<R>    : InverseLBFGSOperator two-loop recursion (src/lbfgs.jl:117-154) as 2A+1
//                            sweeps, the axpy of step i fused with the dot of step i+1.
//                            Algorithmic DRAM bytes (8A+2)*8*n.
#pragma once
#include "b2o_stream.cuh"

enum { OP_LBFGS_FWD = 0, OP_LSR1 = 1, OP_INV_COMPACT = 2, OP_PUSH_A = 3, OP_PUSH_L = 4 };
// OP_PUSH_A: one step of the a_k rebuild inside push! (src/lbfgs.jl:236-250): a_k = s_k/γ + Σ_{l<k} [(b_l·s_k) b_l − (a_l·s_k) a_l]
// is the forward apply of the operator truncated to the pairs older than k, taken at x = s_k -- same two streaming phases; the
// combine keeps the reference's TWO statements per pair (.+= then .-=) and the dot s_k·a_k (:248) is taken on the way out
// (per-CTA partials at partials[grid*ncols + 2*cta]).
// OP_PUSH_L: the same for L-SR1 (src/lsr1.jl:169-179): a_k = (y_k − s_k/γ) − Σ_{l<k} ((a_l·s_k)/as_l) a_l with x = s_k, y2 = y_k;
// takes as_k = a_k·s_k and ‖a_k‖² on the way out (partials[grid*ncols + 2*cta + {0,1}]).
enum { MODE_FUSED = 0, MODE_PHASE1 = 1, MODE_PHASE2 = 2 };

template <typename T>
struct CompactArgsT {
  const T *cols[B2O_MAX_COLS];       // active columns in reference order (LBFGS: a_k,b_k pairs oldest->newest)
  double cdiv[B2O_MAX_COLS];         // LSR1: as[k]
  int ncols;
  const T *x;
  T *res;
  int64_t n, ntiles;
  double alpha, beta, gamma;
  int scaling;
  int x_al16, res_al16;
  double *partials;                  // [grid][ncols]
  double *dots;                      // [ncols] (split mode)
  unsigned long long *bar;           // grid barrier counter (monotonic)
  unsigned long long bar_target;
  unsigned long long *arrive;        // last-block counter (split mode)
  int mode, stages, group;
  int accumulate;                    // split mode: dots[c] += this launch's partial (row-chunked host pipeline)
  uint32_t accs_off, coef_off, bar_off;
  MboxDev mbox;                      // nranks > 1: the dots are all-reduced in-kernel through the NVLink peer mailbox
  const T *y2;                       // OP_PUSH_L: y_k (library-owned column)
  const double *W;                   // OP_INV_COMPACT: ncols x ncols middle matrix (row-major), coefficients = W * dots
  int base_div;                      // OP_INV_COMPACT: base term x/γ (compact FORWARD form) instead of γx (compact inverse)
  double *dbg;                       // [0] += ns CTA 0 waited for the local CTAs, [1] += ns in the mailbox exchange, [2] += epochs
};
using CompactArgs = CompactArgsT<double>;

// T = double: the reference's Float64 operators.  T = float: LBFGSOperator(Float32, n) etc. (test/test_lbfgs.jl:162-178) --
// columns, x and res are Float32, every elementwise statement runs in Float32, every inner product is accumulated in double
// and rounded to Float32 where the reference's `dot` returns one (the coefficient casts below are no-ops for T = double).
template <int R, int OP, typename T = double>
__global__ void __launch_bounds__(B2O_NTHREADS, 1) qn_compact_kernel(const __grid_constant__ CompactArgsT<T> p) {
  constexpr int EPT = R / B2O_NCONS;
#ifdef B2O_SIMT_EMU
  unsigned char *smem_raw = emu::dyn_smem();
#else
  extern __shared__ __align__(128) unsigned char smem_raw[];
#endif
  Ring rg;
  rg.buf = smem_raw;
  double *accs = reinterpret_cast<double *>(smem_raw + p.accs_off);
  double *coef = reinterpret_cast<double *>(smem_raw + p.coef_off);
  rg.full = reinterpret_cast<uint64_t *>(smem_raw + p.bar_off);
  rg.empty = rg.full + p.stages;
  rg.stages = p.stages;
  __shared__ bool s_is_last;

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
  const int ncols = p.ncols;

  // ------------------------------------------------------------------ phase 1: all dots against x
  if (p.mode != MODE_PHASE2 && ncols > 0) {
    for (int g0 = 0; g0 < ncols; g0 += p.group) {
      const int gc = min(p.group, ncols - g0);
      if (is_producer) {
        if (lane == 0) {
          for (int64_t i = 0; i < my_tiles; ++i) {
            const int64_t t = blockIdx.x + i * grid;
            for (int c = 0; c < gc; ++c) producer_push<R>(rg, pos, p.cols[g0 + c] + t * R);
          }
        }
        __syncwarp();
      } else {
        for (int c = 0; c < gc; ++c) accs[c * B2O_NCONS + tid] = 0.0;
        T xr[EPT], xn[EPT];
        if (my_tiles > 0) load_user_tile<R>(p.x, (int64_t)blockIdx.x * R, p.n, p.x_al16, xn);
        for (int64_t i = 0; i < my_tiles; ++i) {
          const int64_t t = blockIdx.x + i * grid;
#pragma unroll
          for (int j = 0; j < EPT; ++j) xr[j] = xn[j];
          if (i + 1 < my_tiles) load_user_tile<R>(p.x, (t + grid) * R, p.n, p.x_al16, xn);
          for (int c = 0; c < gc; ++c) {
            mbar_wait(&rg.full[pos.slot], pos.par);
            T b[EPT];
            tile_from_ring<R>(rg, pos.slot, b);
            accs[c * B2O_NCONS + tid] += tile_dot<EPT>(b, xr);
            consumer_release(rg, pos.slot);
            pos.advance();
          }
        }
        consumers_sync();
        for (int c = warp; c < gc; c += B2O_CONS_WARPS) {
          double s = 0.0;
#pragma unroll
          for (int w = 0; w < B2O_CONS_WARPS; ++w) s += accs[c * B2O_NCONS + w * 32 + lane];
          s = warp_sum(s);
          if (lane == 0) p.partials[(size_t)blockIdx.x * ncols + g0 + c] = s;
        }
        consumers_sync();
      }
    }
  }

  // ------------------------------------------------------------------ reduce the dots
  unsigned long long bar_target = p.bar_target;
  if (ncols > 0) {
    if (p.mode == MODE_FUSED && p.mbox.nranks > 1) {
      // row-partitioned, one launch per GPU: CTA 0 gathers the local partials, all-reduces them with the peers through the
      // NVLink mailbox and publishes the global dots; the other CTAs wait for the publication instead of a second barrier.
      const unsigned long long epoch = p.mbox.epoch_base + 1;
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicAdd(p.bar, 1ULL);
      }
      if (blockIdx.x == 0) {
        unsigned long long t0 = 0, t1 = 0;
        if (tid == 0) {
          t0 = globaltimer_ns();
          while (ld_acquire_u64(p.bar) < bar_target) { __nanosleep(32); }
          __threadfence();
          t1 = globaltimer_ns();
        }
        __syncthreads();
        if (!is_producer) {
          for (int c = warp; c < ncols; c += B2O_CONS_WARPS) {
            double s = 0.0;
            for (int b = lane; b < (int)grid; b += 32) s += __ldcg(&p.partials[(size_t)b * ncols + c]);
            s = warp_sum(s);
            if (lane == 0) coef[c] = s;
          }
        }
        __syncthreads();
        if (warp == 0) mbox_allreduce_warp(p.mbox, epoch, coef, ncols);
        if (tid == 0 && p.dbg) {
          const unsigned long long t2 = globaltimer_ns();
          p.dbg[0] += (double)(t1 - t0);
          p.dbg[1] += (double)(t2 - t1);
          p.dbg[2] += 1.0;
        }
        __syncthreads();
        for (int c = tid; c < ncols; c += B2O_NTHREADS) p.dots[c] = coef[c];
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release_gpu_u64(p.mbox.ready, epoch);
      } else {
        if (tid == 0)
          while (ld_acquire_u64(p.mbox.ready) < epoch) { __nanosleep(32); }
        __syncthreads();
        for (int c = tid; c < ncols; c += B2O_NTHREADS) coef[c] = __ldcg(&p.dots[c]);
        __syncthreads();
      }
      bar_target += gridDim.x;
    } else if (p.mode == MODE_FUSED) {
      grid_barrier(p.bar, bar_target);
      bar_target += gridDim.x;
      if (!is_producer) {
        for (int c = warp; c < ncols; c += B2O_CONS_WARPS) {
          double s = 0.0;
          for (int b = lane; b < (int)grid; b += 32) s += __ldcg(&p.partials[(size_t)b * ncols + c]);
          s = warp_sum(s);
          if (lane == 0) {
            coef[c] = s;
            if (blockIdx.x == 0) p.dots[c] = s;   // left for the parity tests (b2o_ctx_debug_read)
          }
        }
      }
      __syncthreads();
    } else if (p.mode == MODE_PHASE1) {
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        unsigned long long tk = atomicAdd(p.arrive, 1ULL);
        s_is_last = (tk == gridDim.x - 1);
      }
      __syncthreads();
      if (s_is_last) {
        __threadfence();
        if (!is_producer) {
          for (int c = warp; c < ncols; c += B2O_CONS_WARPS) {
            double s = 0.0;
            for (int b = lane; b < (int)grid; b += 32) s += __ldcg(&p.partials[(size_t)b * ncols + c]);
            s = warp_sum(s);
            if (lane == 0) p.dots[c] = p.accumulate ? __ldcg(&p.dots[c]) + s : s;
          }
        }
        if (tid == 0) *p.arrive = 0ULL;
      }
      return;
    } else {
      for (int c = tid; c < ncols; c += B2O_NTHREADS) coef[c] = __ldcg(&p.dots[c]);
      __syncthreads();
    }
  } else if (p.mode == MODE_PHASE1) {
    return;
  }

  if (OP == OP_INV_COMPACT && ncols > 0) {
    // compact representation (Byrd-Nocedal-Schnabel): coefficients = W * [Sᵀx; Yᵀx]; every CTA does the tiny matvec
    // redundantly in the same order -> identical coefficients everywhere
    for (int j = tid; j < ncols; j += B2O_NTHREADS) {
      double s = 0.0;
      for (int k = 0; k < ncols; ++k) s = fma(__ldcg(&p.W[(size_t)j * ncols + k]), coef[k], s);
      accs[j] = s;
    }
    __syncthreads();
    for (int j = tid; j < ncols; j += B2O_NTHREADS) coef[j] = accs[j];
    __syncthreads();
  }

  // ------------------------------------------------------------------ phase 2: combine and write res
  if (is_producer) {
    if (lane == 0) {
      for (int64_t i = my_tiles - 1; i >= 0; --i) {   // reverse order: the tail of phase 1 is still in L2
        const int64_t t = blockIdx.x + i * grid;
        for (int c = 0; c < ncols; ++c) producer_push<R>(rg, pos, p.cols[c] + t * R);
      }
    }
    __syncwarp();
  } else {
    const T alpha = (T)p.alpha, beta = (T)p.beta, gamma = (T)p.gamma;
    T xn[EPT], q[EPT], rold[EPT];
    T xc[(OP == OP_PUSH_A || OP == OP_PUSH_L) ? EPT : 1];
    double racc = 0.0, racc2 = 0.0;
    if (my_tiles > 0) load_user_tile<R>(p.x, (blockIdx.x + (my_tiles - 1) * grid) * R, p.n, p.x_al16, xn);
    for (int64_t i = my_tiles - 1; i >= 0; --i) {
      const int64_t t = blockIdx.x + i * grid;
      if (beta != (T)0) load_user_tile<R>(p.res, t * R, p.n, p.res_al16, rold);
      if (OP == OP_LBFGS_FWD) {
        // q .= x ; scaling && (q ./= γ)                                   src/lbfgs.jl:183-186
#pragma unroll
        for (int j = 0; j < EPT; ++j) q[j] = p.scaling ? xn[j] / gamma : xn[j];
      } else if (OP == OP_PUSH_A) {
        // a[k] .= s[k] ./ γ                                               src/lbfgs.jl:239
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
          xc[j] = xn[j];
          q[j] = xn[j] / gamma;
        }
      } else if (OP == OP_PUSH_L) {
        // a[k] .= y[k] .- s[k] ./ γ                                       src/lsr1.jl:169
        load_user_tile<R>(p.y2, t * R, p.n, true, rold);
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
          xc[j] = xn[j];
          q[j] = rold[j] - xn[j] / gamma;
        }
      } else if (OP == OP_INV_COMPACT) {
        // H0 x = γ x (γ = 1 without scaling)
#pragma unroll
        for (int j = 0; j < EPT; ++j) q[j] = !p.scaling ? xn[j] : (p.base_div ? xn[j] / gamma : xn[j] * gamma);
      } else {
        // q .= α .* x ./ γ (.+ β .* q)                                    src/lsr1.jl:92-96
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
          T v = (alpha * xn[j]) / gamma;
          q[j] = (beta != (T)0) ? v + beta * rold[j] : v;
        }
      }
      if (i > 0) load_user_tile<R>(p.x, (t - grid) * R, p.n, p.x_al16, xn);
      if (OP == OP_LBFGS_FWD || OP == OP_PUSH_A) {
        for (int c = 0; c < ncols; c += 2) {
          // q .+= bx .* b[k] .- ax .* a[k]                                src/lbfgs.jl:194
          const uint32_t sa = pos.slot, pa = pos.par;
          pos.advance();
          const uint32_t sb = pos.slot, pb = pos.par;
          pos.advance();
          const T ax = (T)coef[c], bx = (T)coef[c + 1];
          mbar_wait(&rg.full[sa], pa);
          mbar_wait(&rg.full[sb], pb);
          T a[EPT], b[EPT];
          tile_from_ring<R>(rg, sa, a);
          tile_from_ring<R>(rg, sb, b);
#pragma unroll
          for (int j = 0; j < EPT; ++j) {
            if (OP == OP_PUSH_A) {
              // a[k] .+= dot(b[l], s[k]) .* b[l] ; a[k] .-= dot(a[l], s[k]) .* a[l]      src/lbfgs.jl:244-245
              q[j] = (q[j] + bx * b[j]) - ax * a[j];
            } else {
              q[j] = q[j] + (bx * b[j] - ax * a[j]);
            }
          }
          consumer_release(rg, sa);
          consumer_release(rg, sb);
        }
        if (OP == OP_PUSH_A) {
#pragma unroll
          for (int j = 0; j < EPT; ++j) racc = fma((double)q[j], (double)xc[j], racc);           // dot(s[k], a[k])   :248
        } else {
#pragma unroll
          for (int j = 0; j < EPT; ++j) q[j] = (beta != (T)0) ? alpha * q[j] + beta * rold[j] : alpha * q[j];  // :197-201
        }
      } else {
        for (int c = 0; c < ncols; ++c) {
          // LSR1: ax = α * dot(a[k], x) / as[k];  q[j] += ax * a[k][j]    src/lsr1.jl:101-104
          // compact inverse: q += c_j * col_j
          // push! of L-SR1: as = dot(a[l], s[k]) / as[l];  a[k] .-= as .* a[l]              src/lsr1.jl:173-174
          const T ax = (OP == OP_INV_COMPACT) ? (T)coef[c]
                       : (OP == OP_PUSH_L)   ? (T)coef[c] / (T)p.cdiv[c]
                                             : (alpha * (T)coef[c]) / (T)p.cdiv[c];
          mbar_wait(&rg.full[pos.slot], pos.par);
          T a[EPT];
          tile_from_ring<R>(rg, pos.slot, a);
#pragma unroll
          for (int j = 0; j < EPT; ++j) {
            if (OP == OP_PUSH_L) q[j] = q[j] - ax * a[j];
            else q[j] = q[j] + ax * a[j];
          }
          consumer_release(rg, pos.slot);
          pos.advance();
        }
        if (OP == OP_INV_COMPACT) {
#pragma unroll
          for (int j = 0; j < EPT; ++j) q[j] = (beta != (T)0) ? alpha * q[j] + beta * rold[j] : alpha * q[j];
        }
        if (OP == OP_PUSH_L) {
#pragma unroll
          for (int j = 0; j < EPT; ++j) {
            racc = fma((double)q[j], (double)xc[j], racc);                                        // as[k] = dot(a[k], s[k])   :177
            racc2 = fma((double)q[j], (double)q[j], racc2);                                       // norm(a[k])^2              :179
          }
        }
      }
      store_user_tile<R>(p.res, t * R, p.n, p.res_al16, q);
    }
    if (OP == OP_PUSH_A || OP == OP_PUSH_L) {
      // per-CTA partials of the outgoing dots in a fixed order (accs is free after phase 1)
      racc = warp_sum(racc);
      racc2 = warp_sum(racc2);
      if (lane == 0) {
        accs[warp] = racc;
        accs[B2O_CONS_WARPS + warp] = racc2;
      }
      consumers_sync();
      if (tid < 2) {
        double sum = 0.0;
        for (int w = 0; w < B2O_CONS_WARPS; ++w) sum += accs[tid * B2O_CONS_WARPS + w];
        p.partials[(size_t)gridDim.x * ncols + 2 * blockIdx.x + tid] = sum;
      }
    }
  }
}

// =====================================================================================================
// Inverse L-BFGS two-loop recursion, src/lbfgs.jl:117-154
// =====================================================================================================
constexpr int B2O_MAX_MEM = 64;

template <typename T>
struct TwoLoopArgsT {
  const T *s[B2O_MAX_MEM];       // active slots ordered newest -> oldest (loop-1 order)
  const T *y[B2O_MAX_MEM];
  double ys[B2O_MAX_MEM];
  int nact;
  const T *x;
  T *res;
  T *q;                          // library-owned work vector (data.Ax), padded pitch
  double *alpha_out;             // device copy of data.α in loop-1 order (may be null)
  int64_t n, ntiles;
  double alpha, beta, gamma;
  int scaling;
  int x_al16, res_al16;
  double *partials;              // [grid]
  double *dots;                  // step mode: dots[0] carries the reduced dot between launches, dots[1+i] the α_i
  unsigned long long *bar;
  unsigned long long bar_target;
  unsigned long long *arrive;
  int stages;
  int sweep_begin, sweep_end;    // sweeps [begin,end) of 0..2A run in this launch (fused: 0..2A+1)
  uint32_t coef_off, bar_off;
  MboxDev mbox;                  // nranks > 1: each inner product is all-reduced in-kernel through the NVLink peer mailbox
  double *sweep_dots;            // fused mode: the reduced inner product of sweep w is also left in sweep_dots[w] (parity tests)
  double *dbg;                   // [0] += ns CTA 0 waited for the local CTAs, [1] += ns in the mailbox exchange, [2] += epochs
};
using TwoLoopArgs = TwoLoopArgsT<double>;

#ifdef B2O_SIMT_EMU
inline void fence_proxy_async() {}
#else
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
#endif

template <int R, typename T = double>
__global__ void __launch_bounds__(B2O_NTHREADS, 1) qn_twoloop_kernel(const __grid_constant__ TwoLoopArgsT<T> p) {
  constexpr int EPT = R / B2O_NCONS;
#ifdef B2O_SIMT_EMU
  unsigned char *smem_raw = emu::dyn_smem();
#else
  extern __shared__ __align__(128) unsigned char smem_raw[];
#endif
  Ring rg;
  rg.buf = smem_raw;
  double *sred = reinterpret_cast<double *>(smem_raw + p.coef_off);  // [8] warp partials + [8..8+64) α_i
  double *alphas = sred + 8;
  rg.full = reinterpret_cast<uint64_t *>(smem_raw + p.bar_off);
  rg.empty = rg.full + p.stages;
  rg.stages = p.stages;
  __shared__ double s_dot;
  __shared__ double s_mb[2];
  __shared__ bool s_is_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_producer = warp == B2O_CONS_WARPS;
  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&rg.full[s], 1);
      mbar_init(&rg.empty[s], B2O_CONS_WARPS);
    }
    mbar_fence_init();
  }
  unsigned long long mb_epoch = p.mbox.epoch_base;
  const int A = p.nact;
  const bool fused = (p.sweep_end - p.sweep_begin) > 1 || (p.sweep_begin == 0 && p.sweep_end == 2 * A + 1);
  if (p.sweep_begin > 0) {
    // step mode: pick up the reduced dot and the α's of earlier launches
    if (tid == 0) s_dot = __ldcg(&p.dots[0]);
    for (int i = tid; i < A; i += B2O_NTHREADS) alphas[i] = __ldcg(&p.dots[1 + i]);
  }
  __syncthreads();

  RingPos pos(p.stages);
  const int64_t grid = gridDim.x;
  const int64_t my_tiles = (p.ntiles > (int64_t)blockIdx.x) ? (p.ntiles - 1 - blockIdx.x) / grid + 1 : 0;
  unsigned long long bar_target = p.bar_target;

  for (int w = p.sweep_begin; w < p.sweep_end; ++w) {
    // ---- describe sweep w (all threads compute the same thing)
    // w = 0          : d = s[0]·x
    // w = 1..A       : loop 1 step i=w  : q = qin - α_i y[i-1];  (i==A: q *= γ);  d = (i<A ? s[i] : y[A-1])·q
    // w = A+1..2A    : loop 2 step i=w-A: slot o = A-i;  β' = α_o - d/ys_o;  q = q + β' s[o];  (i<A: d = y[o-1]·q) else res
    const bool dot_only = (w == 0);
    const bool loop1 = (w >= 1 && w <= A);
    const bool last = (w == 2 * A);
    const T *v1 = nullptr, *v2 = nullptr;
    T c1 = (T)0;
    bool qin_is_x = false;
    bool apply_gamma = false;
    if (dot_only) {
      v2 = p.s[0];
      qin_is_x = true;
    } else if (loop1) {
      const int i = w;
      const T ak = (T)s_dot / (T)p.ys[i - 1];                     // αk = dot(s[k], q) / ys[k]      :133
      if (tid == 0) alphas[i - 1] = (double)ak;
      c1 = ak;
      v1 = p.y[i - 1];
      qin_is_x = (i == 1);
      apply_gamma = (i == A) && p.scaling;
      v2 = (i < A) ? p.s[i] : p.y[A - 1];
    } else {
      const int i = w - A, o = A - i;
      c1 = (T)alphas[o] - (T)s_dot / (T)p.ys[o];                  // β = αk - dot(y[k], q) / ys[k]  :144-145
      v1 = p.s[o];
      v2 = last ? nullptr : p.y[o - 1];
    }
    __syncthreads();  // alphas[] write above vs reads in later sweeps; s_dot reads vs next write

    if (is_producer) {
      if (lane == 0) {
        fence_proxy_async();
        for (int64_t i = 0; i < my_tiles; ++i) {
          const int64_t t = blockIdx.x + i * grid;
          if (!qin_is_x) producer_push<R>(rg, pos, p.q + t * R);
          if (v1) producer_push<R>(rg, pos, v1 + t * R);
          if (v2) producer_push<R>(rg, pos, v2 + t * R);
        }
      }
      __syncwarp();
    } else {
      double acc = 0.0;
      T q[EPT], xn[EPT], rold[EPT];
      const T alpha = (T)p.alpha, beta = (T)p.beta, gamma = (T)p.gamma;
      if (qin_is_x && my_tiles > 0) load_user_tile<R>(p.x, (int64_t)blockIdx.x * R, p.n, p.x_al16, xn);
      for (int64_t i = 0; i < my_tiles; ++i) {
        const int64_t t = blockIdx.x + i * grid;
        uint32_t q_slot = 0xffffffffu;
        if (qin_is_x) {
#pragma unroll
          for (int j = 0; j < EPT; ++j) q[j] = xn[j];                                  // q .= x   :127-128
          if (i + 1 < my_tiles) load_user_tile<R>(p.x, (t + grid) * R, p.n, p.x_al16, xn);
        } else {
          mbar_wait(&rg.full[pos.slot], pos.par);
          tile_from_ring<R>(rg, pos.slot, q);
          // the slot goes back only after the arithmetic below has consumed the loaded registers: an mbarrier arrive
          // issued behind still-in-flight LDS can let the TMA refill race the read (tests/test_sass_invariants.py)
          q_slot = pos.slot;
          pos.advance();
        }
        if (last && beta != (T)0) load_user_tile<R>(p.res, t * R, p.n, p.res_al16, rold);
        if (v1) {
          mbar_wait(&rg.full[pos.slot], pos.par);
          T v[EPT];
          tile_from_ring<R>(rg, pos.slot, v);
#pragma unroll
          for (int j = 0; j < EPT; ++j) {
            if (loop1) q[j] = q[j] - c1 * v[j];                                        // q .-= αk .* y[k]   :135
            else q[j] = q[j] + c1 * v[j];                                              // q .+= β .* s[k]    :146
          }
          if (q_slot != 0xffffffffu) consumer_release(rg, q_slot);
          consumer_release(rg, pos.slot);
          pos.advance();
          if (apply_gamma) {
#pragma unroll
            for (int j = 0; j < EPT; ++j) q[j] = q[j] * gamma;                         // q .*= γ            :139
          }
          if (last) {
#pragma unroll
            for (int j = 0; j < EPT; ++j) q[j] = (beta != (T)0) ? alpha * q[j] + beta * rold[j] : alpha * q[j];  // :149-153
            store_user_tile<R>(p.res, t * R, p.n, p.res_al16, q);
          } else {
            // q is library-owned: pitch is padded, rows >= n hold zeros and stay zero (x loads 0 there)
            tile_to_owned<R>(p.q + t * R, q);
          }
        }
        else if (q_slot != 0xffffffffu) consumer_release(rg, q_slot);   // (no sweep reads q without an update vector)
        if (v2) {
          mbar_wait(&rg.full[pos.slot], pos.par);
          T v[EPT];
          tile_from_ring<R>(rg, pos.slot, v);
          acc += tile_dot<EPT>(v, q);
          consumer_release(rg, pos.slot);
          pos.advance();
        }
      }
      if (v2) {
        double s = warp_sum(acc);
        if (lane == 0) sred[warp] = s;
      }
      fence_proxy_async();  // our generic-proxy stores to q precede next sweep's TMA reads of q
    }
    if (!v2) break;  // last sweep wrote res
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int wv = 0; wv < B2O_CONS_WARPS; ++wv) s += sred[wv];
      p.partials[blockIdx.x] = s;
    }
    if (fused && p.mbox.nranks > 1) {
      // one inner product, all-reduced across GPUs without leaving the kernel (NVLink peer mailbox)
      const unsigned long long epoch = ++mb_epoch;
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicAdd(p.bar, 1ULL);
      }
      if (blockIdx.x == 0) {
        unsigned long long t0 = 0, t1 = 0;
        if (tid == 0) {
          t0 = globaltimer_ns();
          while (ld_acquire_u64(p.bar) < bar_target) { __nanosleep(32); }
          __threadfence();
          t1 = globaltimer_ns();
        }
        __syncthreads();
        if (warp == 0) {
          double s = 0.0;
          for (int b = lane; b < (int)grid; b += 32) s += __ldcg(&p.partials[b]);
          s = warp_sum(s);
          if (lane == 0) s_mb[0] = s;
          __syncwarp();
          mbox_allreduce_warp(p.mbox, epoch, s_mb, 1);
          if (lane == 0) {
            s_dot = s_mb[0];
            p.dots[0] = s_mb[0];
            if (p.sweep_dots) p.sweep_dots[w] = s_mb[0];
            __threadfence();
            st_release_gpu_u64(p.mbox.ready, epoch);
            if (p.dbg) {
              const unsigned long long t2 = globaltimer_ns();
              p.dbg[0] += (double)(t1 - t0);
              p.dbg[1] += (double)(t2 - t1);
              p.dbg[2] += 1.0;
            }
          }
        }
      } else {
        if (tid == 0) {
          while (ld_acquire_u64(p.mbox.ready) < epoch) { __nanosleep(32); }
          s_dot = __ldcg(&p.dots[0]);
        }
      }
      bar_target += gridDim.x;
      __syncthreads();
    } else if (fused) {
      grid_barrier(p.bar, bar_target);
      bar_target += gridDim.x;
      if (warp == 0) {
        double s = 0.0;
        for (int b = lane; b < (int)grid; b += 32) s += __ldcg(&p.partials[b]);
        s = warp_sum(s);
        if (lane == 0) {
          s_dot = s;
          if (blockIdx.x == 0 && p.sweep_dots) p.sweep_dots[w] = s;
        }
      }
      __syncthreads();
    } else {
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        unsigned long long tk = atomicAdd(p.arrive, 1ULL);
        s_is_last = (tk == gridDim.x - 1);
      }
      __syncthreads();
      if (s_is_last) {
        __threadfence();
        if (warp == 0) {
          double s = 0.0;
          for (int b = lane; b < (int)grid; b += 32) s += __ldcg(&p.partials[b]);
          s = warp_sum(s);
          if (lane == 0) {
            p.dots[0] = s;
            *p.arrive = 0ULL;
          }
        }
        for (int i = tid; i < A; i += B2O_NTHREADS) p.dots[1 + i] = alphas[i];
      }
    }
  }
  if (p.alpha_out && blockIdx.x == 0 && p.sweep_end == 2 * A + 1)
    for (int i = tid; i < A; i += B2O_NTHREADS) p.alpha_out[i] = alphas[i];
}
