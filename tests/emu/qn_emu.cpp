// qn_emu.cpp -- host build of the streaming quasi-Newton kernels (csrc/b2o_stream.cuh, b2o_qn_kernels.cuh, b2o_qn_multi.cuh) under the
// SIMT emulator (TEST INFRASTRUCTURE).  One block of 288 OS threads runs the product's own kernel bodies: the TMA ring protocol
// (mbarrier phases, slot hand-back, producer / consumer item order), the sweep bookkeeping of the two-loop recursion and of its block
// variant, ragged tiles and unaligned user vectors are checked on a CPU box -- including rings with FEWER slots than a tile has items
// (the deadlock class that cost a GPU lease in round 2).  It says nothing about PTX, memory ordering or speed.
#include <atomic>
#include "simt_emu.h"

namespace emu_qn {
#include "../../linearoperators.jl_b200/csrc/b2o_shared_defs.h"
// ---- CUDA vocabulary of these headers that simt_emu.h does not carry
struct float4 {
  float x, y, z, w;
};
inline double2 make_double2(double a, double b) { return double2{a, b}; }
inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
inline double2 ldg_stream2(const double *p) { return double2{p[0], p[1]}; }
inline void stg_stream2(double *p, double2 v) {
  p[0] = v.x;
  p[1] = v.y;
}
inline unsigned __float_as_uint(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
using std::fma;
inline int __double2hiint(double d) {
  uint64_t u;
  memcpy(&u, &d, 8);
  return (int)(u >> 32);
}
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline unsigned long long ld_acquire_u64(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release_gpu_u64(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline unsigned long long globaltimer_ns() { return 0; }
inline void mbox_allreduce_warp(const MboxDev &, unsigned long long, double *, int) {
  fprintf(stderr, "emu: the NVLink mailbox is not emulated\n");
  abort();
}
// blocks run one after the other under the emulator: a grid barrier can only be met by a grid of ONE block
inline void grid_barrier(unsigned long long *ctr, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (gridDim.x != 1) {
      fprintf(stderr, "emu: grid barrier with %u blocks\n", gridDim.x);
      abort();
    }
    atomicAdd(ctr, 1ULL);
    while (ld_acquire_u64(ctr) < target) std::this_thread::yield();
  }
  __syncthreads();
}
#include "../../linearoperators.jl_b200/csrc/b2o_qn_multi.cuh"
}  // namespace emu_qn
using namespace emu_qn;
#define EMU_API __attribute__((visibility("default")))

namespace {
struct Work {
  std::vector<double> partials = std::vector<double>((size_t)B2O_MAX_GRID * B2O_MAX_COLS, 0.0), dots = std::vector<double>(B2O_WS_DOTS, 0.0);
  unsigned long long bar[2] = {0, 0};
};
constexpr int R = 1024;
}  // namespace

extern "C" {
// forward LBFGSOperator / LSR1Operator apply on ONE emulated CTA.  cols: [ncols][pitch] (pitch a multiple of R, zero padded, 16-byte aligned)
// compact representations (op 2 = OP_INV_COMPACT): coefficients = W * [cols' x]; base_div selects x/γ (compact forward form) or γx (compact inverse)
static const double *g_W = nullptr;
static int g_base_div = 0;
EMU_API void emu_qn_set_compact(const double *W, int base_div) {
  g_W = W;
  g_base_div = base_div;
}
EMU_API int emu_qn_compact(int op, int64_t n, int64_t pitch, int ncols, const double *cols, const double *cdiv, const double *x, double *res,
                           double alpha, double beta, double gamma, int scaling, int stages) {
  Work w;
  CompactArgsT<double> a;
  memset(&a, 0, sizeof(a));
  for (int c = 0; c < ncols; ++c) {
    a.cols[c] = cols + (size_t)c * pitch;
    a.cdiv[c] = cdiv ? cdiv[c] : 1.0;
  }
  a.ncols = ncols;
  a.x = x;
  a.res = res;
  a.n = n;
  a.ntiles = (n + R - 1) / R;
  a.alpha = alpha;
  a.beta = beta;
  a.gamma = gamma;
  a.scaling = scaling;
  a.x_al16 = ((uintptr_t)x % 16) == 0;
  a.res_al16 = ((uintptr_t)res % 16) == 0;
  a.partials = w.partials.data();
  a.dots = w.dots.data();
  a.bar = &w.bar[0];
  a.arrive = &w.bar[1];
  a.bar_target = 1;
  a.mode = ncols > 0 ? MODE_FUSED : MODE_PHASE2;
  a.stages = stages;
  a.group = std::max(1, std::min(ncols, 40));
  const SmemLayout L = smem_layout(R, stages, a.group);
  a.accs_off = (uint32_t)L.accs_off;
  a.coef_off = (uint32_t)L.coef_off;
  a.bar_off = (uint32_t)L.bar_off;
  a.mbox.nranks = 1;
  a.W = g_W;
  a.base_div = g_base_div;
  if (op == OP_LBFGS_FWD) {
    void (*k)(const CompactArgsT<double>) = qn_compact_kernel<R, OP_LBFGS_FWD, double>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), L.total, nullptr, a);
  } else if (op == OP_INV_COMPACT) {
    void (*k)(const CompactArgsT<double>) = qn_compact_kernel<R, OP_INV_COMPACT, double>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), L.total, nullptr, a);
  } else {
    void (*k)(const CompactArgsT<double>) = qn_compact_kernel<R, OP_LSR1, double>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), L.total, nullptr, a);
  }
  return 0;
}

// one step of the a_k rebuild inside push! (src/lbfgs.jl:236-250 -> OP_PUSH_A; src/lsr1.jl:166-181 -> OP_PUSH_L): writes the unnormalised a_k to
// `res` and returns the outgoing inner products (s_k.a_k; for L-SR1 also |a_k|^2) summed over the CTA partials like sum_partials_kernel
EMU_API int emu_qn_push_step(int lsr1, int64_t n, int64_t pitch, int ncols, const double *cols, const double *cdiv, const double *sk, const double *yk,
                             double *res, double gamma, int stages, double *out2) {
  Work w;
  CompactArgsT<double> a;
  memset(&a, 0, sizeof(a));
  for (int c = 0; c < ncols; ++c) {
    a.cols[c] = cols + (size_t)c * pitch;
    a.cdiv[c] = cdiv ? cdiv[c] : 1.0;
  }
  a.ncols = ncols;
  a.x = sk;
  a.res = res;
  a.y2 = yk;
  a.n = n;
  a.ntiles = (n + R - 1) / R;
  a.alpha = 1.0;
  a.beta = 0.0;
  a.gamma = gamma;
  a.scaling = 1;
  a.x_al16 = ((uintptr_t)sk % 16) == 0;
  a.res_al16 = ((uintptr_t)res % 16) == 0;
  a.partials = w.partials.data();
  a.dots = w.dots.data();
  a.bar = &w.bar[0];
  a.arrive = &w.bar[1];
  a.bar_target = 1;
  a.mode = ncols > 0 ? MODE_FUSED : MODE_PHASE2;
  a.stages = stages;
  a.group = std::max(1, std::min(ncols, 40));
  const SmemLayout L = smem_layout(R, stages, a.group);
  a.accs_off = (uint32_t)L.accs_off;
  a.coef_off = (uint32_t)L.coef_off;
  a.bar_off = (uint32_t)L.bar_off;
  a.mbox.nranks = 1;
  if (lsr1) {
    void (*k)(const CompactArgsT<double>) = qn_compact_kernel<R, OP_PUSH_L, double>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), L.total, nullptr, a);
  } else {
    void (*k)(const CompactArgsT<double>) = qn_compact_kernel<R, OP_PUSH_A, double>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), L.total, nullptr, a);
  }
  out2[0] = w.partials[(size_t)1 * ncols + 0];   // grid = 1: partials[grid*ncols + 2*cta + {0, 1}]
  out2[1] = w.partials[(size_t)1 * ncols + 1];
  return 0;
}

// the same kernels instantiated for Float32 (LBFGSOperator(Float32, n)): forward apply (inverse = 0) or two-loop recursion (inverse = 1).
// cols / S, Y: float [..][pitch]; q: float [pitch] zeroed
EMU_API int emu_qn_f32(int inverse, int64_t n, int64_t pitch, int ncols, const float *c0, const float *c1, const double *ys, const float *x, float *res,
                       float *q, double alpha, double beta, double gamma, int scaling, int stages) {
  Work w;
  if (!inverse) {
    CompactArgsT<float> a;
    memset(&a, 0, sizeof(a));
    for (int c = 0; c < ncols; ++c) {
      a.cols[c] = c0 + (size_t)c * pitch;
      a.cdiv[c] = 1.0;
    }
    a.ncols = ncols;
    a.x = x;
    a.res = res;
    a.n = n;
    a.ntiles = (n + R - 1) / R;
    a.alpha = alpha;
    a.beta = beta;
    a.gamma = gamma;
    a.scaling = scaling;
    a.x_al16 = ((uintptr_t)x % 16) == 0;
    a.res_al16 = ((uintptr_t)res % 16) == 0;
    a.partials = w.partials.data();
    a.dots = w.dots.data();
    a.bar = &w.bar[0];
    a.arrive = &w.bar[1];
    a.bar_target = 1;
    a.mode = ncols > 0 ? MODE_FUSED : MODE_PHASE2;
    a.stages = stages;
    a.group = std::max(1, std::min(ncols, 40));
    const SmemLayout L = smem_layout(R, stages, a.group, sizeof(float));
    a.accs_off = (uint32_t)L.accs_off;
    a.coef_off = (uint32_t)L.coef_off;
    a.bar_off = (uint32_t)L.bar_off;
    a.mbox.nranks = 1;
    void (*k)(const CompactArgsT<float>) = qn_compact_kernel<R, OP_LBFGS_FWD, float>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), L.total, nullptr, a);
    return 0;
  }
  TwoLoopArgsT<float> a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < ncols; ++i) {
    a.s[i] = c0 + (size_t)i * pitch;
    a.y[i] = c1 + (size_t)i * pitch;
    a.ys[i] = ys[i];
  }
  a.nact = ncols;
  a.x = x;
  a.res = res;
  a.q = q;
  a.n = n;
  a.ntiles = (n + R - 1) / R;
  a.alpha = alpha;
  a.beta = beta;
  a.gamma = gamma;
  a.scaling = scaling;
  a.x_al16 = ((uintptr_t)x % 16) == 0;
  a.res_al16 = ((uintptr_t)res % 16) == 0;
  a.partials = w.partials.data();
  a.dots = w.dots.data();
  a.bar = &w.bar[0];
  a.arrive = &w.bar[1];
  a.bar_target = 1;
  a.stages = stages;
  const SmemLayout L = smem_layout(R, stages, 0, sizeof(float));
  a.coef_off = (uint32_t)L.coef_off;
  a.bar_off = (uint32_t)L.bar_off;
  a.sweep_begin = 0;
  a.sweep_end = 2 * ncols + 1;
  a.mbox.nranks = 1;
  void (*k)(const TwoLoopArgsT<float>) = qn_twoloop_kernel<R, float>;
  B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), L.total, nullptr, a);
  return 0;
}

// block apply mul!(Res, op, X) of the forward operator (op 0) or L-SR1 (op 1): qn_multi_kernel<NR, OP> on one emulated CTA
EMU_API int emu_qn_multi(int op, int NR, int64_t n, int64_t pitch, int ncols, const double *cols, const double *cdiv, const double *x, int64_t ldx,
                         double *res, int64_t ldr, int nrhs, double alpha, double beta, double gamma, int scaling, int stages) {
  Work w;
  MultiArgs a;
  memset(&a, 0, sizeof(a));
  for (int c = 0; c < ncols; ++c) {
    a.cols[c] = cols + (size_t)c * pitch;
    a.cdiv[c] = cdiv ? cdiv[c] : 1.0;
  }
  a.ncols = ncols;
  a.x = x;
  a.res = res;
  a.ldx = ldx;
  a.ldr = ldr;
  a.nrhs = nrhs;
  a.n = n;
  const int RR = NR == 8 ? 1024 : 2048;
  a.ntiles = (n + RR - 1) / RR;
  a.alpha = alpha;
  a.beta = beta;
  a.gamma = gamma;
  a.scaling = scaling;
  a.x_al16 = ((uintptr_t)x % 16) == 0 && (nrhs == 1 || ldx % 2 == 0);
  a.res_al16 = ((uintptr_t)res % 16) == 0 && (nrhs == 1 || ldr % 2 == 0);
  const int nv = ncols * NR;
  a.partials = w.partials.data();
  a.dots = w.dots.data();
  a.bar = &w.bar[0];
  a.bar_target = 1;
  a.stages = stages;
  a.wacc_off = (uint32_t)((size_t)stages * RR * sizeof(double));
  a.coef_off = (uint32_t)(a.wacc_off + (size_t)B2O_CONS_WARPS * ncols * 32 * sizeof(double));
  a.bar_off = (uint32_t)(a.coef_off + (size_t)nv * sizeof(double));
  a.landed_off = (uint32_t)(a.bar_off + (size_t)2 * stages * sizeof(uint64_t));
  const size_t total = a.landed_off + B2O_NCONS * sizeof(unsigned);
  a.mbox.nranks = 1;
#define EMU_MULTI(NRv, OPv)                                                 \
  {                                                                         \
    void (*k)(const MultiArgs) = qn_multi_kernel<NRv, OPv>;                 \
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), total, nullptr, a);          \
    return 0;                                                               \
  }
  if (op == 0 && NR == 2) EMU_MULTI(2, OP_LBFGS_FWD)
  if (op == 0 && NR == 4) EMU_MULTI(4, OP_LBFGS_FWD)
  if (op == 0 && NR == 8) EMU_MULTI(8, OP_LBFGS_FWD)
  if (op == 1 && NR == 4) EMU_MULTI(4, OP_LSR1)
  if (op == 1 && NR == 8) EMU_MULTI(8, OP_LSR1)
#undef EMU_MULTI
  return 1;
}

// Float32 block apply of the forward operator: qn_multi_kernel<NR, LBFGS_FWD, float>
EMU_API int emu_qn_multi_f32(int NR, int64_t n, int64_t pitch, int ncols, const float *cols, const float *x, int64_t ldx, float *res, int64_t ldr, int nrhs,
                             double alpha, double beta, double gamma, int stages) {
  Work w;
  MultiArgsT<float> a;
  memset(&a, 0, sizeof(a));
  for (int c = 0; c < ncols; ++c) {
    a.cols[c] = cols + (size_t)c * pitch;
    a.cdiv[c] = 1.0;
  }
  a.ncols = ncols;
  a.x = x;
  a.res = res;
  a.ldx = ldx;
  a.ldr = ldr;
  a.nrhs = nrhs;
  a.n = n;
  const int RR = NR == 8 ? 2048 : 4096;
  a.ntiles = (n + RR - 1) / RR;
  a.alpha = alpha;
  a.beta = beta;
  a.gamma = gamma;
  a.scaling = 1;
  a.x_al16 = ((uintptr_t)x % 16) == 0 && (nrhs == 1 || ldx % 4 == 0);
  a.res_al16 = ((uintptr_t)res % 16) == 0 && (nrhs == 1 || ldr % 4 == 0);
  const int nv = ncols * NR;
  a.partials = w.partials.data();
  a.dots = w.dots.data();
  a.bar = &w.bar[0];
  a.bar_target = 1;
  a.stages = stages;
  a.wacc_off = (uint32_t)((size_t)stages * RR * sizeof(float));
  a.coef_off = (uint32_t)(a.wacc_off + (size_t)B2O_CONS_WARPS * ncols * 32 * sizeof(double));
  a.bar_off = (uint32_t)(a.coef_off + (size_t)nv * sizeof(double));
  a.landed_off = (uint32_t)(a.bar_off + (size_t)2 * stages * sizeof(uint64_t));
  const size_t total = a.landed_off + B2O_NCONS * sizeof(unsigned);
  a.mbox.nranks = 1;
  if (NR == 8) {
    void (*k)(const MultiArgsT<float>) = qn_multi_kernel<8, OP_LBFGS_FWD, float>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), total, nullptr, a);
  } else if (NR == 4) {
    void (*k)(const MultiArgsT<float>) = qn_multi_kernel<4, OP_LBFGS_FWD, float>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), total, nullptr, a);
  } else {
    void (*k)(const MultiArgsT<float>) = qn_multi_kernel<2, OP_LBFGS_FWD, float>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), total, nullptr, a);
  }
  return 0;
}

// inverse two-loop recursion, vector (nrhs == 0: x, res are vectors) or block (nrhs columns with leading dimensions ldx, ldr; NR = 4 or 8).
// S, Y: [A][pitch] newest -> oldest; q: [max(1, NR)][pitch] zeroed work vectors
EMU_API int emu_qn_twoloop(int64_t n, int64_t pitch, int A, const double *S, const double *Y, const double *ys, const double *x, int64_t ldx,
                           double *res, int64_t ldr, int nrhs, int NR, double *q, double alpha, double beta, double gamma, int scaling, int stages) {
  Work w;
  if (nrhs == 0) {
    TwoLoopArgsT<double> a;
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < A; ++i) {
      a.s[i] = S + (size_t)i * pitch;
      a.y[i] = Y + (size_t)i * pitch;
      a.ys[i] = ys[i];
    }
    a.nact = A;
    a.x = x;
    a.res = res;
    a.q = q;
    a.n = n;
    a.ntiles = (n + R - 1) / R;
    a.alpha = alpha;
    a.beta = beta;
    a.gamma = gamma;
    a.scaling = scaling;
    a.x_al16 = ((uintptr_t)x % 16) == 0;
    a.res_al16 = ((uintptr_t)res % 16) == 0;
    a.partials = w.partials.data();
    a.dots = w.dots.data();
    a.bar = &w.bar[0];
    a.arrive = &w.bar[1];
    a.bar_target = 1;
    a.stages = stages;
    const SmemLayout L = smem_layout(R, stages, 0);
    a.coef_off = (uint32_t)L.coef_off;
    a.bar_off = (uint32_t)L.bar_off;
    a.sweep_begin = 0;
    a.sweep_end = 2 * A + 1;
    a.mbox.nranks = 1;
    void (*k)(const TwoLoopArgsT<double>) = qn_twoloop_kernel<R, double>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), L.total, nullptr, a);
    return 0;
  }
  TwoLoopMultiArgs a;
  memset(&a, 0, sizeof(a));
  for (int i = 0; i < A; ++i) {
    a.s[i] = S + (size_t)i * pitch;
    a.y[i] = Y + (size_t)i * pitch;
    a.ys[i] = ys[i];
  }
  a.nact = A;
  a.x = x;
  a.res = res;
  a.ldx = ldx;
  a.ldr = ldr;
  a.nrhs = nrhs;
  a.q = q;
  a.qpitch = pitch;
  a.n = n;
  a.ntiles = (n + R - 1) / R;
  a.alpha = alpha;
  a.beta = beta;
  a.gamma = gamma;
  a.scaling = scaling;
  a.x_al16 = ((uintptr_t)x % 16) == 0 && (nrhs == 1 || ldx % 2 == 0);
  a.res_al16 = ((uintptr_t)res % 16) == 0 && (nrhs == 1 || ldr % 2 == 0);
  a.partials = w.partials.data();
  a.bar = &w.bar[0];
  a.bar_target = 1;
  a.stages = stages;
  const size_t scal = (size_t)(B2O_MAX_MEM + B2O_CONS_WARPS + 1) * NR * sizeof(double);
  a.scal_off = (uint32_t)((size_t)stages * R * sizeof(double));
  a.bar_off = (uint32_t)(a.scal_off + scal);
  a.landed_off = (uint32_t)(a.bar_off + (size_t)2 * stages * sizeof(uint64_t));
  const size_t total = a.landed_off + B2O_NCONS * sizeof(unsigned);
  if (NR == 4) {
    void (*k)(const TwoLoopMultiArgs) = qn_twoloop_multi_kernel<R, 4>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), total, nullptr, a);
  } else {
    void (*k)(const TwoLoopMultiArgs) = qn_twoloop_multi_kernel<R, 8>;
    B2O_LAUNCH(k, dim3(1), dim3(B2O_NTHREADS), total, nullptr, a);
  }
  return 0;
}
}
