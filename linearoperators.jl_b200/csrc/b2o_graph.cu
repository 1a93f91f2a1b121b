// b2o_graph.cu -- fused evaluator for STATIC operator trees.
//
// The reference evaluates `(opHouseholder(h)*opDiagonal(d) + 0.1*opEye(n)) * v` as a closure tree: prod_op!
// (src/operations.jl:117-128), sum_prod! (:187-197), scalar folding (:163-177), unary minus (:102-115) -- one memory
// pass per leaf statement plus temporaries.  Here the same tree is lowered ONCE on the host, by the same recursive
// descent `mul!` performs (so α/β threading and the statement-level rounding order are preserved), into
//   * per-row expression trees over the input vectors, and
//   * reductions (dot(h, ·) of opHouseholder, sum(·) of opOnes) that cut the evaluation into passes,
// and then into a tiny stack program per pass.  ONE cooperative kernel runs all passes: each thread preloads its rows
// of every array the pass touches (16-byte loads, all in flight together), interprets the program on a register
// resident stack (static indices via a jump table -- no local memory), and the passes are separated by the same
// deterministic grid barrier + fixed-order reduction as the quasi-Newton kernels.
// cfg3 lowers to 2 passes: 3n reads, then 3n reads + n writes = 7n*8 bytes, one launch.
#include "b2o_internal.cuh"
#include <math.h>
#include <algorithm>
#include <memory>

constexpr int G_MAX_PROG = 96, G_MAX_PASS = 4, G_MAX_ARR = 12, G_MAX_SCAL = 48, G_MAX_RED = 8;
constexpr int G_DEPTH = 6, G_SLOTS = 6, G_NT = 256, G_RED_PER_PASS = 4, G_MAX_SOP = 8;
// abstract ops used by the host code generator; the device sees FLAT opcodes (op, stack depth and array slot folded
// into one number so that one jump-table dispatch reaches code with static register indices)
enum { I_PUSH_ARR = 0, I_PUSH_SCAL = 1, I_ADD = 2, I_SUB = 3, I_MUL = 4, I_RED = 5, I_STORE = 6, I_BINA = 7, I_BINS = 8 };
enum { F_PUSHA = 0, F_PUSHS = 48, F_BIN = 56, F_BINA = 80, F_BINS = 224, F_RED = 248, F_STORE = 256 };
enum { GK_DIAG = 0, GK_EYE = 1, GK_ZEROS = 2, GK_ONES = 3, GK_HOUSE = 4, GK_SUM = 10, GK_PROD = 11, GK_SCALE = 12, GK_NEG = 13,
       GK_TRANS = 14 };

struct GraphArgs {
  const double *arr[G_MAX_ARR];
  unsigned char arr_al16[G_MAX_ARR];
  double *out;
  int out_al16;
  int64_t n;
  int npass;
  int prog_len[G_MAX_PASS];
  uint32_t prog[G_MAX_PASS][G_MAX_PROG];        // flat opcode | arg<<16
  int narr_pass[G_MAX_PASS];
  unsigned char arr_of_pass[G_MAX_PASS][G_SLOTS];
  double scal[G_MAX_SCAL];
  int nsop[G_MAX_PASS];
  unsigned char sop[G_MAX_PASS][G_MAX_SOP][3];  // before pass p: scal[dst] = scal[a] * red[r]
  int nred_pass[G_MAX_PASS];
  unsigned char red_of_pass[G_MAX_PASS][G_RED_PER_PASS];
  double *partials;
  double *dots;
  unsigned long long *bar;
  unsigned long long bar_target;
  unsigned long long *arrive;
  int pass_begin, pass_end, fused;
};

#define G_FOR_J _Pragma("unroll") for (int j = 0; j < G_EPT; ++j)
#define G_ROW6(M, A) M(A, 0) M(A, 1) M(A, 2) M(A, 3) M(A, 4) M(A, 5)
#define G_ROW6S(M, A) M(A, 1) M(A, 2) M(A, 3) M(A, 4) M(A, 5) M(A, 6)

// Two instantiations share the body: the SMALL machine (stack depth <= 3, <= 4 vectors per pass: cfg3 and every tree of that
// size) serves G_EPT = 4 rows per dispatch inside 128 registers; the general one (depth 6, 6 vectors) keeps 2 rows per dispatch.
// Opcodes are the same for both: static indices beyond a machine's limits are clamped (those cases are unreachable for
// programs the host routes to it).
#define SI(x) ((x) < DEPTH ? (x) : 0)
#define AI(x) ((x) < SLOTS ? (x) : 0)
template <int MINB, int DEPTH, int SLOTS, int G_EPT>
__global__ void __launch_bounds__(G_NT, MINB) graph_kernel(const __grid_constant__ GraphArgs p) {
  __shared__ double s_scal[G_MAX_SCAL];
  __shared__ double s_red[G_MAX_RED];
  __shared__ double s_w[G_NT / 32][G_RED_PER_PASS];
  __shared__ uint32_t s_prog[G_MAX_PROG + 1];
  __shared__ bool s_is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < G_MAX_SCAL; i += G_NT) s_scal[i] = p.scal[i];
  if (p.pass_begin > 0)
    for (int i = tid; i < G_MAX_RED; i += G_NT) s_red[i] = __ldcg(&p.dots[i]);
  __syncthreads();
  if (p.pass_begin > 0 && tid == 0)   // split mode: rebuild the reduction-derived scalars of earlier passes
    for (int q = 0; q < p.pass_begin; ++q)
      for (int i = 0; i < p.nsop[q]; ++i) s_scal[p.sop[q][i][0]] = s_scal[p.sop[q][i][1]] * s_red[p.sop[q][i][2]];
  __syncthreads();
  unsigned long long bar_target = p.bar_target;
  const int64_t tile_rows = (int64_t)G_NT * G_EPT;
  const int64_t ntiles = (p.n + tile_rows - 1) / tile_rows;

  for (int pass = p.pass_begin; pass < p.pass_end; ++pass) {
    if (tid == 0)
      for (int i = 0; i < p.nsop[pass]; ++i) s_scal[p.sop[pass][i][0]] = s_scal[p.sop[pass][i][1]] * s_red[p.sop[pass][i][2]];
    const int len = p.prog_len[pass];
    for (int i = tid; i <= len; i += G_NT) s_prog[i] = i < len ? p.prog[pass][i] : 0u;
    __syncthreads();
    const int narr = p.narr_pass[pass];
    double racc[G_RED_PER_PASS];
#pragma unroll
    for (int r = 0; r < G_RED_PER_PASS; ++r) racc[r] = 0.0;
    const bool reverse = pass & 1;  // alternate direction: the tail of the previous pass is still in L2

    for (int64_t tt = blockIdx.x; tt < ntiles; tt += gridDim.x) {
      const int64_t t = reverse ? (ntiles - 1 - tt) : tt;
      // this thread's rows: G_EPT/2 pairs, pair q at tile row 2*(q*G_NT + tid) (a warp reads 512 contiguous bytes per pair).
      // One interpreter dispatch now serves G_EPT = 4 rows (round 1: 2): the dispatch latency per row halves and twice the
      // loads are in flight per thread -- the interpreter was latency-bound (30 % DRAM throughput, profiles/r1_ncu_all_kernels.md).
      int64_t rq[G_EPT / 2];
      bool in[G_EPT];
#pragma unroll
      for (int q = 0; q < G_EPT / 2; ++q) {
        rq[q] = t * tile_rows + 2 * ((int64_t)q * G_NT + tid);
        in[2 * q] = rq[q] < p.n;
        in[2 * q + 1] = rq[q] + 1 < p.n;
      }
      double av[SLOTS][G_EPT];
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) {
        G_FOR_J av[s][j] = 0.0;
        if (s < narr) {
          const int a = p.arr_of_pass[pass][s];
          const double *base = p.arr[a];
#pragma unroll
          for (int q = 0; q < G_EPT / 2; ++q) {
            if (p.arr_al16[a] && in[2 * q + 1]) {
              double2 v = *reinterpret_cast<const double2 *>(base + rq[q]);
              av[s][2 * q] = v.x;
              av[s][2 * q + 1] = v.y;
            } else {
              if (in[2 * q]) av[s][2 * q] = base[rq[q]];
              if (in[2 * q + 1]) av[s][2 * q + 1] = base[rq[q] + 1];
            }
          }
        }
      }
      double st[DEPTH][G_EPT];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d) G_FOR_J st[d][j] = 0.0;

      uint32_t ins = s_prog[0];
      for (int pc = 0; pc < len; ++pc) {
        const uint32_t cur = ins;
        ins = s_prog[pc + 1];                       // prefetch the next instruction
        const int arg = cur >> 16;
        switch (cur & 0xffffu) {
#define C_PUSHA(SL, SP) case F_PUSHA + SL * 8 + SP: G_FOR_J st[SI(SP)][j] = av[AI(SL)][j]; break;
          G_ROW6(C_PUSHA, 0) G_ROW6(C_PUSHA, 1) G_ROW6(C_PUSHA, 2) G_ROW6(C_PUSHA, 3) G_ROW6(C_PUSHA, 4) G_ROW6(C_PUSHA, 5)
#define C_PUSHS(U, SP) case F_PUSHS + SP: { const double c = s_scal[arg]; G_FOR_J st[SI(SP)][j] = c; } break;
          G_ROW6(C_PUSHS, 0)
#define C_ADD(U, SP) case F_BIN + 0 * 8 + SP: G_FOR_J st[SI(SP - 2)][j] = st[SI(SP - 2)][j] + st[SI(SP - 1)][j]; break;
#define C_SUB(U, SP) case F_BIN + 1 * 8 + SP: G_FOR_J st[SI(SP - 2)][j] = st[SI(SP - 2)][j] - st[SI(SP - 1)][j]; break;
#define C_MUL(U, SP) case F_BIN + 2 * 8 + SP: G_FOR_J st[SI(SP - 2)][j] = st[SI(SP - 2)][j] * st[SI(SP - 1)][j]; break;
          C_ADD(0, 2) C_ADD(0, 3) C_ADD(0, 4) C_ADD(0, 5) C_ADD(0, 6)
          C_SUB(0, 2) C_SUB(0, 3) C_SUB(0, 4) C_SUB(0, 5) C_SUB(0, 6)
          C_MUL(0, 2) C_MUL(0, 3) C_MUL(0, 4) C_MUL(0, 5) C_MUL(0, 6)
          // top (op)= array slot
#define C_ADDA(SL, SP) case F_BINA + (0 * 6 + SL) * 8 + SP: G_FOR_J st[SI(SP - 1)][j] = st[SI(SP - 1)][j] + av[AI(SL)][j]; break;
#define C_SUBA(SL, SP) case F_BINA + (1 * 6 + SL) * 8 + SP: G_FOR_J st[SI(SP - 1)][j] = st[SI(SP - 1)][j] - av[AI(SL)][j]; break;
#define C_MULA(SL, SP) case F_BINA + (2 * 6 + SL) * 8 + SP: G_FOR_J st[SI(SP - 1)][j] = st[SI(SP - 1)][j] * av[AI(SL)][j]; break;
          G_ROW6S(C_ADDA, 0) G_ROW6S(C_ADDA, 1) G_ROW6S(C_ADDA, 2) G_ROW6S(C_ADDA, 3) G_ROW6S(C_ADDA, 4) G_ROW6S(C_ADDA, 5)
          G_ROW6S(C_SUBA, 0) G_ROW6S(C_SUBA, 1) G_ROW6S(C_SUBA, 2) G_ROW6S(C_SUBA, 3) G_ROW6S(C_SUBA, 4) G_ROW6S(C_SUBA, 5)
          G_ROW6S(C_MULA, 0) G_ROW6S(C_MULA, 1) G_ROW6S(C_MULA, 2) G_ROW6S(C_MULA, 3) G_ROW6S(C_MULA, 4) G_ROW6S(C_MULA, 5)
          // top (op)= scalar
#define C_ADDS(U, SP) case F_BINS + 0 * 8 + SP: { const double c = s_scal[arg]; G_FOR_J st[SI(SP - 1)][j] = st[SI(SP - 1)][j] + c; } break;
#define C_SUBS(U, SP) case F_BINS + 1 * 8 + SP: { const double c = s_scal[arg]; G_FOR_J st[SI(SP - 1)][j] = st[SI(SP - 1)][j] - c; } break;
#define C_MULS(U, SP) case F_BINS + 2 * 8 + SP: { const double c = s_scal[arg]; G_FOR_J st[SI(SP - 1)][j] = st[SI(SP - 1)][j] * c; } break;
          G_ROW6S(C_ADDS, 0) G_ROW6S(C_SUBS, 0) G_ROW6S(C_MULS, 0)
          // masked sum of the top of stack into local reduction `arg` (rows >= n contribute nothing)
#define C_RED(U, SP)                                                   \
  case F_RED + SP: {                                                   \
    double s = 0.0;                                                    \
    G_FOR_J s += in[j] ? st[SI(SP - 1)][j] : 0.0;                          \
    if (arg == 0) racc[0] += s;                                        \
    else if (arg == 1) racc[1] += s;                                   \
    else if (arg == 2) racc[2] += s;                                   \
    else racc[3] += s;                                                 \
  } break;
          G_ROW6S(C_RED, 0)
#define C_STORE(U, SP)                                                                       \
  case F_STORE + SP: {                                                                       \
    _Pragma("unroll") for (int q = 0; q < G_EPT / 2; ++q) {                                   \
      if (p.out_al16 && in[2 * q + 1]) stg_stream2(p.out + rq[q], make_double2(st[SI(SP - 1)][2 * q], st[SI(SP - 1)][2 * q + 1])); \
      else {                                                                                 \
        if (in[2 * q]) p.out[rq[q]] = st[SI(SP - 1)][2 * q];                                     \
        if (in[2 * q + 1]) p.out[rq[q] + 1] = st[SI(SP - 1)][2 * q + 1];                         \
      }                                                                                      \
    }                                                                                        \
  } break;
          G_ROW6S(C_STORE, 0)
          default: break;
        }
      }
    }

    const int nred = p.nred_pass[pass];
    if (nred == 0) continue;   // only the final pass has no reductions
#pragma unroll
    for (int r = 0; r < G_RED_PER_PASS; ++r) {
      double s = warp_sum(racc[r]);
      if (lane == 0) s_w[warp][r] = s;
    }
    __syncthreads();
    if (tid < nred) {
      double s = 0.0;
      for (int w = 0; w < G_NT / 32; ++w) s += s_w[w][tid];
      p.partials[(size_t)blockIdx.x * G_RED_PER_PASS + tid] = s;
    }
    if (p.fused) {
      grid_barrier(p.bar, bar_target);
      bar_target += gridDim.x;
      if (warp < nred) {
        double s = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(&p.partials[(size_t)b * G_RED_PER_PASS + warp]);
        s = warp_sum(s);
        if (lane == 0) s_red[p.red_of_pass[pass][warp]] = s;
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
        if (warp < nred) {
          double s = 0.0;
          for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(&p.partials[(size_t)b * G_RED_PER_PASS + warp]);
          s = warp_sum(s);
          if (lane == 0) p.dots[p.red_of_pass[pass][warp]] = s;
        }
        if (tid == 0) *p.arrive = 0ULL;
      }
    }
  }
}

#undef SI
#undef AI
// ====================================================================================== host: tree -> passes -> stack programs
namespace {

struct GNode {
  int kind, a, b;
  double x;
  const double *ptr;
};

// symbolic scalars: evaluated on the host at apply time unless they involve a reduction
struct SExpr {
  enum Kind { CONST, ALPHA, BETA, MUL, NEG, REDMUL } kind;
  double c = 0;
  int a = -1, b = -1;   // MUL: a*b ; NEG: -a ; REDMUL: a * red[b]
};
// per-row values
struct VExpr {
  enum Kind { ARR, SCAL, BIN } kind;
  int id = -1;          // ARR: array id ; SCAL: scalar id
  int op = 0, l = -1, r = -1;
};
struct Red {
  int expr;             // VExpr summed over rows
  int level = 0;
};

struct Lowering {
  std::vector<SExpr> S;
  std::vector<VExpr> V;
  std::vector<Red> R;
  std::vector<const double *> arrays;   // 0 = v (input), 1 = res (in/out), then leaves
  int s_one, s_zero;

  int sc(SExpr e) { S.push_back(e); return (int)S.size() - 1; }
  int sconst(double c) { SExpr e; e.kind = SExpr::CONST; e.c = c; return sc(e); }
  int smul(int a, int b) {
    if (is_one(a)) return b;
    if (is_one(b)) return a;
    SExpr e; e.kind = SExpr::MUL; e.a = a; e.b = b; return sc(e);
  }
  int sneg(int a) { SExpr e; e.kind = SExpr::NEG; e.a = a; return sc(e); }
  int sredmul(int a, int red) { SExpr e; e.kind = SExpr::REDMUL; e.a = a; e.b = red; return sc(e); }
  bool is_one(int s) const { return S[s].kind == SExpr::CONST && S[s].c == 1.0; }
  bool is_zero(int s) const { return s == s_zero; }

  int varr(int id) { VExpr e; e.kind = VExpr::ARR; e.id = id; V.push_back(e); return (int)V.size() - 1; }
  int vscal(int s) { VExpr e; e.kind = VExpr::SCAL; e.id = s; V.push_back(e); return (int)V.size() - 1; }
  int vbin(int op, int l, int r) { VExpr e; e.kind = VExpr::BIN; e.op = op; e.l = l; e.r = r; V.push_back(e); return (int)V.size() - 1; }
  int array_id(const double *p) {
    for (size_t i = 2; i < arrays.size(); ++i)
      if (arrays[i] == p) return (int)i;
    arrays.push_back(p);
    return (int)arrays.size() - 1;
  }
  // scalar*vector with the bit-exact shortcut 1*x == x
  int vscale(int s, int v) { return is_one(s) ? v : vbin(I_MUL, vscal(s), v); }
  // t (+ β*res): `.+ β .* res`
  int with_beta(int t, int beta, int res) {
    if (is_zero(beta)) return t;
    return vbin(I_ADD, t, vscale(beta, res));
  }
  int red_level_of_scalar(int s) const {
    const SExpr &e = S[s];
    switch (e.kind) {
      case SExpr::MUL: return std::max(red_level_of_scalar(e.a), red_level_of_scalar(e.b));
      case SExpr::NEG: return red_level_of_scalar(e.a);
      case SExpr::REDMUL: return std::max(red_level_of_scalar(e.a), R[e.b].level);
      default: return 0;
    }
  }
  int level_of(int v) const {
    const VExpr &e = V[v];
    if (e.kind == VExpr::ARR) return 0;
    if (e.kind == VExpr::SCAL) return red_level_of_scalar(e.id);
    return std::max(level_of(e.l), level_of(e.r));
  }
  int new_red(int expr) {
    Red r; r.expr = expr; r.level = level_of(expr) + 1;
    R.push_back(r);
    return (int)R.size() - 1;
  }

  // the recursive descent of mul!(res, node, in, α, β): returns the VExpr of the new res
  int emit(const std::vector<GNode> &N, int node, int in, int alpha, int beta, int res, bool tr) {
    const GNode &g = N[node];
    switch (g.kind) {
      case GK_DIAG: {   // res .= α .* d .* v (.+ β .* res)            special-operators.jl:125-131
        int t = vbin(I_MUL, vscale(alpha, varr(array_id(g.ptr))), in);
        return with_beta(t, beta, res);
      }
      case GK_EYE:      // res .= α .* v (.+ β .* res)                  :36-44 (square)
        return with_beta(vscale(alpha, in), beta, res);
      case GK_ZEROS:    // res .= 0 | res .*= β                          :102-108
        return is_zero(beta) ? vscal(s_zero) : vbin(I_MUL, res, vscal(beta));
      case GK_ONES: {   // res .= (α * sum(v)) (.+ β .* res)             :79-85
        int r = new_red(in);
        return with_beta(vscal(sredmul(alpha, r)), beta, res);
      }
      case GK_HOUSE: {  // res .= α .* (v .- 2 * dot(h, v) .* h) (.+ β .* res)   linalg.jl:77-83
        int h = array_id(g.ptr);
        int r = new_red(vbin(I_MUL, varr(h), in));
        int t2 = sredmul(sconst(2.0), r);
        int t = vscale(alpha, vbin(I_SUB, in, vbin(I_MUL, vscal(t2), varr(h))));
        return with_beta(t, beta, res);
      }
      case GK_SUM: {    // sum_prod!                                    operations.jl:187-197
        int r1 = emit(N, g.a, in, alpha, beta, res, tr);
        return emit(N, g.b, in, alpha, s_one, r1, tr);
      }
      case GK_PROD: {   // prod_op!                                     operations.jl:117-128 (transpose swaps the factors)
        int first = tr ? g.a : g.b, second = tr ? g.b : g.a;
        int vt = emit(N, first, in, s_one, s_zero, -1, tr);
        return emit(N, second, vt, alpha, beta, res, tr);
      }
      case GK_SCALE:    // mul!(res, op, v, x * α, β)                   operations.jl:163-177
        return emit(N, g.a, in, smul(sconst(g.x), alpha), beta, res, tr);
      case GK_NEG:      // mul!(res, op, v, -α, β)                      operations.jl:102-115
        return emit(N, g.a, in, sneg(alpha), beta, res, tr);
      case GK_TRANS:
        return emit(N, g.a, in, alpha, beta, res, !tr);
    }
    return -1;
  }
};

struct Compiled {
  bool valid = false;
  GraphArgs args;
  std::vector<SExpr> S;                 // scalar expressions; slot i of args.scal
  int alg_arrays_read = 0;
  int max_depth = 0, max_slots = 0;     // stack depth / vectors per pass the programs need (selects the interpreter instantiation)
  // specialised instantiation (NVRTC): the same passes as straight-line code
  std::string jit_src;
  void *jit_module = nullptr;           // CUmodule
  void *jit_func = nullptr;             // CUfunction
  int jit_state = 0;                    // 0 not tried, 1 NVRTC module ready, 2 ahead-of-time instantiation found, -1 unavailable (interpreter)
  const void *aot_fn = nullptr;         // kernel of csrc/b2o_graph_aot.cu whose source hash matches jit_src
  int jit_blocks_per_sm = 0;
};

}  // namespace

struct b2o_graph_s {
  b2o_ctx *ctx;
  int64_t n;
  std::vector<GNode> nodes;
  int root = -1;
  Compiled prog[2][2];                  // [transposed][beta != 0]
  std::string err;
};

static int slot_of(std::vector<int> &slots, int id, std::string &err) {
  for (size_t i = 0; i < slots.size(); ++i)
    if (slots[i] == id) return (int)i;
  if ((int)slots.size() >= G_SLOTS) { err = "too many distinct vectors in one pass"; return -1; }
  slots.push_back(id);
  return (int)slots.size() - 1;
}
static inline int op3(int op) { return op == I_ADD ? 0 : op == I_SUB ? 1 : 2; }

// post-order code generation; binary ops whose right (or, if commutative, left) operand is a leaf fold the operand
// into the instruction (st op= array / scalar): same arithmetic, fewer dispatches.
static bool gen_code(Lowering &L, int v, std::vector<uint32_t> &code, int &sp, int &maxsp, std::vector<int> &slots,
                     std::string &err) {
  const VExpr &e = L.V[v];
  if (e.kind == VExpr::BIN) {
    int lhs = e.l, rhs = e.r;
    const bool commut = e.op != I_SUB;
    if (L.V[rhs].kind == VExpr::BIN && L.V[lhs].kind != VExpr::BIN && commut) std::swap(lhs, rhs);  // a+b == b+a, a*b == b*a bitwise
    if (L.V[rhs].kind != VExpr::BIN) {
      if (!gen_code(L, lhs, code, sp, maxsp, slots, err)) return false;
      if (L.V[rhs].kind == VExpr::ARR) {
        int sl = slot_of(slots, L.V[rhs].id, err);
        if (sl < 0) return false;
        code.push_back((uint32_t)(F_BINA + (op3(e.op) * 6 + sl) * 8 + sp));
      } else {
        code.push_back((uint32_t)(F_BINS + op3(e.op) * 8 + sp) | ((uint32_t)L.V[rhs].id << 16));
      }
      return true;
    }
    if (!gen_code(L, lhs, code, sp, maxsp, slots, err)) return false;
    if (!gen_code(L, rhs, code, sp, maxsp, slots, err)) return false;
    code.push_back((uint32_t)(F_BIN + op3(e.op) * 8 + sp));
    sp -= 1;
    return true;
  }
  if (sp >= G_DEPTH) { err = "expression too deep for the fused evaluator"; return false; }
  if (e.kind == VExpr::ARR) {
    int sl = slot_of(slots, e.id, err);
    if (sl < 0) return false;
    code.push_back((uint32_t)(F_PUSHA + sl * 8 + sp));
  } else {
    code.push_back((uint32_t)(F_PUSHS + sp) | ((uint32_t)e.id << 16));
  }
  sp += 1;
  maxsp = std::max(maxsp, sp);
  return true;
}


// ====================================================================================== specialised instantiation via NVRTC
// The interpreter above pays one dispatch per op per tile.  For a STATIC tree the passes are known when the graph is
// compiled, so the same hand-written pass template (vector loads, masked reductions, deterministic grid barrier) is
// instantiated with the per-row expressions written out as straight-line __dmul_rn/__dadd_rn/__dsub_rn calls and
// compiled for sm_100a with NVRTC.  No NVRTC / driver -> the interpreter kernel remains the (all-CUDA) path.
#include <dlfcn.h>
#include <sstream>
namespace {

struct JitRT {          // must match `struct RT` in the generated source
  const double *arr[G_MAX_ARR];
  double *out;
  long long n;
  double scal[G_MAX_SCAL];
  double *partials;
  double *dots;
  unsigned long long *bar;
  unsigned long long bar_target;
  unsigned long long *arrive;
  int pass_begin, pass_end, fused, vec;
};

struct JitApi {
  bool tried = false, ok = false;
  // nvrtc
  int (*CreateProgram)(void **, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  int (*CompileProgram)(void *, int, const char *const *) = nullptr;
  int (*GetCUBINSize)(void *, size_t *) = nullptr;
  int (*GetCUBIN)(void *, char *) = nullptr;
  int (*GetProgramLogSize)(void *, size_t *) = nullptr;
  int (*GetProgramLog)(void *, char *) = nullptr;
  int (*DestroyProgram)(void **) = nullptr;
  // driver
  int (*ModuleLoadData)(void **, const void *) = nullptr;
  int (*ModuleGetFunction)(void **, void *, const char *) = nullptr;
  int (*ModuleUnload)(void *) = nullptr;
  int (*LaunchCooperativeKernel)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **) = nullptr;
  int (*LaunchKernel)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **, void **) = nullptr;
  int (*OccupancyMaxActiveBlocksPerMultiprocessor)(int *, void *, int, size_t) = nullptr;
  bool have_nvrtc = false, have_driver = false;
};
static JitApi g_jit;

static void jit_load_api() {
  if (g_jit.tried) return;
  g_jit.tried = true;
  void *n = dlopen("libnvrtc.so.12", RTLD_NOW);
  if (!n) n = dlopen("libnvrtc.so", RTLD_NOW);
  if (n) {
    *(void **)&g_jit.CreateProgram = dlsym(n, "nvrtcCreateProgram");
    *(void **)&g_jit.CompileProgram = dlsym(n, "nvrtcCompileProgram");
    *(void **)&g_jit.GetCUBINSize = dlsym(n, "nvrtcGetCUBINSize");
    *(void **)&g_jit.GetCUBIN = dlsym(n, "nvrtcGetCUBIN");
    *(void **)&g_jit.GetProgramLogSize = dlsym(n, "nvrtcGetProgramLogSize");
    *(void **)&g_jit.GetProgramLog = dlsym(n, "nvrtcGetProgramLog");
    *(void **)&g_jit.DestroyProgram = dlsym(n, "nvrtcDestroyProgram");
    g_jit.have_nvrtc = g_jit.CreateProgram && g_jit.CompileProgram && g_jit.GetCUBINSize && g_jit.GetCUBIN && g_jit.DestroyProgram;
  }
  void *d = dlopen("libcuda.so.1", RTLD_NOW);
  if (d) {
    *(void **)&g_jit.ModuleLoadData = dlsym(d, "cuModuleLoadData");
    *(void **)&g_jit.ModuleGetFunction = dlsym(d, "cuModuleGetFunction");
    *(void **)&g_jit.ModuleUnload = dlsym(d, "cuModuleUnload");
    *(void **)&g_jit.LaunchCooperativeKernel = dlsym(d, "cuLaunchCooperativeKernel");
    *(void **)&g_jit.LaunchKernel = dlsym(d, "cuLaunchKernel");
    *(void **)&g_jit.OccupancyMaxActiveBlocksPerMultiprocessor = dlsym(d, "cuOccupancyMaxActiveBlocksPerMultiprocessor");
    g_jit.have_driver = g_jit.ModuleLoadData && g_jit.ModuleGetFunction && g_jit.LaunchCooperativeKernel && g_jit.LaunchKernel &&
                        g_jit.OccupancyMaxActiveBlocksPerMultiprocessor;
  }
  g_jit.ok = g_jit.have_nvrtc && g_jit.have_driver;
}

static void emit_expr(const Lowering &L, int v, const std::vector<int> &slots, std::ostringstream &o) {
  const VExpr &e = L.V[v];
  if (e.kind == VExpr::ARR) {
    int sl = 0;
    for (size_t i = 0; i < slots.size(); ++i)
      if (slots[i] == e.id) sl = (int)i;
    o << "a" << sl;
  } else if (e.kind == VExpr::SCAL) {
    o << "S[" << e.id << "]";
  } else {
    o << (e.op == I_ADD ? "__dadd_rn(" : e.op == I_SUB ? "__dsub_rn(" : "__dmul_rn(");
    emit_expr(L, e.l, slots, o);
    o << ", ";
    emit_expr(L, e.r, slots, o);
    o << ")";
  }
}
static void collect_slots(const Lowering &L, int v, std::vector<int> &slots) {
  const VExpr &e = L.V[v];
  if (e.kind == VExpr::ARR) {
    for (int s : slots)
      if (s == e.id) return;
    slots.push_back(e.id);
  } else if (e.kind == VExpr::BIN) {
    collect_slots(L, e.l, slots);
    collect_slots(L, e.r, slots);
  }
}

static const char *kJitPrologue = R"SRC(
typedef unsigned long long u64;
struct RT {
  const double *arr[12]; double *out; long long n; double scal[48];
  double *partials; double *dots; u64 *bar; u64 bar_target; u64 *arrive;
  int pass_begin, pass_end, fused, vec;
};
__device__ __forceinline__ double wsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ u64 ldacq(const u64 *p) { u64 v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void gbar(u64 *ctr, u64 target) {
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); atomicAdd(ctr, 1ULL); while (ldacq(ctr) < target) { __nanosleep(32); } __threadfence(); }
  __syncthreads();
}
__device__ __forceinline__ void st2(double *p, double x, double y) { asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" :: "l"(p), "d"(x), "d"(y) : "memory"); }
)SRC";

// emits the CUDA source of the specialised kernel for one compiled variant
static std::string jit_source(const Lowering &L, const GraphArgs &A, int out_expr) {
  std::ostringstream o;
  o << kJitPrologue;
  const int npass = A.npass;
  std::vector<std::vector<int>> pass_slots(npass);
  std::vector<std::vector<int>> pass_reds(npass);
  for (int p = 0; p < npass; ++p) {
    for (size_t r = 0; r < L.R.size(); ++r)
      if (L.R[r].level == p + 1) {
        pass_reds[p].push_back((int)r);
        collect_slots(L, L.R[r].expr, pass_slots[p]);
      }
    if (p == npass - 1) collect_slots(L, out_expr, pass_slots[p]);
  }
  auto params = [&](int p) {
    std::ostringstream q;
    for (size_t i = 0; i < pass_slots[p].size(); ++i) q << "double a" << i << ", ";
    q << "const double *S";
    return q.str();
  };
  auto argsj = [&](int p, const char *suffix) {
    std::ostringstream q;
    for (size_t i = 0; i < pass_slots[p].size(); ++i) q << "x" << i << suffix << ", ";
    q << "S";
    return q.str();
  };
  for (int p = 0; p < npass; ++p) {
    for (size_t k = 0; k < pass_reds[p].size(); ++k) {
      o << "__device__ __forceinline__ double f_p" << p << "_r" << k << "(" << params(p) << ") { return ";
      emit_expr(L, L.R[pass_reds[p][k]].expr, pass_slots[p], o);
      o << "; }\n";
    }
    if (p == npass - 1) {
      o << "__device__ __forceinline__ double f_out(" << params(p) << ") { return ";
      emit_expr(L, out_expr, pass_slots[p], o);
      o << "; }\n";
    }
  }
  o << "extern \"C\" __global__ void __launch_bounds__(256) b2o_fused(const __grid_constant__ RT p) {\n"
       "  __shared__ double S[48]; __shared__ double R[8]; __shared__ double W[8][4]; __shared__ bool s_last;\n"
       "  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;\n"
       "  for (int i = tid; i < 48; i += 256) S[i] = p.scal[i];\n"
       "  if (p.pass_begin > 0) for (int i = tid; i < 8; i += 256) R[i] = __ldcg(&p.dots[i]);\n"
       "  __syncthreads();\n"
       "  u64 bt = p.bar_target;\n"
       "  const long long npair = p.n >> 1;\n"
       "  const long long stride = (long long)gridDim.x * 256;\n"
       "  const long long nit = (npair + stride - 1) / stride;   // grid-stride iterations over 16-byte pairs\n";
  for (int p = 0; p < npass; ++p) {
    const int ns = (int)pass_slots[p].size(), nr = (int)pass_reds[p].size();
    const bool last = p == npass - 1;
    // scalars that depend on reductions finished before this pass (recomputed when a launch starts at a later pass)
    o << "  if (tid == 0 && " << p << " < p.pass_end) {\n";
    for (int i = 0; i < A.nsop[p]; ++i)
      o << "    S[" << (int)A.sop[p][i][0] << "] = __dmul_rn(S[" << (int)A.sop[p][i][1] << "], R[" << (int)A.sop[p][i][2] << "]);\n";
    o << "  }\n  __syncthreads();\n";
    o << "  if (p.pass_begin <= " << p << " && " << p << " < p.pass_end) {\n";
    for (int i = 0; i < ns; ++i) o << "    const double *A" << i << " = p.arr[" << pass_slots[p][i] << "];\n";
    for (int k = 0; k < nr; ++k) o << "    double acc" << k << " = 0.0;\n";
    const bool rev = p & 1;
    o << "    if (p.vec) {\n"
         "      for (long long it0 = 0; it0 < nit; it0 += 2) {\n"
         "        const long long ita = " << (rev ? "nit - 1 - it0" : "it0") << ", itb = " << (rev ? "nit - 2 - it0" : "it0 + 1") << ";\n"
         "        const long long ea = ita * stride + (long long)blockIdx.x * 256 + tid;\n"
         "        const long long eb = itb * stride + (long long)blockIdx.x * 256 + tid;\n"
         "        const bool va = ea < npair, vb = " << (rev ? "itb >= 0" : "itb < nit") << " && eb < npair;\n";
    for (int i = 0; i < ns; ++i)
      o << "        double2 u" << i << "a = make_double2(0.0, 0.0), u" << i << "b = make_double2(0.0, 0.0);\n";
    for (int i = 0; i < ns; ++i) o << "        if (va) u" << i << "a = *reinterpret_cast<const double2 *>(A" << i << " + 2 * ea);\n";
    for (int i = 0; i < ns; ++i) o << "        if (vb) u" << i << "b = *reinterpret_cast<const double2 *>(A" << i << " + 2 * eb);\n";
    for (const char *h : {"a", "b"}) {
      o << "        if (v" << h << ") {\n";
      for (int i = 0; i < ns; ++i) o << "          const double x" << i << "0 = u" << i << h << ".x, x" << i << "1 = u" << i << h << ".y;\n";
      for (int k = 0; k < nr; ++k)
        o << "          acc" << k << " += f_p" << p << "_r" << k << "(" << argsj(p, "0") << ") + f_p" << p << "_r" << k << "(" << argsj(p, "1") << ");\n";
      if (last) o << "          st2(p.out + 2 * e" << h << ", f_out(" << argsj(p, "0") << "), f_out(" << argsj(p, "1") << "));\n";
      o << "        }\n";
    }
    o << "      }\n"
         "      if ((p.n & 1) && blockIdx.x == 0 && tid == 0) {\n"
         "        const long long r = p.n - 1;\n";
    for (int i = 0; i < ns; ++i) o << "        const double x" << i << "0 = A" << i << "[r];\n";
    for (int k = 0; k < nr; ++k) o << "        acc" << k << " += f_p" << p << "_r" << k << "(" << argsj(p, "0") << ");\n";
    if (last) o << "        p.out[r] = f_out(" << argsj(p, "0") << ");\n";
    o << "      }\n"
         "    } else {\n"
         "      for (long long r = (long long)blockIdx.x * 256 + tid; r < p.n; r += stride) {\n";
    for (int i = 0; i < ns; ++i) o << "        const double x" << i << "0 = A" << i << "[r];\n";
    for (int k = 0; k < nr; ++k) o << "        acc" << k << " += f_p" << p << "_r" << k << "(" << argsj(p, "0") << ");\n";
    if (last) o << "        p.out[r] = f_out(" << argsj(p, "0") << ");\n";
    o << "      }\n    }\n";
    if (nr > 0) {
      for (int k = 0; k < nr; ++k) o << "    { double s = wsum(acc" << k << "); if (lane == 0) W[warp][" << k << "] = s; }\n";
      o << "    __syncthreads();\n"
           "    if (tid < " << nr << ") { double s = 0.0; for (int w = 0; w < 8; ++w) s += W[w][tid]; p.partials[(size_t)blockIdx.x * 4 + tid] = s; }\n"
           "    if (p.fused) {\n"
           "      gbar(p.bar, bt); bt += gridDim.x;\n"
           "      if (warp < " << nr << ") {\n"
           "        double s = 0.0; for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(&p.partials[(size_t)b * 4 + warp]);\n"
           "        s = wsum(s);\n"
           "        if (lane == 0) { const int ids[4] = {";
      for (int k = 0; k < 4; ++k) o << (k < nr ? pass_reds[p][k] : 0) << (k < 3 ? ", " : "");
      o << "}; R[ids[warp]] = s; }\n"
           "      }\n"
           "      __syncthreads();\n"
           "    } else {\n"
           "      __threadfence(); __syncthreads();\n"
           "      if (tid == 0) { u64 tk = atomicAdd(p.arrive, 1ULL); s_last = (tk == gridDim.x - 1); }\n"
           "      __syncthreads();\n"
           "      if (s_last) {\n"
           "        __threadfence();\n"
           "        if (warp < " << nr << ") {\n"
           "          double s = 0.0; for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(&p.partials[(size_t)b * 4 + warp]);\n"
           "          s = wsum(s);\n"
           "          if (lane == 0) { const int ids[4] = {";
      for (int k = 0; k < 4; ++k) o << (k < nr ? pass_reds[p][k] : 0) << (k < 3 ? ", " : "");
      o << "}; p.dots[ids[warp]] = s; }\n"
           "        }\n"
           "        if (tid == 0) *p.arrive = 0ULL;\n"
           "      }\n"
           "    }\n";
    }
    o << "  }\n";
  }
  o << "}\n";
  return o.str();
}

// NVRTC: source -> sm_100a cubin.  Usable on a CPU-only box (this is what tests/test_abi.py checks).
static int jit_compile_cubin(const std::string &src, std::vector<char> &cubin, std::string &log) {
  jit_load_api();
  if (!g_jit.have_nvrtc) { log = "libnvrtc not available"; return -1; }
  void *prog = nullptr;
  if (g_jit.CreateProgram(&prog, src.c_str(), "b2o_fused.cu", 0, nullptr, nullptr) != 0) { log = "nvrtcCreateProgram failed"; return -1; }
  const char *opts[] = {"--gpu-architecture=sm_100a", "--fmad=false", "-lineinfo", "--std=c++17"};
  int rc = g_jit.CompileProgram(prog, 4, opts);
  if (rc != 0) {
    size_t ls = 0;
    if (g_jit.GetProgramLogSize && g_jit.GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
      log.resize(ls);
      g_jit.GetProgramLog(prog, &log[0]);
    }
    g_jit.DestroyProgram(&prog);
    return -1;
  }
  size_t cs = 0;
  g_jit.GetCUBINSize(prog, &cs);
  cubin.resize(cs);
  g_jit.GetCUBIN(prog, cubin.data());
  g_jit.DestroyProgram(&prog);
  return 0;
}

}  // namespace

static int compile_variant(b2o_graph *g, bool tr, bool beta_nz, Compiled &C) {
  Lowering L;
  L.arrays = {nullptr, nullptr};
  L.s_zero = L.sconst(0.0);
  L.s_one = L.sconst(1.0);
  SExpr ea; ea.kind = SExpr::ALPHA; int s_alpha = L.sc(ea);
  SExpr eb; eb.kind = SExpr::BETA; int s_beta = beta_nz ? L.sc(eb) : L.s_zero;
  int in = L.varr(0), res = beta_nz ? L.varr(1) : -1;
  int out = L.emit(g->nodes, g->root, in, s_alpha, s_beta, res, tr);
  if (out < 0) B2O_FAIL(B2O_EUNSUPPORTED, "graph: node kind cannot be fused");
  if ((int)L.S.size() > G_MAX_SCAL) B2O_FAIL(B2O_EUNSUPPORTED, "graph: too many scalars");
  if ((int)L.R.size() > G_MAX_RED) B2O_FAIL(B2O_EUNSUPPORTED, "graph: too many reductions");
  if ((int)L.arrays.size() > G_MAX_ARR) B2O_FAIL(B2O_EUNSUPPORTED, "graph: too many vectors");
  int maxlevel = 0;
  for (auto &r : L.R) maxlevel = std::max(maxlevel, r.level);
  const int npass = maxlevel + 1;
  if (npass > G_MAX_PASS) B2O_FAIL(B2O_EUNSUPPORTED, "graph: too many dependent reductions");
  GraphArgs &A = C.args;
  memset(&A, 0, sizeof(A));
  A.npass = npass;
  C.alg_arrays_read = 0;
  C.max_depth = C.max_slots = 0;
  for (int pass = 0; pass < npass; ++pass) {
    std::vector<uint32_t> code;
    std::vector<int> slots;
    int nlocal = 0;
    std::string err;
    for (size_t r = 0; r < L.R.size(); ++r) {
      if (L.R[r].level != pass + 1) continue;
      if (nlocal >= G_RED_PER_PASS) B2O_FAIL(B2O_EUNSUPPORTED, "graph: too many reductions in one pass");
      int sp = 0, maxsp = 0;
      if (!gen_code(L, L.R[r].expr, code, sp, maxsp, slots, err)) B2O_FAIL(B2O_EUNSUPPORTED, "graph: %s", err.c_str());
      C.max_depth = std::max(C.max_depth, maxsp);
      code.push_back((uint32_t)(F_RED + 1) | ((uint32_t)nlocal << 16));
      A.red_of_pass[pass][nlocal++] = (unsigned char)r;
    }
    A.nred_pass[pass] = nlocal;
    if (pass == npass - 1) {
      int sp = 0, maxsp = 0;
      if (!gen_code(L, out, code, sp, maxsp, slots, err)) B2O_FAIL(B2O_EUNSUPPORTED, "graph: %s", err.c_str());
      C.max_depth = std::max(C.max_depth, maxsp);
      code.push_back((uint32_t)(F_STORE + 1));
    }
    if ((int)code.size() > G_MAX_PROG) B2O_FAIL(B2O_EUNSUPPORTED, "graph: program too long (%zu)", code.size());
    A.prog_len[pass] = (int)code.size();
    for (size_t i = 0; i < code.size(); ++i) A.prog[pass][i] = code[i];
    A.narr_pass[pass] = (int)slots.size();
    C.max_slots = std::max(C.max_slots, (int)slots.size());
    for (size_t i = 0; i < slots.size(); ++i) A.arr_of_pass[pass][i] = (unsigned char)slots[i];
    C.alg_arrays_read += (int)slots.size();
    // scalars that become computable once the reductions of earlier passes are known
    int ns = 0;
    for (size_t s = 0; s < L.S.size(); ++s) {
      if (L.S[s].kind != SExpr::REDMUL) continue;
      if (L.R[L.S[s].b].level != pass) continue;   // reductions of level `pass` finished in pass-1
      if (ns >= G_MAX_SOP) B2O_FAIL(B2O_EUNSUPPORTED, "graph: too many reduction scalars");
      A.sop[pass][ns][0] = (unsigned char)s;
      A.sop[pass][ns][1] = (unsigned char)L.S[s].a;
      A.sop[pass][ns][2] = (unsigned char)L.S[s].b;
      ns++;
    }
    A.nsop[pass] = ns;
  }
  // REDMUL scalars whose own factor depends on a reduction are not supported (never produced by the descent above)
  for (auto &s : L.S)
    if (s.kind == SExpr::REDMUL && L.red_level_of_scalar(s.a) != 0)
      B2O_FAIL(B2O_EUNSUPPORTED, "graph: nested reduction scalars");
  for (size_t i = 2; i < L.arrays.size(); ++i) {
    A.arr[i] = L.arrays[i];
    A.arr_al16[i] = ((uintptr_t)L.arrays[i] % 16) == 0;
  }
  C.S = L.S;
  C.jit_src = jit_source(L, A, out);
  C.jit_state = 0;
  C.valid = true;
  return B2O_OK;
}

extern "C" int b2o_graph_create(b2o_ctx *ctx, int64_t n, b2o_graph **out) {
  // ctx may be NULL: a "dry" graph can be built, lowered and NVRTC-compiled on a CPU-only box, but not applied
  if (!out) B2O_FAIL(B2O_EARG, "null argument");
  if (n < 0) B2O_FAIL(B2O_EARG, "negative size");
  b2o_graph *g = new b2o_graph_s();
  g->ctx = ctx;
  g->n = n;
  *out = g;
  return B2O_OK;
}
extern "C" int b2o_graph_destroy(b2o_graph *g) {
  if (!g) return B2O_OK;
  for (int tr = 0; tr < 2; ++tr)
    for (int bz = 0; bz < 2; ++bz)
      if (g->prog[tr][bz].jit_module && g_jit.ModuleUnload) g_jit.ModuleUnload(g->prog[tr][bz].jit_module);
  delete g;
  return B2O_OK;
}

// generated CUDA source of one variant (for inspection / tests); returns the length, copies at most cap-1 bytes
extern "C" int b2o_graph_jit_source(b2o_graph *g, int transposed, double beta, char *buf, int64_t cap, int64_t *len) {
  if (!g || g->root < 0) B2O_FAIL(B2O_EARG, "graph not compiled");
  const Compiled &C = g->prog[transposed ? 1 : 0][beta != 0.0];
  if (len) *len = (int64_t)C.jit_src.size();
  if (buf && cap > 0) {
    size_t k = std::min<size_t>((size_t)cap - 1, C.jit_src.size());
    memcpy(buf, C.jit_src.data(), k);
    buf[k] = 0;
  }
  return B2O_OK;
}
// NVRTC-compile every variant for sm_100a (works without a GPU); B2O_EUNSUPPORTED when NVRTC is absent
extern "C" int b2o_graph_jit_check(b2o_graph *g, int64_t *cubin_bytes) {
  if (!g || g->root < 0) B2O_FAIL(B2O_EARG, "graph not compiled");
  int64_t total = 0;
  for (int tr = 0; tr < 2; ++tr)
    for (int bz = 0; bz < 2; ++bz) {
      std::vector<char> cubin;
      std::string log;
      if (jit_compile_cubin(g->prog[tr][bz].jit_src, cubin, log) != 0) {
        if (!g_jit.have_nvrtc) B2O_FAIL(B2O_EUNSUPPORTED, "NVRTC not available");
        B2O_FAIL(B2O_ECUDA, "NVRTC compilation failed: %s", log.c_str());
      }
      total += (int64_t)cubin.size();
    }
  if (cubin_bytes) *cubin_bytes = total;
  return B2O_OK;
}

// Ahead-of-time instantiations (csrc/b2o_graph_aot.cu, generated by tools/gen_graph_aot.py from this file's own code
// generator): the specialised pass kernels of the common static chains -- BASELINE config 3 and its transpose / beta != 0
// variants, H*D, D1*D2, D + c*I, H1*H2, H*D*H -- compiled by nvcc with the rest of the library, keyed by the FNV-1a hash of
// their generated source.  A tree whose lowering produces the same source runs the specialised kernel WITHOUT libnvrtc.
struct B2oAotKernel {
  unsigned long long hash;
  const void *fn;
};
extern const B2oAotKernel g_b2o_graph_aot[];
extern const int g_b2o_graph_aot_n;
static unsigned long long fnv1a64(const std::string &s) {
  unsigned long long h = 1469598103934665603ULL;
  for (unsigned char ch : s) {
    h ^= ch;
    h *= 1099511628211ULL;
  }
  return h;
}

// c->graph_jit: 0 interpreter only, 1 (default) ahead-of-time table, then NVRTC, then interpreter, 2 ahead-of-time table only
static int ensure_jit(b2o_ctx *c, Compiled &C) {
  if (C.jit_state == 0 || (C.jit_state == -2 && c->graph_jit == 1)) {
    C.jit_state = -1;
    const unsigned long long h = fnv1a64(C.jit_src);
    for (int i = 0; i < g_b2o_graph_aot_n; ++i)
      if (g_b2o_graph_aot[i].hash == h) {
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, g_b2o_graph_aot[i].fn, 256, 0) == cudaSuccess && nb >= 1) {
          C.aot_fn = g_b2o_graph_aot[i].fn;
          C.jit_blocks_per_sm = nb;
          C.jit_state = 2;
          return 2;
        }
        cudaGetLastError();
      }
    if (c->graph_jit == 2) {
      C.jit_state = -2;     // not in the table and NVRTC not allowed: retried if the option goes back to 1
      return -1;
    }
    jit_load_api();
    if (!g_jit.ok) return -1;
    std::vector<char> cubin;
    std::string log;
    if (jit_compile_cubin(C.jit_src, cubin, log) != 0) return -1;
    cudaFree(0);  // make sure the runtime's primary context is current for the driver calls
    void *mod = nullptr, *fn = nullptr;
    if (g_jit.ModuleLoadData(&mod, cubin.data()) != 0) return -1;
    if (g_jit.ModuleGetFunction(&fn, mod, "b2o_fused") != 0) return -1;
    int nb = 0;
    if (g_jit.OccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, 256, 0) != 0 || nb < 1) return -1;
    C.jit_module = mod;
    C.jit_func = fn;
    C.jit_blocks_per_sm = nb;
    C.jit_state = 1;
    return 1;
  }
  if (C.jit_state == 1 && c->graph_jit == 2) return -1;   // "ahead-of-time only" excludes an NVRTC module built earlier
  return C.jit_state;
}

// 1 when the last apply of this variant ran the NVRTC-specialised kernel, 0 for the interpreter
extern "C" int b2o_graph_uses_jit(b2o_graph *g, int transposed, double beta, int *out) {
  if (!g || g->root < 0 || !out) B2O_FAIL(B2O_EARG, "graph not compiled");
  *out = g->prog[transposed ? 1 : 0][beta != 0.0].jit_state >= 1;
  return B2O_OK;
}
// which executor the last apply of this variant used: 0 interpreter, 1 NVRTC-specialised, 2 ahead-of-time instantiation; and
// the FNV-1a hash of the variant's generated source (the key of the ahead-of-time table)
extern "C" int b2o_graph_variant(b2o_graph *g, int transposed, double beta, int *executor, uint64_t *source_hash) {
  if (!g || g->root < 0) B2O_FAIL(B2O_EARG, "graph not compiled");
  const Compiled &C = g->prog[transposed ? 1 : 0][beta != 0.0];
  if (executor) *executor = C.jit_state >= 1 ? C.jit_state : 0;
  if (source_hash) *source_hash = fnv1a64(C.jit_src);
  return B2O_OK;
}
extern "C" int b2o_graph_leaf(b2o_graph *g, int kind, const void *ptr, int *node) {
  if (!g || !node) B2O_FAIL(B2O_EARG, "null argument");
  if (kind < GK_DIAG || kind > GK_HOUSE) B2O_FAIL(B2O_EARG, "bad leaf kind %d", kind);
  if ((kind == GK_DIAG || kind == GK_HOUSE) && !ptr && g->n > 0) B2O_FAIL(B2O_EARG, "leaf needs a vector");
  if ((uintptr_t)ptr % 8) B2O_FAIL(B2O_EARG, "vectors must be 8-byte aligned");
  g->nodes.push_back(GNode{kind, -1, -1, 0.0, (const double *)ptr});
  *node = (int)g->nodes.size() - 1;
  return B2O_OK;
}
extern "C" int b2o_graph_unary(b2o_graph *g, int kind, int child, double x, int *node) {
  if (!g || !node) B2O_FAIL(B2O_EARG, "null argument");
  if (kind != GK_SCALE && kind != GK_NEG && kind != GK_TRANS) B2O_FAIL(B2O_EARG, "bad unary kind %d", kind);
  if (child < 0 || child >= (int)g->nodes.size()) B2O_FAIL(B2O_EARG, "bad child");
  g->nodes.push_back(GNode{kind, child, -1, x, nullptr});
  *node = (int)g->nodes.size() - 1;
  return B2O_OK;
}
extern "C" int b2o_graph_binary(b2o_graph *g, int kind, int a, int b, int *node) {
  if (!g || !node) B2O_FAIL(B2O_EARG, "null argument");
  if (kind != GK_SUM && kind != GK_PROD) B2O_FAIL(B2O_EARG, "bad binary kind %d", kind);
  if (a < 0 || b < 0 || a >= (int)g->nodes.size() || b >= (int)g->nodes.size()) B2O_FAIL(B2O_EARG, "bad child");
  g->nodes.push_back(GNode{kind, a, b, 0.0, nullptr});
  *node = (int)g->nodes.size() - 1;
  return B2O_OK;
}
extern "C" int b2o_graph_compile(b2o_graph *g, int root) {
  if (!g) B2O_FAIL(B2O_EARG, "null graph");
  if (root < 0 || root >= (int)g->nodes.size()) B2O_FAIL(B2O_EARG, "bad root");
  g->root = root;
  for (int tr = 0; tr < 2; ++tr)
    for (int bz = 0; bz < 2; ++bz) B2O_TRY(compile_variant(g, tr != 0, bz != 0, g->prog[tr][bz]));
  return B2O_OK;
}
extern "C" int b2o_graph_info(b2o_graph *g, int transposed, double beta, int *npasses, int *nreductions, double *alg_bytes) {
  if (!g || g->root < 0) B2O_FAIL(B2O_EARG, "graph not compiled");
  const Compiled &C = g->prog[transposed ? 1 : 0][beta != 0.0];
  int nred = 0;
  for (int p = 0; p < C.args.npass; ++p) nred += C.args.nred_pass[p];
  if (npasses) *npasses = C.args.npass;
  if (nreductions) *nreductions = nred;
  if (alg_bytes) *alg_bytes = 8.0 * (double)g->n * (C.alg_arrays_read + 1);
  return B2O_OK;
}

static double eval_scalar(const std::vector<SExpr> &S, int s, double alpha, double beta) {
  const SExpr &e = S[s];
  switch (e.kind) {
    case SExpr::CONST: return e.c;
    case SExpr::ALPHA: return alpha;
    case SExpr::BETA: return beta;
    case SExpr::MUL: return eval_scalar(S, e.a, alpha, beta) * eval_scalar(S, e.b, alpha, beta);
    case SExpr::NEG: return -eval_scalar(S, e.a, alpha, beta);
    case SExpr::REDMUL: return 0.0;  // filled on the device
  }
  return 0.0;
}

extern "C" int b2o_graph_apply(b2o_graph *g, int transposed, void *res, int64_t res_len, const void *v, int64_t v_len,
                               double alpha, double beta) {
  if (!g || g->root < 0) B2O_FAIL(B2O_EARG, "graph not compiled");
  if (res_len != g->n || v_len != g->n) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if (g->n == 0) return B2O_OK;
  if (!res || !v) B2O_FAIL(B2O_EARG, "null vector");
  if (((uintptr_t)res | (uintptr_t)v) % 8) B2O_FAIL(B2O_EARG, "vectors must be 8-byte aligned");
  b2o_ctx *c = g->ctx;
  if (!c) B2O_FAIL(B2O_EARG, "graph was created without a context (dry graph)");
  B2O_CUDA(cudaSetDevice(c->device));
  Compiled &C = g->prog[transposed ? 1 : 0][beta != 0.0];
  const int jit_kind = c->graph_jit ? ensure_jit(c, C) : -1;
  if (jit_kind >= 1) {
    JitRT R;
    memset(&R, 0, sizeof(R));
    for (int i = 2; i < G_MAX_ARR; ++i) R.arr[i] = C.args.arr[i];
    R.arr[0] = (const double *)v;
    R.arr[1] = (const double *)res;
    R.out = (double *)res;
    R.n = g->n;
    for (size_t sidx = 0; sidx < C.S.size(); ++sidx) R.scal[sidx] = eval_scalar(C.S, (int)sidx, alpha, beta);
    R.partials = c->d_partials;
    R.dots = c->d_dots + 400;
    R.bar = c->d_bar;
    R.arrive = c->d_bar + 1;
    bool vec = (((uintptr_t)v | (uintptr_t)res) % 16) == 0;
    for (int i = 2; i < G_MAX_ARR; ++i) vec = vec && (((uintptr_t)C.args.arr[i]) % 16 == 0);
    R.vec = vec ? 1 : 0;
    const int64_t units = (vec ? (g->n >> 1) : g->n);
    const int64_t want = (units + 2 * 256 - 1) / (2 * 256);
    int grid = (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)c->num_sms * C.jit_blocks_per_sm));
    grid = std::min(grid, B2O_MAX_GRID);
    int nbar = 0;
    for (int p = 0; p < C.args.npass; ++p) nbar += C.args.nred_pass[p] > 0;
    void *params[] = {(void *)&R};
    if (c->time_kernels) B2O_CUDA(cudaEventRecord(c->ev0, c->stream));
    if (c->nranks <= 1) {
      R.pass_begin = 0;
      R.pass_end = C.args.npass;
      R.fused = 1;
      R.bar_target = c->bar_base + (unsigned long long)grid;
      if (jit_kind == 2) {
        B2O_CUDA(cudaLaunchCooperativeKernel(C.aot_fn, dim3(grid), dim3(256), params, 0, c->stream));
      } else {
        int rc = g_jit.LaunchCooperativeKernel(C.jit_func, grid, 1, 1, 256, 1, 1, 0, (void *)c->stream, params);
        if (rc != 0) B2O_FAIL(B2O_ECUDA, "cuLaunchCooperativeKernel failed (%d)", rc);
      }
      c->bar_base += (unsigned long long)grid * nbar;
      c->launches++;
    } else {
      for (int p = 0; p < C.args.npass; ++p) {
        R.pass_begin = p;
        R.pass_end = p + 1;
        R.fused = 0;
        if (jit_kind == 2) {
          B2O_CUDA(cudaLaunchKernel(C.aot_fn, dim3(grid), dim3(256), params, 0, c->stream));
        } else {
          int rc = g_jit.LaunchKernel(C.jit_func, grid, 1, 1, 256, 1, 1, 0, (void *)c->stream, params, nullptr);
          if (rc != 0) B2O_FAIL(B2O_ECUDA, "cuLaunchKernel failed (%d)", rc);
        }
        c->launches++;
        for (int r = 0; r < C.args.nred_pass[p]; ++r) B2O_TRY(b2o_allreduce_sum_f64(c, R.dots + C.args.red_of_pass[p][r], 1));
      }
    }
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
  GraphArgs A = C.args;
  A.arr[0] = (const double *)v;
  A.arr_al16[0] = ((uintptr_t)v % 16) == 0;
  A.arr[1] = (const double *)res;
  A.arr_al16[1] = ((uintptr_t)res % 16) == 0;
  A.out = (double *)res;
  A.out_al16 = A.arr_al16[1];
  A.n = g->n;
  for (size_t s = 0; s < C.S.size(); ++s) A.scal[s] = eval_scalar(C.S, (int)s, alpha, beta);
  A.partials = c->d_partials;
  A.dots = c->d_dots + 400;
  A.bar = c->d_bar;
  A.arrive = c->d_bar + 1;
  // small programs (cfg3 and its variants: depth <= 3, <= 4 vectors per pass) run on the 4-rows-per-dispatch machine
  const bool small = C.max_depth <= 3 && C.max_slots <= 4 && c->graph_interp != 2;
  const int ept = small ? 4 : 2;
  const void *kern = small ? (c->graph_blocks >= 2 ? (const void *)graph_kernel<2, 3, 4, 4>      // 117 registers, no spills; 3 CTAs/SM would spill
                                                   : (const void *)graph_kernel<1, 3, 4, 4>)
                           : (c->graph_blocks >= 4   ? (const void *)graph_kernel<4, G_DEPTH, G_SLOTS, 2>
                              : c->graph_blocks == 3 ? (const void *)graph_kernel<3, G_DEPTH, G_SLOTS, 2>
                              : c->graph_blocks == 2 ? (const void *)graph_kernel<2, G_DEPTH, G_SLOTS, 2>
                                                     : (const void *)graph_kernel<1, G_DEPTH, G_SLOTS, 2>);
  int blocks_per_sm = 0;
  B2O_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, G_NT, 0));
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  if (c->graph_blocks > 0) blocks_per_sm = std::min(blocks_per_sm, small ? std::min(c->graph_blocks, 2) : c->graph_blocks);
  const int64_t ntiles = (g->n + (int64_t)G_NT * ept - 1) / ((int64_t)G_NT * ept);
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t)c->num_sms * blocks_per_sm));
  grid = std::min(grid, B2O_MAX_GRID);
  void *kargs[] = {(void *)&A};
  int nbar = 0;
  for (int p = 0; p < A.npass; ++p) nbar += A.nred_pass[p] > 0;
  if (c->nranks <= 1) {
    A.pass_begin = 0;
    A.pass_end = A.npass;
    A.fused = 1;
    A.bar_target = c->bar_base + (unsigned long long)grid;
    if (c->time_kernels) B2O_CUDA(cudaEventRecord(c->ev0, c->stream));
    B2O_CUDA(cudaLaunchCooperativeKernel(kern, dim3(grid), dim3(G_NT), kargs, 0, c->stream));
    c->bar_base += (unsigned long long)grid * nbar;
    c->launches++;
    if (c->time_kernels) {
      B2O_CUDA(cudaEventRecord(c->ev1, c->stream));
      B2O_CUDA(cudaEventSynchronize(c->ev1));
      float ms = 0.f;
      B2O_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
      c->kern_ms += ms;
      c->kern_n++;
    }
  } else {
    // row-partitioned: one launch per pass, reductions all-reduced in between
    for (int p = 0; p < A.npass; ++p) {
      A.pass_begin = p;
      A.pass_end = p + 1;
      A.fused = 0;
      B2O_CUDA(cudaLaunchKernel(kern, dim3(grid), dim3(G_NT), kargs, 0, c->stream));
      c->launches++;
      for (int r = 0; r < A.nred_pass[p]; ++r) B2O_TRY(b2o_allreduce_sum_f64(c, A.dots + A.red_of_pass[p][r], 1));
    }
  }
  return B2O_OK;
}
