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
constexpr int G_DEPTH = 6, G_SLOTS = 6, G_EPT = 4, G_NT = 256, G_RED_PER_PASS = 4, G_MAX_SOP = 8;
enum { I_PUSH_ARR = 0, I_PUSH_SCAL = 1, I_ADD = 2, I_SUB = 3, I_MUL = 4, I_RED = 5, I_STORE = 6 };
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
  uint32_t prog[G_MAX_PASS][G_MAX_PROG];        // op | sp<<8 | arg<<16
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

template <int MINB>
__global__ void __launch_bounds__(G_NT, MINB) graph_kernel(const __grid_constant__ GraphArgs p) {
  __shared__ double s_scal[G_MAX_SCAL];
  __shared__ double s_red[G_MAX_RED];
  __shared__ double s_w[G_NT / 32][G_RED_PER_PASS];
  __shared__ uint32_t s_prog[G_MAX_PROG];
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
    for (int i = tid; i < len; i += G_NT) s_prog[i] = p.prog[pass][i];
    __syncthreads();
    const int narr = p.narr_pass[pass];
    double racc[G_RED_PER_PASS];
#pragma unroll
    for (int r = 0; r < G_RED_PER_PASS; ++r) racc[r] = 0.0;
    const bool reverse = pass & 1;  // alternate direction: the tail of the previous pass is still in L2

    for (int64_t tt = blockIdx.x; tt < ntiles; tt += gridDim.x) {
      const int64_t t = reverse ? (ntiles - 1 - tt) : tt;
      // rows of this thread: pairs e = jj*G_NT + tid  (jj < G_EPT/2) -> rows 2e, 2e+1 of the tile
      const int64_t row0 = t * tile_rows;
      double av[G_SLOTS][G_EPT];
#pragma unroll
      for (int s = 0; s < G_SLOTS; ++s) {
        if (s < narr) {
          const int a = p.arr_of_pass[pass][s];
          const double *base = p.arr[a];
          const bool al = p.arr_al16[a];
#pragma unroll
          for (int jj = 0; jj < G_EPT / 2; ++jj) {
            const int64_t r = row0 + 2 * ((int64_t)jj * G_NT + tid);
            if (al && r + 1 < p.n) {
              double2 v = *reinterpret_cast<const double2 *>(base + r);
              av[s][2 * jj] = v.x;
              av[s][2 * jj + 1] = v.y;
            } else {
              av[s][2 * jj] = r < p.n ? base[r] : 0.0;
              av[s][2 * jj + 1] = r + 1 < p.n ? base[r + 1] : 0.0;
            }
          }
        } else {
          G_FOR_J av[s][j] = 0.0;
        }
      }
      double st[G_DEPTH][G_EPT];
#pragma unroll
      for (int d = 0; d < G_DEPTH; ++d) G_FOR_J st[d][j] = 0.0;

      for (int pc = 0; pc < len; ++pc) {
        const uint32_t ins = s_prog[pc];
        const int op = ins & 0xff, sp = (ins >> 8) & 0xff, arg = ins >> 16;
        switch (op) {
          case I_PUSH_ARR: {
            double tmp[G_EPT];
            switch (arg) {
#define G_CASE_SLOT(S) case S: G_FOR_J tmp[j] = av[S][j]; break;
              G_CASE_SLOT(0) G_CASE_SLOT(1) G_CASE_SLOT(2) G_CASE_SLOT(3) G_CASE_SLOT(4) G_CASE_SLOT(5)
#undef G_CASE_SLOT
              default: G_FOR_J tmp[j] = 0.0;
            }
            switch (sp) {
#define G_CASE_SP(S) case S: G_FOR_J st[S][j] = tmp[j]; break;
              G_CASE_SP(0) G_CASE_SP(1) G_CASE_SP(2) G_CASE_SP(3) G_CASE_SP(4) G_CASE_SP(5)
#undef G_CASE_SP
            }
            break;
          }
          case I_PUSH_SCAL: {
            const double c = s_scal[arg];
            switch (sp) {
#define G_CASE_SP(S) case S: G_FOR_J st[S][j] = c; break;
              G_CASE_SP(0) G_CASE_SP(1) G_CASE_SP(2) G_CASE_SP(3) G_CASE_SP(4) G_CASE_SP(5)
#undef G_CASE_SP
            }
            break;
          }
          case I_ADD:
            switch (sp) {
#define G_CASE_SP(S) case S: G_FOR_J st[S - 2][j] = st[S - 2][j] + st[S - 1][j]; break;
              G_CASE_SP(2) G_CASE_SP(3) G_CASE_SP(4) G_CASE_SP(5) G_CASE_SP(6)
#undef G_CASE_SP
            }
            break;
          case I_SUB:
            switch (sp) {
#define G_CASE_SP(S) case S: G_FOR_J st[S - 2][j] = st[S - 2][j] - st[S - 1][j]; break;
              G_CASE_SP(2) G_CASE_SP(3) G_CASE_SP(4) G_CASE_SP(5) G_CASE_SP(6)
#undef G_CASE_SP
            }
            break;
          case I_MUL:
            switch (sp) {
#define G_CASE_SP(S) case S: G_FOR_J st[S - 2][j] = st[S - 2][j] * st[S - 1][j]; break;
              G_CASE_SP(2) G_CASE_SP(3) G_CASE_SP(4) G_CASE_SP(5) G_CASE_SP(6)
#undef G_CASE_SP
            }
            break;
          case I_RED: {
            // masked sum of the top of stack into local reduction `arg` (rows >= n contribute nothing)
            double tmp[G_EPT];
            switch (sp) {
#define G_CASE_SP(S) case S: G_FOR_J tmp[j] = st[S - 1][j]; break;
              G_CASE_SP(1) G_CASE_SP(2) G_CASE_SP(3) G_CASE_SP(4) G_CASE_SP(5) G_CASE_SP(6)
#undef G_CASE_SP
              default: G_FOR_J tmp[j] = 0.0;
            }
            double s = 0.0;
#pragma unroll
            for (int jj = 0; jj < G_EPT / 2; ++jj) {
              const int64_t r = row0 + 2 * ((int64_t)jj * G_NT + tid);
              if (r < p.n) s += tmp[2 * jj];
              if (r + 1 < p.n) s += tmp[2 * jj + 1];
            }
            switch (arg) {
              case 0: racc[0] += s; break;
              case 1: racc[1] += s; break;
              case 2: racc[2] += s; break;
              case 3: racc[3] += s; break;
            }
            break;
          }
          case I_STORE: {
            double tmp[G_EPT];
            switch (sp) {
#define G_CASE_SP(S) case S: G_FOR_J tmp[j] = st[S - 1][j]; break;
              G_CASE_SP(1) G_CASE_SP(2) G_CASE_SP(3) G_CASE_SP(4) G_CASE_SP(5) G_CASE_SP(6)
#undef G_CASE_SP
              default: G_FOR_J tmp[j] = 0.0;
            }
#pragma unroll
            for (int jj = 0; jj < G_EPT / 2; ++jj) {
              const int64_t r = row0 + 2 * ((int64_t)jj * G_NT + tid);
              if (p.out_al16 && r + 1 < p.n) {
                stg_stream2(p.out + r, make_double2(tmp[2 * jj], tmp[2 * jj + 1]));
              } else {
                if (r < p.n) p.out[r] = tmp[2 * jj];
                if (r + 1 < p.n) p.out[r + 1] = tmp[2 * jj + 1];
              }
            }
            break;
          }
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

static bool gen_code(Lowering &L, int v, std::vector<uint32_t> &code, int &sp, int &maxsp, std::vector<int> &slots,
                     std::string &err) {
  const VExpr &e = L.V[v];
  if (e.kind == VExpr::BIN) {
    if (!gen_code(L, e.l, code, sp, maxsp, slots, err)) return false;
    if (!gen_code(L, e.r, code, sp, maxsp, slots, err)) return false;
    code.push_back((uint32_t)e.op | ((uint32_t)sp << 8));
    sp -= 1;
    return true;
  }
  if (sp >= G_DEPTH) { err = "expression too deep for the fused evaluator"; return false; }
  if (e.kind == VExpr::ARR) {
    int slot = -1;
    for (size_t i = 0; i < slots.size(); ++i)
      if (slots[i] == e.id) slot = (int)i;
    if (slot < 0) {
      if ((int)slots.size() >= G_SLOTS) { err = "too many distinct vectors in one pass"; return false; }
      slots.push_back(e.id);
      slot = (int)slots.size() - 1;
    }
    code.push_back((uint32_t)I_PUSH_ARR | ((uint32_t)sp << 8) | ((uint32_t)slot << 16));
  } else {
    code.push_back((uint32_t)I_PUSH_SCAL | ((uint32_t)sp << 8) | ((uint32_t)e.id << 16));
  }
  sp += 1;
  maxsp = std::max(maxsp, sp);
  return true;
}

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
      code.push_back((uint32_t)I_RED | (1u << 8) | ((uint32_t)nlocal << 16));
      A.red_of_pass[pass][nlocal++] = (unsigned char)r;
    }
    A.nred_pass[pass] = nlocal;
    if (pass == npass - 1) {
      int sp = 0, maxsp = 0;
      if (!gen_code(L, out, code, sp, maxsp, slots, err)) B2O_FAIL(B2O_EUNSUPPORTED, "graph: %s", err.c_str());
      code.push_back((uint32_t)I_STORE | (1u << 8));
    }
    if ((int)code.size() > G_MAX_PROG) B2O_FAIL(B2O_EUNSUPPORTED, "graph: program too long (%zu)", code.size());
    A.prog_len[pass] = (int)code.size();
    for (size_t i = 0; i < code.size(); ++i) A.prog[pass][i] = code[i];
    A.narr_pass[pass] = (int)slots.size();
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
  C.valid = true;
  return B2O_OK;
}

extern "C" int b2o_graph_create(b2o_ctx *ctx, int64_t n, b2o_graph **out) {
  if (!ctx || !out) B2O_FAIL(B2O_EARG, "null argument");
  if (n < 0) B2O_FAIL(B2O_EARG, "negative size");
  b2o_graph *g = new b2o_graph_s();
  g->ctx = ctx;
  g->n = n;
  *out = g;
  return B2O_OK;
}
extern "C" int b2o_graph_destroy(b2o_graph *g) {
  delete g;
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
  B2O_CUDA(cudaSetDevice(c->device));
  Compiled &C = g->prog[transposed ? 1 : 0][beta != 0.0];
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
  const bool two = c->graph_blocks == 2;
  const void *kern = two ? (const void *)graph_kernel<2> : (const void *)graph_kernel<1>;
  int blocks_per_sm = 0;
  B2O_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kern, G_NT, 0));
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  if (c->graph_blocks > 0) blocks_per_sm = std::min(blocks_per_sm, c->graph_blocks);
  const int64_t ntiles = (g->n + (int64_t)G_NT * G_EPT - 1) / ((int64_t)G_NT * G_EPT);
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
