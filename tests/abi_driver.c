/*
 * abi_driver.c -- a plain C99 host for libb2o.so: no Python, no torch, no ctypes.
 *
 * This is the shape of caller the reference-side binding is (a Julia `ccall` is a C call): every argument below goes
 * through the prototypes of include/b2o.h as compiled by gcc, so a by-value / by-pointer slip, a wrong integer width or
 * a missing export fails HERE at compile / link / run time instead of hiding behind a dynamically typed binding.
 * Flow: context -> leaf operators -> index operators -> L-BFGS / inverse L-BFGS / L-SR1 push! + apply (device and host
 * buffers) -> diagonal quasi-Newton push! -> fused static tree (BASELINE config 3) -> kron on the tensor cores ->
 * error codes.  Every result is compared with the CPU oracle (oracle/libb2o_oracle.so, the checker -- test
 * infrastructure, linked only into this test binary).
 *
 * Build + run: tests/test_abi_driver.py (gcc -std=c99 -Wall -Wextra -Werror).  Prints "ABI_DRIVER_OK <checks>".
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/b2o.h"
#include "../oracle/b2o_oracle.h"

static int g_checks = 0;
static b2o_ctx *ctx = NULL;

#define OK(expr)                                                                                  \
  do {                                                                                            \
    int st_ = (expr);                                                                             \
    if (st_ != B2O_OK) {                                                                          \
      fprintf(stderr, "%s:%d: %s -> status %d: %s\n", __FILE__, __LINE__, #expr, st_, b2o_last_error()); \
      exit(1);                                                                                    \
    }                                                                                             \
  } while (0)
#define CHECK(cond, ...)                                    \
  do {                                                      \
    if (!(cond)) {                                          \
      fprintf(stderr, "%s:%d: CHECK(%s) failed: ", __FILE__, __LINE__, #cond); \
      fprintf(stderr, __VA_ARGS__);                         \
      fprintf(stderr, "\n");                                \
      exit(1);                                              \
    }                                                       \
    ++g_checks;                                             \
  } while (0)

static double *dev_f64(int64_t n) {
  void *p = NULL;
  OK(b2o_malloc(ctx, (size_t)(n > 0 ? n : 1) * sizeof(double), &p));
  return (double *)p;
}
static double *host_f64(int64_t n) {
  double *p = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
  if (!p) exit(2);
  return p;
}
static void h2d(void *d, const void *h, size_t bytes) { OK(b2o_memcpy_h2d(ctx, d, h, bytes)); }
static void d2h(void *h, const void *d, size_t bytes) { OK(b2o_memcpy_d2h(ctx, h, d, bytes)); }
static double rel_err(const double *a, const double *b, int64_t n) {
  long double num = 0, den = 0;
  for (int64_t i = 0; i < n; ++i) {
    num += ((long double)a[i] - b[i]) * ((long double)a[i] - b[i]);
    den += (long double)b[i] * b[i];
  }
  return (double)sqrtl(num / (den > 0 ? den : 1));
}
static int same_bits(const double *a, const double *b, int64_t n) { return memcmp(a, b, (size_t)n * sizeof(double)) == 0; }

/* ---- leaf operators: bit-exact where no reduction is involved ------------------------------------------------ */
static void test_leaves(void) {
  const int64_t n = 100003;
  double *d = dev_f64(n), *v = dev_f64(n), *res = dev_f64(n);
  double *hd = host_f64(n), *hv = host_f64(n), *hr = host_f64(n), *ref = host_f64(n);
  OK(b2o_fill_uniform(ctx, B2O_F64, d, n, 1, 0.0, 1.0));
  OK(b2o_fill_uniform(ctx, B2O_F64, v, n, 2, 0.0, 1.0));
  OK(b2o_fill_uniform(ctx, B2O_F64, res, n, 9, 0.0, 1.0));
  orc_fill_uniform(hd, n, 1, 0.0, 1.0);
  orc_fill_uniform(hv, n, 2, 0.0, 1.0);
  orc_fill_uniform(ref, n, 9, 0.0, 1.0);
  d2h(hr, d, (size_t)n * 8);
  CHECK(same_bits(hr, hd, n), "device and oracle generators differ");
  /* mulSquareOpDiagonal!, 5-arg form (test/test_linop.jl:308-319 uses alpha = beta = 2) */
  OK(b2o_diag_apply(ctx, B2O_F64, n, n, d, n, res, n, v, n, 2.0, 2.0));
  orc_diag_square(ref, hd, hv, n, 2.0, 2.0);
  d2h(hr, res, (size_t)n * 8);
  CHECK(same_bits(hr, ref, n), "opDiagonal 5-arg");
  /* mulOpEye!, rectangular: tail is the scalar beta (quirk Q2) */
  OK(b2o_eye_apply(ctx, B2O_F64, n, n - 7, res, n, v, n - 7, 1.5, -0.5));
  orc_eye(ref, n, hv, n - 7, 1.5, -0.5, n - 7);
  d2h(hr, res, (size_t)n * 8);
  CHECK(same_bits(hr, ref, n), "opEye rectangular");
  /* mulOpZeros! */
  OK(b2o_zeros_apply(ctx, B2O_F64, n, n, res, n, n, 1.0, 3.0));
  orc_zeros(ref, n, 1.0, 3.0);
  d2h(hr, res, (size_t)n * 8);
  CHECK(same_bits(hr, ref, n), "opZeros");
  /* mulOpOnes!, mulHouseholder!: reductions, 1e-12 */
  OK(b2o_ones_apply(ctx, B2O_F64, n, n, res, n, v, n, 0.5, 0.0));
  orc_ones(ref, n, hv, n, 0.5, 0.0);
  d2h(hr, res, (size_t)n * 8);
  CHECK(rel_err(hr, ref, n) <= 1e-12, "opOnes %g", rel_err(hr, ref, n));
  double nrm = orc_nrm2(hd, n);
  for (int64_t i = 0; i < n; ++i) hd[i] /= nrm;
  h2d(d, hd, (size_t)n * 8);
  OK(b2o_householder_apply(ctx, B2O_F64, n, d, res, n, v, n, 1.0, 0.0));
  orc_householder(ref, hd, hv, n, 1.0, 0.0);
  d2h(hr, res, (size_t)n * 8);
  CHECK(rel_err(hr, ref, n) <= 1e-12, "opHouseholder %g", rel_err(hr, ref, n));
  /* shape mismatch is reported before any work, with the reference's message (src/operations.jl:23-24) */
  CHECK(b2o_diag_apply(ctx, B2O_F64, n, n, d, n, res, n - 1, v, n, 1.0, 0.0) == B2O_ESHAPE, "shape check");
  CHECK(strstr(b2o_last_error(), "shape mismatch") != NULL, "message: %s", b2o_last_error());
  OK(b2o_free(ctx, d)); OK(b2o_free(ctx, v)); OK(b2o_free(ctx, res));
  free(hd); free(hv); free(hr); free(ref);
}

/* ---- opRestriction / opExtension: index work, exact (test/test_linop.jl:437-467), duplicates: last wins ------- */
static void test_index(void) {
  const int64_t ncol = 50000, k = 20000;
  int64_t *idx = (int64_t *)malloc((size_t)k * sizeof(int64_t));
  unsigned long long s = 12345;
  for (int64_t i = 0; i < k; ++i) {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    idx[i] = (int64_t)((s >> 33) % (unsigned long long)ncol) + 1;       /* 1-based, with duplicates */
  }
  b2o_index *ix = NULL;
  OK(b2o_index_create(ctx, idx, k, ncol, &ix));
  double *v = dev_f64(ncol), *u = dev_f64(k), *rk = dev_f64(k), *rn = dev_f64(ncol);
  double *hv = host_f64(ncol), *hu = host_f64(k), *got = host_f64(ncol), *ref = host_f64(ncol);
  orc_fill_uniform(hv, ncol, 3, -1.0, 1.0);
  orc_fill_uniform(hu, k, 4, -1.0, 1.0);
  h2d(v, hv, (size_t)ncol * 8);
  h2d(u, hu, (size_t)k * 8);
  OK(b2o_restrict_apply(ix, B2O_F64, rk, k, v, ncol));
  orc_restrict(ref, idx, k, hv);
  d2h(got, rk, (size_t)k * 8);
  CHECK(same_bits(got, ref, k), "opRestriction");
  OK(b2o_extend_apply(ix, B2O_F64, rn, ncol, u, k));
  orc_extend(ref, ncol, idx, k, hu);
  d2h(got, rn, (size_t)ncol * 8);
  CHECK(same_bits(got, ref, ncol), "opExtension");
  OK(b2o_index_destroy(ix));
  idx[0] = ncol + 1;
  CHECK(b2o_index_create(ctx, idx, k, ncol, &ix) == B2O_EARG, "index range check");
  CHECK(strstr(b2o_last_error(), "indices should be between 1 and") != NULL, "message: %s", b2o_last_error());
  OK(b2o_free(ctx, v)); OK(b2o_free(ctx, u)); OK(b2o_free(ctx, rk)); OK(b2o_free(ctx, rn));
  free(idx); free(hv); free(hu); free(got); free(ref);
}

/* ---- quasi-Newton operators: push! + apply against the oracle's push! + apply --------------------------------- */
static void test_qn(void) {
  const int64_t n = 50021;
  const int mem = 5, npush = 7;
  double *s = dev_f64(n), *y = dev_f64(n), *x = dev_f64(n), *res = dev_f64(n);
  double *hs = host_f64(n), *hy = host_f64(n), *hu = host_f64(n), *hx = host_f64(n), *got = host_f64(n), *ref = host_f64(n);
  orc_set_mode(1, 1);
  orc_fill_uniform(hx, n, 7, 0.0, 1.0);
  h2d(x, hx, (size_t)n * 8);
  for (int kind = 0; kind < 3; ++kind) {          /* 0 forward L-BFGS, 1 inverse L-BFGS, 2 L-SR1 */
    b2o_qn *op = NULL;
    orc_lbfgs *ol = NULL;
    orc_lsr1 *os = NULL;
    if (kind < 2) {
      OK(b2o_lbfgs_create(ctx, B2O_F64, n, mem, 1, 0, 0.99, 10.0, kind == 1, &op));
      ol = orc_lbfgs_create(n, mem, 1, 0, 0.99, 10.0, kind == 1);
    } else {
      OK(b2o_lsr1_create(ctx, B2O_F64, n, mem, 1, &op));
      os = orc_lsr1_create(n, mem, 1);
    }
    for (int i = 0; i < npush; ++i) {
      orc_fill_uniform(hs, n, 100 + (uint64_t)i, 0.0, 1.0);
      orc_fill_uniform(hu, n, 200 + (uint64_t)i, 0.0, 1.0);
      for (int64_t j = 0; j < n; ++j) hy[j] = kind < 2 ? hs[j] + 0.1 * hu[j] : 2.0 * hs[j] + 0.3 * hu[j];
      h2d(s, hs, (size_t)n * 8);
      h2d(y, hy, (size_t)n * 8);
      int acc = -1;
      OK(b2o_qn_push(op, s, y, n, &acc));
      const int oacc = kind < 2 ? orc_lbfgs_push(ol, hs, hy) : orc_lsr1_push(os, hs, hy);
      CHECK(acc == oacc, "push! acceptance kind %d step %d: %d vs %d", kind, i, acc, oacc);
    }
    int ins = 0;
    double gamma = 0, ub = 0, ys[8], aux[8];
    OK(b2o_qn_get_scalars(op, &ins, &gamma, &ub, ys, aux));
    const double og = kind < 2 ? orc_lbfgs_gamma(ol) : orc_lsr1_gamma(os);
    CHECK(ins == (kind < 2 ? orc_lbfgs_insert(ol) : orc_lsr1_insert(os)), "ring index");
    CHECK(fabs(gamma - og) <= 1e-13 * fabs(og), "scaling factor %g vs %g", gamma, og);
    /* mul!(res, op, x, 1.5, -0.25) */
    orc_fill_uniform(ref, n, 8, 0.0, 1.0);
    h2d(res, ref, (size_t)n * 8);
    OK(b2o_qn_apply(op, res, n, x, n, 1.5, -0.25));
    if (kind < 2) orc_lbfgs_apply(ol, ref, hx, 1.5, -0.25);
    else orc_lsr1_apply(os, ref, hx, 1.5, -0.25);
    d2h(got, res, (size_t)n * 8);
    CHECK(rel_err(got, ref, n) <= 1e-12, "apply kind %d: %g", kind, rel_err(got, ref, n));
    /* host-buffer entry (H2D + apply + D2H inside) */
    OK(b2o_qn_apply_host(op, got, hx, n, 1.0, 0.0));
    if (kind < 2) orc_lbfgs_apply(ol, ref, hx, 1.0, 0.0);
    else orc_lsr1_apply(os, ref, hx, 1.0, 0.0);
    CHECK(rel_err(got, ref, n) <= 1e-12, "apply_host kind %d: %g", kind, rel_err(got, ref, n));
    if (kind != 1) {                               /* diag! (forward operators only) */
      OK(b2o_qn_diag(op, res, n));
      if (kind == 0) orc_lbfgs_diag(ol, ref);
      else orc_lsr1_diag(os, ref);
      d2h(got, res, (size_t)n * 8);
      CHECK(rel_err(got, ref, n) <= 1e-12, "diag! kind %d: %g", kind, rel_err(got, ref, n));
    } else {
      CHECK(b2o_qn_diag(op, res, n) == B2O_ESTATE, "diag! of an inverse operator must be refused");
    }
    CHECK(b2o_qn_apply(op, res, n, x, n - 1, 1.0, 0.0) == B2O_ESHAPE, "shape check");
    double bytes = 0;
    OK(b2o_qn_apply_bytes(op, 0.0, &bytes));
    CHECK(bytes == (kind == 0 ? 4.0 * mem + 3 : kind == 1 ? 8.0 * mem + 2 : 2.0 * mem + 3) * 8.0 * (double)n, "algorithmic bytes %g", bytes);
    OK(b2o_qn_reset(op));
    OK(b2o_qn_apply(op, res, n, x, n, 1.0, 0.0));   /* identity after reset! (test/test_lbfgs.jl:13-20) */
    d2h(got, res, (size_t)n * 8);
    CHECK(same_bits(got, hx, n), "reset! -> identity");
    OK(b2o_qn_destroy(op));
    if (ol) orc_lbfgs_destroy(ol);
    if (os) orc_lsr1_destroy(os);
  }
  /* diagonal quasi-Newton push! (DiagonalPSB), pinned on the oracle */
  orc_fill_uniform(hs, n, 31, -1.0, 1.0);
  orc_fill_uniform(hy, n, 32, -1.0, 1.0);
  for (int64_t j = 0; j < n; ++j) ref[j] = 1.0;
  h2d(s, hs, (size_t)n * 8);
  h2d(y, hy, (size_t)n * 8);
  h2d(res, ref, (size_t)n * 8);
  OK(b2o_diagqn_push(ctx, 0, res, n, s, y, n));
  CHECK(orc_diagqn_push(0, ref, hs, hy, n) == 0, "oracle diagqn");
  d2h(got, res, (size_t)n * 8);
  CHECK(rel_err(got, ref, n) <= 1e-12, "DiagonalPSB push!: %g", rel_err(got, ref, n));
  OK(b2o_free(ctx, s)); OK(b2o_free(ctx, y)); OK(b2o_free(ctx, x)); OK(b2o_free(ctx, res));
  free(hs); free(hy); free(hu); free(hx); free(got); free(ref);
}

/* ---- LBFGSOperator(Float32, n) / InverseLBFGSOperator / LSR1Operator(Float32, n): test/test_lbfgs.jl:162-178, test/test_lsr1.jl:74-86.
 * The reference's precision test pushes s = y = ones and applies to v = (-1)^i.  With s = y the updates leave the identity:
 * forward: gamma = ys/yy = 1 and a_k == b_k bit for bit, so q + (bx b - ax a) = q exactly -> B v == v in every bit; L-SR1:
 * y - B s = 0 rejects the pair -> identity; inverse: H = I mathematically, the two-loop recursion returns v to Float32 rounding.
 * A closed-form check that needs no oracle. */
static void test_qn_f32(void) {
  const int64_t n = 10007;
  float *hs = (float *)malloc((size_t)n * 4), *hv = (float *)malloc((size_t)n * 4), *got = (float *)malloc((size_t)n * 4);
  void *s, *v, *res;
  OK(b2o_malloc(ctx, (size_t)n * 4, &s));
  OK(b2o_malloc(ctx, (size_t)n * 4, &v));
  OK(b2o_malloc(ctx, (size_t)n * 4, &res));
  for (int64_t i = 0; i < n; ++i) {
    hs[i] = 1.0f;
    hv[i] = (i & 1) ? 1.0f : -1.0f;
  }
  h2d(s, hs, (size_t)n * 4);
  h2d(v, hv, (size_t)n * 4);
  for (int kind = 0; kind < 3; ++kind) {
    b2o_qn *op = NULL;
    if (kind < 2) OK(b2o_lbfgs_create(ctx, B2O_F32, n, 5, 1, 0, 0.99, 10.0, kind == 1, &op));
    else OK(b2o_lsr1_create(ctx, B2O_F32, n, 5, 1, &op));
    OK(b2o_qn_apply(op, res, n, v, n, 1.0, 0.0));
    d2h(got, res, (size_t)n * 4);
    CHECK(memcmp(got, hv, (size_t)n * 4) == 0, "Float32 kind %d: identity before the first push", kind);
    int acc = -1;
    OK(b2o_qn_push(op, s, s, n, &acc));
    CHECK(acc == (kind < 2 ? 1 : 0), "Float32 kind %d: push!(op, ones, ones) acceptance %d", kind, acc);
    OK(b2o_qn_apply(op, res, n, v, n, 1.0, 0.0));
    d2h(got, res, (size_t)n * 4);
    if (kind != 1) {
      CHECK(memcmp(got, hv, (size_t)n * 4) == 0, "Float32 kind %d: B v == v after push!(op, ones, ones)", kind);
    } else {
      double worst = 0;
      for (int64_t i = 0; i < n; ++i) worst = fmax(worst, fabs((double)got[i] - (double)hv[i]));
      CHECK(worst <= 1e-6, "Float32 inverse: H v = v to rounding, worst %g", worst);
    }
    OK(b2o_qn_apply(op, res, n, v, n, 2.0, 0.0));
    d2h(got, res, (size_t)n * 4);
    double worst2 = 0;
    for (int64_t i = 0; i < n; ++i) worst2 = fmax(worst2, fabs((double)got[i] - 2.0 * (double)hv[i]));
    CHECK(worst2 <= (kind == 1 ? 2e-6 : 0.0), "Float32 kind %d: alpha = 2, worst %g", kind, worst2);
    double bytes = 0;
    OK(b2o_qn_apply_bytes(op, 0.0, &bytes));
    CHECK(bytes == (kind == 0 ? 4.0 + 3 : kind == 1 ? 8.0 + 2 : 2.0) * 4.0 * (double)n, "Float32 algorithmic bytes %g", bytes);
    if (kind != 1) {   /* diag!: 1/gamma + b^2 - a^2 with a == b (forward), the identity's diagonal (L-SR1 pair rejected) */
      OK(b2o_qn_diag(op, res, n));
      d2h(got, res, (size_t)n * 4);
      CHECK(memcmp(got, hs, (size_t)n * 4) == 0, "Float32 kind %d: diag! == ones", kind);
    } else {
      CHECK(b2o_qn_diag(op, res, n) == B2O_ESTATE, "diag! of an inverse operator must be refused");
    }
    OK(b2o_qn_apply_host(op, got, hv, n, 1.0, 0.0));   /* host buffers: staged copy-in, apply, copy-out */
    {
      double worst3 = 0;
      for (int64_t i = 0; i < n; ++i) worst3 = fmax(worst3, fabs((double)got[i] - (double)hv[i]));
      CHECK(worst3 <= (kind == 1 ? 1e-6 : 0.0), "Float32 kind %d: apply_host, worst %g", kind, worst3);
    }
    CHECK(b2o_qn_set_option(op, "forward_mode", 1) == B2O_EUNSUPPORTED || kind != 0, "compact forms are Float64 only");
    OK(b2o_qn_destroy(op));
  }
  b2o_qn *bad = NULL;
  CHECK(b2o_lbfgs_create(ctx, B2O_BF16, n, 5, 1, 0, 0.99, 10.0, 0, &bad) == B2O_EUNSUPPORTED, "bf16 quasi-Newton must be refused");
  OK(b2o_free(ctx, s)); OK(b2o_free(ctx, v)); OK(b2o_free(ctx, res));
  free(hs); free(hv); free(got);
}

/* ---- BASELINE config 3 as a fused static tree: (opHouseholder(h) * opDiagonal(d) + 0.1 * opEye(n)) * v --------- */
static void test_graph(void) {
  const int64_t n = 300007;
  double *h = dev_f64(n), *d = dev_f64(n), *v = dev_f64(n), *res = dev_f64(n);
  double *hh = host_f64(n), *hd = host_f64(n), *hv = host_f64(n), *t1 = host_f64(n), *ref = host_f64(n), *got = host_f64(n);
  orc_fill_uniform(hh, n, 3, 0.0, 1.0);
  orc_fill_uniform(hd, n, 4, 0.5, 1.5);
  orc_fill_uniform(hv, n, 5, 0.0, 1.0);
  const double nrm = orc_nrm2(hh, n);
  for (int64_t i = 0; i < n; ++i) hh[i] /= nrm;
  h2d(h, hh, (size_t)n * 8);
  h2d(d, hd, (size_t)n * 8);
  h2d(v, hv, (size_t)n * 8);
  b2o_graph *g = NULL;
  int H = 0, D = 0, E = 0, P = 0, S = 0, R = 0;
  OK(b2o_graph_create(ctx, n, &g));
  OK(b2o_graph_leaf(g, 4, h, &H));
  OK(b2o_graph_leaf(g, 0, d, &D));
  OK(b2o_graph_leaf(g, 1, NULL, &E));
  OK(b2o_graph_binary(g, 11, H, D, &P));
  OK(b2o_graph_unary(g, 12, E, 0.1, &S));
  OK(b2o_graph_binary(g, 10, P, S, &R));
  OK(b2o_graph_compile(g, R));
  OK(b2o_graph_apply(g, 0, res, n, v, n, 1.0, 0.0));
  /* the closure tree as the reference evaluates it: sum_prod! (src/operations.jl:187-197) = mul!(res, H*D, v, 1, 0) then
   * mul!(res, 0.1*I, v, 1, 1);  prod_op! (:117-128) = mul!(tmp, D, v); mul!(res, H, tmp, 1, 0) */
  orc_diag_square(t1, hd, hv, n, 1.0, 0.0);
  orc_householder(ref, hh, t1, n, 1.0, 0.0);
  orc_eye(ref, n, hv, n, 0.1, 1.0, n);
  d2h(got, res, (size_t)n * 8);
  CHECK(rel_err(got, ref, n) <= 1e-12, "fused cfg3 chain: %g", rel_err(got, ref, n));
  int npass = 0, nred = 0;
  double bytes = 0;
  OK(b2o_graph_info(g, 0, 0.0, &npass, &nred, &bytes));
  CHECK(npass == 2 && nred == 1 && bytes == 7.0 * 8.0 * (double)n, "passes %d reductions %d bytes %g", npass, nred, bytes);
  OK(b2o_graph_destroy(g));
  OK(b2o_free(ctx, h)); OK(b2o_free(ctx, d)); OK(b2o_free(ctx, v)); OK(b2o_free(ctx, res));
  free(hh); free(hd); free(hv); free(t1); free(ref); free(got);
}

/* ---- kron(A, B) * x on the tensor cores: bf16 operands, fp32 result, against the Float64 oracle ---------------- */
static void test_kron(void) {
  const int64_t m = 72, n = 40, p = 136, q = 64;          /* ragged: not multiples of the 64 / 128 tile sizes */
  const int64_t nx = n * q, nr = m * p;
  double *Ad = host_f64(m * n), *Bd = host_f64(p * q), *xd = host_f64(nx), *ref = host_f64(nr), *got = host_f64(nr);
  uint16_t *Ab = (uint16_t *)malloc((size_t)(m * n) * 2), *Bb = (uint16_t *)malloc((size_t)(p * q) * 2), *xb = (uint16_t *)malloc((size_t)nx * 2);
  float *rf = (float *)malloc((size_t)nr * 4);
  orc_fill_uniform(Ad, m * n, 11, -1.0, 1.0);
  orc_fill_uniform(Bd, p * q, 12, -1.0, 1.0);
  orc_fill_uniform(xd, nx, 13, -1.0, 1.0);
  for (int64_t i = 0; i < m * n; ++i) { Ab[i] = orc_f32_to_bf16((float)Ad[i]); Ad[i] = orc_bf16_to_f32(Ab[i]); }
  for (int64_t i = 0; i < p * q; ++i) { Bb[i] = orc_f32_to_bf16((float)Bd[i]); Bd[i] = orc_bf16_to_f32(Bb[i]); }
  for (int64_t i = 0; i < nx; ++i) { xb[i] = orc_f32_to_bf16((float)xd[i]); xd[i] = orc_bf16_to_f32(xb[i]); }
  void *dA = NULL, *dB = NULL, *dx = NULL, *dr = NULL;
  OK(b2o_malloc(ctx, (size_t)(m * n) * 2, &dA));
  OK(b2o_malloc(ctx, (size_t)(p * q) * 2, &dB));
  OK(b2o_malloc(ctx, (size_t)nx * 2, &dx));
  OK(b2o_malloc(ctx, (size_t)nr * 4, &dr));
  h2d(dA, Ab, (size_t)(m * n) * 2);
  h2d(dB, Bb, (size_t)(p * q) * 2);
  h2d(dx, xb, (size_t)nx * 2);
  b2o_kron *K = NULL;
  OK(b2o_kron_create(ctx, B2O_BF16, dA, m, n, dB, p, q, 1, &K));     /* column-major A (m x n), B (p x q): Julia layout */
  OK(b2o_kron_apply(K, 0, dr, B2O_F32, nr, dx, nx, 1, 1.0, 0.0));
  orc_kron(ref, Ad, m, n, Bd, p, q, xd, 1.0, 0.0, 0);
  d2h(rf, dr, (size_t)nr * 4);
  for (int64_t i = 0; i < nr; ++i) got[i] = rf[i];
  CHECK(rel_err(got, ref, nr) <= 1e-5, "kron prod! (fp32 result): %g", rel_err(got, ref, nr));
  double fl = 0;
  OK(b2o_kron_flops(K, 1, &fl));
  CHECK(fl == 2.0 * (double)(p * q * n) + 2.0 * (double)(p * n * m), "flops %g", fl);
  CHECK(b2o_kron_apply(K, 0, dr, B2O_F32, nr, dx, nx - 8, 1, 1.0, 0.0) == B2O_ESHAPE, "kron shape check");
  OK(b2o_kron_destroy(K));
  OK(b2o_free(ctx, dA)); OK(b2o_free(ctx, dB)); OK(b2o_free(ctx, dx)); OK(b2o_free(ctx, dr));
  free(Ad); free(Bd); free(xd); free(ref); free(got); free(Ab); free(Bb); free(xb); free(rf);
}

int main(void) {
  if (b2o_version() < 100) return 3;
  int st = b2o_ctx_create(0, NULL, &ctx);
  if (st != B2O_OK) {
    /* no GPU: the library must say so and refuse -- there is no CPU fallback */
    printf("ABI_DRIVER_NO_GPU status=%d msg=%s\n", st, b2o_last_error());
    return strstr(b2o_last_error(), "no CPU fallback") ? 77 : 4;
  }
  test_leaves();
  test_index();
  test_qn();
  test_qn_f32();
  test_graph();
  test_kron();
  int64_t launches = 0;
  OK(b2o_ctx_launch_count(ctx, &launches));
  CHECK(launches > 50, "the CUDA path must have run: %lld launches", (long long)launches);
  OK(b2o_ctx_destroy(ctx));
  printf("ABI_DRIVER_OK %d checks, %lld kernel launches\n", g_checks, (long long)launches);
  return 0;
}
