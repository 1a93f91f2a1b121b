/*
 * b2o_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).  See b2o_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (oracle/Makefile).
 * -ffp-contract=off matters: Julia does not contract a*b+c into an fma, so the
 * elementwise statements below round exactly like the reference's broadcasts.
 *
 * Index convention: Julia's 1-based ring index `insert` is kept 0-based internally
 * (ins0 = insert-1); every loop states the Julia formula it restates.
 */
#include "b2o_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_accurate = 1;
static int g_threads = 1;

void orc_set_mode(int accurate, int threads) {
  g_accurate = accurate;
  g_threads = threads < 1 ? 1 : threads;
}
int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- deterministic generator (shared with csrc/b2o_util.cu, bit for bit) ---- */
static inline uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static inline double u01(uint64_t seed, uint64_t i) {
  uint64_t z = mix64((i + 1) * 0x9E3779B97F4A7C15ULL + seed * 0xD1B54A32D192ED03ULL);
  return (double)(z >> 11) * 0x1.0p-53;
}
void orc_fill_uniform(double *x, int64_t n, uint64_t seed, double lo, double hi) {
  double w = hi - lo;
#pragma omp parallel for num_threads(g_threads) if (g_threads > 1)
  for (int64_t i = 0; i < n; ++i) x[i] = lo + w * u01(seed, (uint64_t)i);
}
void orc_fill_uniform_f32(float *x, int64_t n, uint64_t seed, float lo, float hi) {
  for (int64_t i = 0; i < n; ++i) x[i] = (float)((double)lo + ((double)hi - (double)lo) * u01(seed, (uint64_t)i));
}

/* ---- reductions: `dot`, `sum`, `norm` (OpenBLAS in the reference; order unspecified) ---- */
double orc_dot(const double *a, const double *b, int64_t n) {
  if (g_accurate) {
    long double acc = 0.0L;
    for (int64_t i = 0; i < n; ++i) acc += (long double)a[i] * (long double)b[i];
    return (double)acc;
  }
  double acc = 0.0;
#pragma omp parallel for reduction(+ : acc) num_threads(g_threads) if (g_threads > 1)
  for (int64_t i = 0; i < n; ++i) acc += a[i] * b[i];
  return acc;
}
double orc_sum(const double *a, int64_t n) {
  if (g_accurate) {
    long double acc = 0.0L;
    for (int64_t i = 0; i < n; ++i) acc += (long double)a[i];
    return (double)acc;
  }
  double acc = 0.0;
#pragma omp parallel for reduction(+ : acc) num_threads(g_threads) if (g_threads > 1)
  for (int64_t i = 0; i < n; ++i) acc += a[i];
  return acc;
}
double orc_nrm2(const double *a, int64_t n) { return sqrt(orc_dot(a, a, n)); }

#define PFOR _Pragma("omp parallel for num_threads(g_threads) if (g_threads > 1)")

/* ---- mulOpEye!  src/special-operators.jl:36-44 (Q2: tail is set to beta, not beta*res) ---- */
void orc_eye(double *res, int64_t nres, const double *v, int64_t nv, double alpha, double beta, int64_t n_min) {
  (void)nv;
  if (beta == 0.0) {
    PFOR for (int64_t i = 0; i < n_min; ++i) res[i] = alpha * v[i];
    for (int64_t i = n_min; i < nres; ++i) res[i] = 0.0;
  } else {
    PFOR for (int64_t i = 0; i < n_min; ++i) res[i] = alpha * v[i] + beta * res[i];
    for (int64_t i = n_min; i < nres; ++i) res[i] = beta;
  }
}

/* ---- mulOpOnes!  src/special-operators.jl:79-85 ---- */
void orc_ones(double *res, int64_t nres, const double *v, int64_t nv, double alpha, double beta) {
  double c = alpha * orc_sum(v, nv);
  if (beta == 0.0) {
    PFOR for (int64_t i = 0; i < nres; ++i) res[i] = c;
  } else {
    PFOR for (int64_t i = 0; i < nres; ++i) res[i] = c + beta * res[i];
  }
}

/* ---- mulOpZeros!  src/special-operators.jl:102-108 ---- */
void orc_zeros(double *res, int64_t nres, double alpha, double beta) {
  (void)alpha;
  if (beta == 0.0) {
    PFOR for (int64_t i = 0; i < nres; ++i) res[i] = 0.0;
  } else {
    PFOR for (int64_t i = 0; i < nres; ++i) res[i] *= beta;
  }
}

/* ---- mulSquareOpDiagonal!  src/special-operators.jl:125-131; `α .* d .* v` = (α*d)*v ---- */
void orc_diag_square(double *res, const double *d, const double *v, int64_t n, double alpha, double beta) {
  if (beta == 0.0) {
    PFOR for (int64_t i = 0; i < n; ++i) res[i] = (alpha * d[i]) * v[i];
  } else {
    PFOR for (int64_t i = 0; i < n; ++i) res[i] = (alpha * d[i]) * v[i] + beta * res[i];
  }
}

/* ---- mulOpDiagonal!  src/special-operators.jl:144-151 (Q3: tail zeroed even if beta != 0) ---- */
void orc_diag_rect(double *res, int64_t nres, const double *d, const double *v, double alpha, double beta, int64_t n_min) {
  orc_diag_square(res, d, v, n_min, alpha, beta);
  for (int64_t i = n_min; i < nres; ++i) res[i] = 0.0;
}

/* ---- mulHouseholder!  src/linalg.jl:77-83: res = α .* (v .- 2*dot(h,v) .* h) (.+ β .* res) ---- */
void orc_householder(double *res, const double *h, const double *v, int64_t n, double alpha, double beta) {
  double t = 2 * orc_dot(h, v, n);
  if (beta == 0.0) {
    PFOR for (int64_t i = 0; i < n; ++i) res[i] = alpha * (v[i] - t * h[i]);
  } else {
    PFOR for (int64_t i = 0; i < n; ++i) res[i] = alpha * (v[i] - t * h[i]) + beta * res[i];
  }
}

/* ---- ComplexF64 leaves (interleaved re, im).  Julia's Complex `*` is (ac - bd) + (ad + bc)i, plain multiplies and adds. ---- */
typedef struct { double re, im; } orc_c;
static inline orc_c c_mul(orc_c a, orc_c b) { orc_c r = {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; return r; }
static inline orc_c c_add(orc_c a, orc_c b) { orc_c r = {a.re + b.re, a.im + b.im}; return r; }
static inline orc_c c_sub(orc_c a, orc_c b) { orc_c r = {a.re - b.re, a.im - b.im}; return r; }
/* mulSquareOpDiagonal! with complex T (src/special-operators.jl:125-131); conj_d: the ctprod! closure's conj.(d) (:140) */
void orc_cdiag(double *res_, const double *d_, const double *v_, int64_t n, int conj_d, const double *alpha, const double *beta) {
  orc_c *res = (orc_c *)res_;
  const orc_c *d = (const orc_c *)d_, *v = (const orc_c *)v_;
  const orc_c a = {alpha[0], alpha[1]}, b = {beta[0], beta[1]};
  const int bz = (b.re == 0.0 && b.im == 0.0);
  for (int64_t i = 0; i < n; ++i) {
    orc_c di = d[i];
    if (conj_d) di.im = -di.im;
    orc_c t = c_mul(c_mul(a, di), v[i]);
    res[i] = bz ? t : c_add(t, c_mul(b, res[i]));
  }
}
/* mulHouseholder! with complex T (src/linalg.jl:77-83): dot(h, v) = sum conj(h_i) v_i */
void orc_chouseholder(double *res_, const double *h_, const double *v_, int64_t n, const double *alpha, const double *beta) {
  orc_c *res = (orc_c *)res_;
  const orc_c *h = (const orc_c *)h_, *v = (const orc_c *)v_;
  const orc_c a = {alpha[0], alpha[1]}, b = {beta[0], beta[1]};
  const int bz = (b.re == 0.0 && b.im == 0.0);
  long double re = 0.0L, im = 0.0L;
  for (int64_t i = 0; i < n; ++i) {
    re += (long double)h[i].re * v[i].re + (long double)h[i].im * v[i].im;
    im += (long double)h[i].re * v[i].im - (long double)h[i].im * v[i].re;
  }
  const orc_c two = {2.0, 0.0}, dot = {(double)re, (double)im};
  const orc_c tau = c_mul(two, dot);
  for (int64_t i = 0; i < n; ++i) {
    orc_c t = c_mul(a, c_sub(v[i], c_mul(tau, h[i])));
    res[i] = bz ? t : c_add(t, c_mul(b, res[i]));
  }
}
/* conj!(res) / conj.(v) (src/adjtrans.jl:128-136) */
void orc_conj(double *dst, const double *src, int64_t n) {
  for (int64_t i = 0; i < n; ++i) { dst[2 * i] = src[2 * i]; dst[2 * i + 1] = -src[2 * i + 1]; }
}

/* ---- mulRestrict! / multRestrict!  src/special-operators.jl:167-174 (Q1 ignore α,β; Q4 last wins) ---- */
void orc_restrict(double *res, const int64_t *idx1, int64_t k, const double *v) {
  for (int64_t i = 0; i < k; ++i) res[i] = v[idx1[i] - 1];
}
void orc_extend(double *res, int64_t ncol, const int64_t *idx1, int64_t k, const double *u) {
  for (int64_t i = 0; i < ncol; ++i) res[i] = 0.0;
  for (int64_t i = 0; i < k; ++i) res[idx1[i] - 1] = u[i];
}

/* ================= L-BFGS  src/lbfgs.jl ================= */
struct orc_lbfgs {
  int64_t n;
  int mem, scaling, damped, inverse;
  double gamma, sigma2, sigma3, opnorm_ub; /* scaling_factor, σ₂, σ₃, opnorm_upper_bound  :4-24 */
  double *s, *y, *a, *b;                   /* mem columns of n, column k0 at +k0*n */
  double *ys, *alpha, *norm_b;
  int ins0;                                /* insert-1 */
  double *Ax;
  /* full-size parity runs (tests only): the columns live on the GPU and are streamed to the host one at a time */
  orc_fetch_fn fetch;
  void *fetch_user;
  double *dots_log;                        /* inner products of the last apply, in the order the reference takes them */
  int ndots;
};

/* column k0 of s/y/a/b (which = 0..3): resident, or fetched into the caller's host buffer `slot` (0 or 1) */
static const double *lbfgs_column(orc_lbfgs *o, int which, int k0, int slot) {
  if (o->fetch) return o->fetch(o->fetch_user, which, k0, slot);
  const double *base = which == 0 ? o->s : which == 1 ? o->y : which == 2 ? o->a : o->b;
  return base + (size_t)k0 * (size_t)o->n;
}
static double logged_dot(orc_lbfgs *o, const double *a, const double *b) {
  double d = orc_dot(a, b, o->n);
  if (o->dots_log && o->ndots < 4 * o->mem + 4) o->dots_log[o->ndots++] = d;
  return d;
}

static inline int pmod(int a, int m) { int r = a % m; return r < 0 ? r + m : r; }

static orc_lbfgs *lbfgs_create(int64_t n, int mem, int scaling, int damped, double sigma2, double sigma3, int inverse, int external);
orc_lbfgs *orc_lbfgs_create(int64_t n, int mem, int scaling, int damped, double sigma2, double sigma3, int inverse) {
  return lbfgs_create(n, mem, scaling, damped, sigma2, sigma3, inverse, 0);
}
/* apply-only operator whose columns are supplied by `fetch` (state scalars through set_gamma / set_insert / ys) */
orc_lbfgs *orc_lbfgs_create_external(int64_t n, int mem, int scaling, int inverse, orc_fetch_fn fetch, void *user) {
  orc_lbfgs *o = lbfgs_create(n, mem, scaling, 0, 0.99, 10.0, inverse, 1);
  o->fetch = fetch;
  o->fetch_user = user;
  return o;
}
const double *orc_lbfgs_last_dots(orc_lbfgs *o, int *count) {
  if (count) *count = o->ndots;
  return o->dots_log;
}
static orc_lbfgs *lbfgs_create(int64_t n, int mem, int scaling, int damped, double sigma2, double sigma3, int inverse, int external) {
  /* LBFGSData ctor :26-57.  Q7: the reference clamps data.mem to max(mem,1) but sizes its arrays with the
   * unclamped value (mem=0 would index out of bounds); the oracle sizes with the clamped value. */
  orc_lbfgs *o = (orc_lbfgs *)calloc(1, sizeof(*o));
  if (mem < 1) mem = 1;
  o->n = n; o->mem = mem; o->scaling = scaling; o->damped = damped; o->inverse = inverse;
  o->gamma = 1.0; o->sigma2 = sigma2; o->sigma3 = sigma3; o->opnorm_ub = 1.0;
  size_t nn = (size_t)(n > 0 ? n : 1) * (size_t)mem;
  o->dots_log = (double *)calloc((size_t)4 * mem + 4, sizeof(double));
  if (!external) {
    o->s = (double *)calloc(nn, sizeof(double));
    o->y = (double *)calloc(nn, sizeof(double));
  }
  if (!inverse && !external) {
    o->a = (double *)calloc(nn, sizeof(double));
    o->b = (double *)calloc(nn, sizeof(double));
  }
  o->ys = (double *)calloc(mem, sizeof(double));
  o->alpha = (double *)calloc(mem, sizeof(double));
  o->norm_b = (double *)calloc(mem, sizeof(double));
  o->Ax = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  o->ins0 = 0;
  return o;
}
void orc_lbfgs_destroy(orc_lbfgs *o) {
  if (!o) return;
  free(o->s); free(o->y); free(o->a); free(o->b); free(o->ys); free(o->alpha); free(o->norm_b); free(o->Ax); free(o->dots_log); free(o);
}
double *orc_lbfgs_col(orc_lbfgs *o, int which, int k0) {
  double *base = which == 0 ? o->s : which == 1 ? o->y : which == 2 ? o->a : o->b;
  return base ? base + (size_t)k0 * (size_t)o->n : NULL;
}
double *orc_lbfgs_ys(orc_lbfgs *o) { return o->ys; }
double orc_lbfgs_gamma(orc_lbfgs *o) { return o->gamma; }
void orc_lbfgs_set_gamma(orc_lbfgs *o, double g) { o->gamma = g; }
int orc_lbfgs_insert(orc_lbfgs *o) { return o->ins0 + 1; }
void orc_lbfgs_set_insert(orc_lbfgs *o, int insert1) { o->ins0 = insert1 - 1; }
double orc_lbfgs_opnorm_upper_bound(orc_lbfgs *o) { return o->opnorm_ub; }

/* inverse two-loop  src/lbfgs.jl:117-154 */
static void lbfgs_apply_inverse(orc_lbfgs *o, double *res, const double *x, double am, double bm) {
  const int64_t n = o->n;
  const int mem = o->mem;
  double *q = o->Ax;
  o->ndots = 0;
  PFOR for (int64_t j = 0; j < n; ++j) q[j] = x[j];                         /* :127-128 */
  for (int i = 1; i <= mem; ++i) {                                          /* :130 */
    int k = pmod(o->ins0 - i, mem);                                         /* k = mod(insert-i-1,mem)+1 */
    if (o->ys[k] != 0) {
      const double *sk = lbfgs_column(o, 0, k, 0), *yk = lbfgs_column(o, 1, k, 1);
      double ak = logged_dot(o, sk, q) / o->ys[k];                          /* :133 */
      o->alpha[k] = ak;
      PFOR for (int64_t j = 0; j < n; ++j) q[j] -= ak * yk[j];              /* :135 */
    }
  }
  if (o->scaling) {
    double g = o->gamma;
    PFOR for (int64_t j = 0; j < n; ++j) q[j] *= g;                         /* :139 */
  }
  for (int i = 1; i <= mem; ++i) {                                          /* :141 */
    int k = pmod(o->ins0 + i - 1, mem);                                     /* k = mod(insert+i-2,mem)+1 */
    if (o->ys[k] != 0) {
      const double *sk = lbfgs_column(o, 0, k, 0), *yk = lbfgs_column(o, 1, k, 1);
      double bb = o->alpha[k] - logged_dot(o, yk, q) / o->ys[k];            /* :144-145 */
      PFOR for (int64_t j = 0; j < n; ++j) q[j] += bb * sk[j];              /* :146 */
    }
  }
  if (bm == 0.0) {
    PFOR for (int64_t j = 0; j < n; ++j) res[j] = am * q[j];                /* :150 */
  } else {
    PFOR for (int64_t j = 0; j < n; ++j) res[j] = am * q[j] + bm * res[j];  /* :152 */
  }
}

/* forward compact form  src/lbfgs.jl:173-202 (Q9: dots against x; `q .+= bx.*b .- ax.*a`) */
static void lbfgs_apply_forward(orc_lbfgs *o, double *res, const double *x, double al, double be) {
  const int64_t n = o->n;
  const int mem = o->mem;
  double *q = o->Ax;
  o->ndots = 0;
  PFOR for (int64_t j = 0; j < n; ++j) q[j] = x[j];                         /* :183-184 */
  if (o->scaling) {
    double g = o->gamma;
    PFOR for (int64_t j = 0; j < n; ++j) q[j] /= g;                         /* :186 */
  }
  for (int i = 1; i <= mem; ++i) {                                          /* :189 */
    int k = pmod(o->ins0 + i - 1, mem);
    if (o->ys[k] != 0) {
      const double *ak = lbfgs_column(o, 2, k, 0), *bk = lbfgs_column(o, 3, k, 1);
      double ax = logged_dot(o, ak, x);                                     /* :192 */
      double bx = logged_dot(o, bk, x);                                     /* :193 */
      PFOR for (int64_t j = 0; j < n; ++j) q[j] = q[j] + (bx * bk[j] - ax * ak[j]); /* :194 */
    }
  }
  if (be == 0.0) {
    PFOR for (int64_t j = 0; j < n; ++j) res[j] = al * q[j];                /* :198 */
  } else {
    PFOR for (int64_t j = 0; j < n; ++j) res[j] = al * q[j] + be * res[j];  /* :200 */
  }
}

void orc_lbfgs_apply(orc_lbfgs *o, double *res, const double *x, double alpha, double beta) {
  if (o->inverse) lbfgs_apply_inverse(o, res, x, alpha, beta);
  else lbfgs_apply_forward(o, res, x, alpha, beta);
}

/* push_common!  src/lbfgs.jl:210-255 */
static void lbfgs_push_common(orc_lbfgs *o, const double *s, const double *y, double ys) {
  const int64_t n = o->n;
  const int mem = o->mem;
  const int ins = o->ins0;
  double *si = o->s + (size_t)ins * n, *yi = o->y + (size_t)ins * n;
  memcpy(si, s, (size_t)n * sizeof(double));                                /* :220 */
  memcpy(yi, y, (size_t)n * sizeof(double));                                /* :221 */
  o->ys[ins] = ys;                                                          /* :222 */
  if (o->scaling) {                                                         /* :223-227 */
    if (o->gamma != 0) o->opnorm_ub -= 1 / o->gamma;
    o->gamma = ys / orc_dot(y, y, n);
    if (o->gamma != 0) o->opnorm_ub += 1 / o->gamma;
  }
  if (!o->inverse) {                                                        /* :230 */
    double *bi = o->b + (size_t)ins * n;
    o->opnorm_ub -= o->norm_b[ins] * o->norm_b[ins];
    double rt = sqrt(ys);
    for (int64_t j = 0; j < n; ++j) bi[j] = y[j] / rt;                      /* :232 */
    o->norm_b[ins] = orc_nrm2(bi, n);
    o->opnorm_ub += o->norm_b[ins] * o->norm_b[ins];
    for (int i = 1; i <= mem; ++i) {                                        /* :236 */
      int k = pmod(ins + i, mem);                                           /* k = mod(insert+i-1,mem)+1 */
      if (o->ys[k] != 0) {
        double *ak = o->a + (size_t)k * n;
        const double *sk = o->s + (size_t)k * n;
        double g = o->gamma;
        for (int64_t j = 0; j < n; ++j) ak[j] = sk[j] / g;                  /* :239 */
        for (int jj = 1; jj <= i - 1; ++jj) {                               /* :241 */
          int l = pmod(ins + jj, mem);
          if (o->ys[l] != 0) {
            const double *bl = o->b + (size_t)l * n, *al = o->a + (size_t)l * n;
            double c1 = orc_dot(bl, sk, n);
            for (int64_t j = 0; j < n; ++j) ak[j] += c1 * bl[j];            /* :244 */
            double c2 = orc_dot(al, sk, n);
            for (int64_t j = 0; j < n; ++j) ak[j] -= c2 * al[j];            /* :245 */
          }
        }
        double nrm = sqrt(orc_dot(sk, ak, n));
        for (int64_t j = 0; j < n; ++j) ak[j] /= nrm;                       /* :248 */
      }
    }
  }
  o->ins0 = pmod(ins + 1, mem);                                             /* :253 insert = mod(insert,mem)+1 */
}

/* push!(op,s,y)  src/lbfgs.jl:269-287 */
int orc_lbfgs_push(orc_lbfgs *o, const double *s, const double *y) {
  if (o->damped) {
    if (o->inverse) return -1; /* push!(op,s,y,Bs) on an inverse operator errors :296-298 */
    double *Bs = (double *)malloc((size_t)o->n * sizeof(double));
    int r = orc_lbfgs_push_damped_fwd(o, s, y, Bs);
    free(Bs);
    return r;
  }
  double ys = orc_dot(y, s, o->n);
  if (ys <= DBL_EPSILON) return 0;                                          /* :281 */
  lbfgs_push_common(o, s, y, ys);
  return 1;
}

/* push!(op,s,y,Bs)  src/lbfgs.jl:289-321 */
int orc_lbfgs_push_damped_fwd(orc_lbfgs *o, const double *s, const double *y, double *Bs) {
  if (!o->damped || o->inverse) return -1;
  const int64_t n = o->n;
  double ys = orc_dot(y, s, n);
  orc_lbfgs_apply(o, Bs, s, 1.0, 0.0);                                      /* :305 */
  double sBs = orc_dot(s, Bs, n);
  int damp = 0;
  double th = 0;
  if (ys < (1 - o->sigma2) * sBs) { th = o->sigma2 * sBs / (sBs - ys); damp = 1; }
  else if (ys > (1 + o->sigma3) * sBs) { th = o->sigma3 * sBs / (ys - sBs); damp = 1; }
  if (damp) {
    double *yd = (double *)malloc((size_t)n * sizeof(double));
    for (int64_t j = 0; j < n; ++j) yd[j] = th * y[j] + (1 - th) * Bs[j];   /* :316 */
    ys = th * ys + (1 - th) * sBs;
    lbfgs_push_common(o, s, yd, ys);
    free(yd);
  } else {
    lbfgs_push_common(o, s, y, ys);
  }
  return 1;
}

/* push!(op,s,y,α,g,Bs)  src/lbfgs.jl:323-357 (mutates y) */
int orc_lbfgs_push_damped_inv(orc_lbfgs *o, const double *s, double *y, double alpha, const double *g, double *Bs) {
  if (!o->damped || !o->inverse) return -1;
  const int64_t n = o->n;
  double ys = orc_dot(y, s, n);
  for (int64_t j = 0; j < n; ++j) Bs[j] = -alpha * g[j];                    /* :341 */
  double sBs = orc_dot(s, Bs, n);
  int damp = 0;
  double th = 0;
  if (ys < (1 - o->sigma2) * sBs) { th = o->sigma2 * sBs / (sBs - ys); damp = 1; }
  else if (ys > (1 + o->sigma3) * sBs) { th = o->sigma3 * sBs / (ys - sBs); damp = 1; }
  if (damp) {
    for (int64_t j = 0; j < n; ++j) y[j] = th * y[j] + (1 - th) * Bs[j];    /* :352 */
    ys = th * ys + (1 - th) * sBs;
  }
  lbfgs_push_common(o, s, y, ys);
  return 1;
}

/* diag!  src/lbfgs.jl:379-395 */
int orc_lbfgs_diag(orc_lbfgs *o, double *d) {
  if (o->inverse) return -1;
  const int64_t n = o->n;
  for (int64_t j = 0; j < n; ++j) d[j] = 1.0;
  if (o->scaling) for (int64_t j = 0; j < n; ++j) d[j] /= o->gamma;
  for (int i = 1; i <= o->mem; ++i) {
    int k = pmod(o->ins0 + i - 1, o->mem);
    if (o->ys[k] != 0) {
      const double *ak = o->a + (size_t)k * n, *bk = o->b + (size_t)k * n;
      for (int64_t j = 0; j < n; ++j) d[j] = d[j] + (bk[j] * bk[j] - ak[j] * ak[j]); /* :391 */
    }
  }
  return 0;
}

/* reset!  src/lbfgs.jl:401-427 (Q8: opnorm_upper_bound / norm_b untouched) */
void orc_lbfgs_reset(orc_lbfgs *o) {
  size_t nn = (size_t)o->n * (size_t)o->mem;
  memset(o->s, 0, nn * sizeof(double));
  memset(o->y, 0, nn * sizeof(double));
  if (!o->inverse) { memset(o->a, 0, nn * sizeof(double)); memset(o->b, 0, nn * sizeof(double)); }
  memset(o->ys, 0, o->mem * sizeof(double));
  memset(o->alpha, 0, o->mem * sizeof(double));
  o->gamma = 1.0;
  o->ins0 = 0;
}

/* solve_shifted_system!  src/utilities.jl:207-248 (Erway, Jain, Marcia 2014): 2*mem rank-one Sherman-Morrison steps */
int orc_lbfgs_solve_shifted(orc_lbfgs *o, double *x, const double *b, double sigma) {
  if (sigma < 0 || o->inverse) return -1;                                       /* :213-215 */
  const int64_t n = o->n;
  const int mem = o->mem, max_i = 2 * mem;
  double *P = (double *)calloc((size_t)n * max_i, sizeof(double));               /* data.shifted_p  n x 2mem */
  double *v = (double *)calloc(max_i, sizeof(double));                           /* data.shifted_v */
  const double gamma_inv = 1 / o->gamma;                                         /* :219 */
  const double x_0 = 1 / (gamma_inv + sigma);
  for (int64_t j = 0; j < n; ++j) x[j] = x_0 * b[j];                             /* :221 */
  int sign_i = 1;
  for (int i = 1; i <= max_i; ++i) {                                             /* :226 */
    const int jj = (i + 1) / 2;
    const int k = pmod(o->ins0 + jj, mem);                                       /* k = mod(insert + j - 1, mem) + 1 */
    const double *u = (sign_i == -1) ? o->b + (size_t)k * n : o->a + (size_t)k * n;   /* :229 */
    double *pi = P + (size_t)(i - 1) * n;
    for (int64_t j = 0; j < n; ++j) pi[j] = x_0 * u[j];                          /* :231 */
    int sign_t = 1;
    for (int t = 1; t <= i - 1; ++t) {                                           /* :234 */
      const double *pt = P + (size_t)(t - 1) * n;
      const double c0 = orc_dot(pt, u, n);
      const double c1 = sign_t * v[t - 1];
      const double c2 = c1 * c0;
      for (int64_t j = 0; j < n; ++j) pi[j] += c2 * pt[j];                       /* :238 */
      sign_t = -sign_t;
    }
    v[i - 1] = 1 / (1 - sign_i * orc_dot(u, pi, n));                             /* :242 */
    const double cx = sign_i * v[i - 1] * orc_dot(pi, b, n);                     /* :243-244 */
    for (int64_t j = 0; j < n; ++j) x[j] += cx * pi[j];
    sign_i = -sign_i;
  }
  free(P);
  free(v);
  return 0;
}

/* ================= L-SR1  src/lsr1.jl ================= */
struct orc_lsr1 {
  int64_t n;
  int mem, scaling;
  double gamma, opnorm_ub;
  double *s, *y, *a, *ys, *as;
  int ins0;
  double *tmp;
};
orc_lsr1 *orc_lsr1_create(int64_t n, int mem, int scaling) {
  orc_lsr1 *o = (orc_lsr1 *)calloc(1, sizeof(*o));
  if (mem < 1) mem = 1;
  o->n = n; o->mem = mem; o->scaling = scaling; o->gamma = 1.0; o->opnorm_ub = 1.0;
  size_t nn = (size_t)(n > 0 ? n : 1) * (size_t)mem;
  o->s = (double *)calloc(nn, sizeof(double));
  o->y = (double *)calloc(nn, sizeof(double));
  o->a = (double *)calloc(nn, sizeof(double));
  o->ys = (double *)calloc(mem, sizeof(double));
  o->as = (double *)calloc(mem, sizeof(double));
  o->tmp = (double *)calloc((size_t)(n > 0 ? n : 1), sizeof(double));
  return o;
}
void orc_lsr1_destroy(orc_lsr1 *o) {
  if (!o) return;
  free(o->s); free(o->y); free(o->a); free(o->ys); free(o->as); free(o->tmp); free(o);
}
double *orc_lsr1_col(orc_lsr1 *o, int which, int k0) {
  double *base = which == 0 ? o->s : which == 1 ? o->y : o->a;
  return base + (size_t)k0 * (size_t)o->n;
}
double *orc_lsr1_ys(orc_lsr1 *o) { return o->ys; }
double *orc_lsr1_as(orc_lsr1 *o) { return o->as; }
double orc_lsr1_gamma(orc_lsr1 *o) { return o->gamma; }
void orc_lsr1_set_gamma(orc_lsr1 *o, double g) { o->gamma = g; }
int orc_lsr1_insert(orc_lsr1 *o) { return o->ins0 + 1; }
void orc_lsr1_set_insert(orc_lsr1 *o, int insert1) { o->ins0 = insert1 - 1; }
double orc_lsr1_opnorm_upper_bound(orc_lsr1 *o) { return o->opnorm_ub; }

/* lsr1_multiply  src/lsr1.jl:89-107 (writes into res in place; dots against x) */
void orc_lsr1_apply(orc_lsr1 *o, double *q, const double *x, double alpha, double beta) {
  const int64_t n = o->n;
  const double g = o->gamma;
  if (beta == 0.0) {
    PFOR for (int64_t j = 0; j < n; ++j) q[j] = alpha * x[j] / g;               /* :93 */
  } else {
    PFOR for (int64_t j = 0; j < n; ++j) q[j] = alpha * x[j] / g + beta * q[j]; /* :95 */
  }
  for (int i = 1; i <= o->mem; ++i) {                                            /* :98 */
    int k = pmod(o->ins0 + i - 1, o->mem);
    if (o->ys[k] != 0) {
      const double *ak = o->a + (size_t)k * n;
      double ax = alpha * orc_dot(ak, x, n) / o->as[k];                          /* :101 */
      PFOR for (int64_t j = 0; j < n; ++j) q[j] += ax * ak[j];                   /* :102-104 */
    }
  }
}

/* push!  src/lsr1.jl:119-184 */
int orc_lsr1_push(orc_lsr1 *o, const double *s, const double *y) {
  const int64_t n = o->n;
  const int mem = o->mem;
  double *t = o->tmp;
  memcpy(t, y, (size_t)n * sizeof(double));                                      /* :124 */
  orc_lsr1_apply(o, t, s, -1.0, 1.0);                                            /* :125 ymBs = y - B s */
  double ys = orc_dot(y, s, n);
  double sNorm = orc_nrm2(s, n);
  double yy = orc_dot(y, y, n);
  const double eps = DBL_EPSILON;
  int well_defined = fabs(orc_dot(t, s, n)) >= eps + eps * orc_nrm2(t, n) * sNorm; /* :131 */
  int sufficient_curvature = 1, scaling_condition = 1;
  if (o->scaling) {                                                              /* :135 */
    double yNorm = sqrt(yy);
    sufficient_curvature = fabs(ys) >= eps * yNorm * sNorm;
    if (sufficient_curvature) {
      double sf = ys / yy;
      for (int64_t j = 0; j < n; ++j) t[j] = y[j] - s[j] / sf;                   /* :140 */
      scaling_condition = orc_nrm2(t, n) >= eps * yNorm * sNorm;
    }
  }
  if (!(well_defined && sufficient_curvature && scaling_condition)) return 0;    /* :145-149 */
  int ins = o->ins0;
  memcpy(o->s + (size_t)ins * n, s, (size_t)n * sizeof(double));
  memcpy(o->y + (size_t)ins * n, y, (size_t)n * sizeof(double));
  o->ys[ins] = ys;
  o->opnorm_ub = 1.0;                                                            /* :156 */
  if (o->scaling) {
    o->gamma = ys / yy;
    if (o->gamma != 0) o->opnorm_ub = 1 / fabs(o->gamma);
  }
  o->ins0 = pmod(ins + 1, mem);                                                  /* :163 */
  for (int i = 1; i <= mem; ++i) {                                               /* :166 */
    int k = pmod(o->ins0 + i - 1, mem);
    if (o->ys[k] != 0) {
      double *ak = o->a + (size_t)k * n;
      const double *sk = o->s + (size_t)k * n, *yk = o->y + (size_t)k * n;
      double g = o->gamma;
      for (int64_t j = 0; j < n; ++j) ak[j] = yk[j] - sk[j] / g;                 /* :169 */
      for (int jj = 1; jj <= i - 1; ++jj) {
        int l = pmod(o->ins0 + jj - 1, mem);
        if (o->ys[l] != 0) {
          const double *al = o->a + (size_t)l * n;
          double as = orc_dot(al, sk, n) / o->as[l];                             /* :173 */
          for (int64_t j = 0; j < n; ++j) ak[j] -= as * al[j];                   /* :174 */
        }
      }
      o->as[k] = orc_dot(ak, sk, n);                                             /* :177 */
      if (o->as[k] != 0) {
        double na = orc_nrm2(ak, n);
        o->opnorm_ub += na * na / fabs(o->as[k]);                                /* :179 */
      }
    }
  }
  return 1;
}

/* diag!  src/lsr1.jl:196-211 */
void orc_lsr1_diag(orc_lsr1 *o, double *d) {
  const int64_t n = o->n;
  for (int64_t j = 0; j < n; ++j) d[j] = 1.0;
  if (o->scaling) for (int64_t j = 0; j < n; ++j) d[j] /= o->gamma;
  for (int i = 1; i <= o->mem; ++i) {
    int k = pmod(o->ins0 + i - 1, o->mem);
    if (o->ys[k] != 0.0) {
      const double *ak = o->a + (size_t)k * n;
      for (int64_t j = 0; j < n; ++j) d[j] += ak[j] * ak[j] / o->as[k];          /* :206 */
    }
  }
}

/* reset!  src/lsr1.jl:217-240 */
void orc_lsr1_reset(orc_lsr1 *o) {
  size_t nn = (size_t)o->n * (size_t)o->mem;
  memset(o->s, 0, nn * sizeof(double));
  memset(o->y, 0, nn * sizeof(double));
  memset(o->a, 0, nn * sizeof(double));
  memset(o->ys, 0, o->mem * sizeof(double));
  memset(o->as, 0, o->mem * sizeof(double));
  o->gamma = 1.0;
  o->ins0 = 0;
}

/* ================= diagonal quasi-Newton push!  src/DiagonalHessianApproximation.jl ================= */
int orc_diagqn_push(int kind, double *d, const double *s, const double *y, int64_t n) {
  if (kind == 3) {                                                              /* SpectralGradient :186-196 */
    int allzero = 1;
    for (int64_t i = 0; i < n; ++i) if (s[i] != 0) allzero = 0;
    if (allzero) return -1;
    d[0] = orc_dot(s, y, n) / orc_dot(s, s, n);
    return 0;
  }
  double sNorm = orc_nrm2(s, n);
  if (sNorm == 0) return -1;
  double sNorm2 = sNorm * sNorm;
  if (kind == 2) {                                                              /* DiagonalBFGS :234-248 */
    double sT_y = orc_dot(s, y, n) / sNorm2;
    long double sum = 0;
    for (int64_t i = 0; i < n; ++i) { d[i] = fabs(y[i]); sum += d[i]; }
    double c = (double)sum / sT_y;
    for (int64_t i = 0; i < n; ++i) d[i] *= c;
    return 0;
  }
  long double s4 = 0, s2d = 0;                                                  /* PSB :45-64, Andrei :120-141 */
  for (int64_t i = 0; i < n; ++i) {
    long double s2 = (long double)s[i] * s[i];
    s4 += s2 * s2;
    s2d += s2 * d[i];
  }
  double trA2 = (double)s4 / (sNorm2 * sNorm2);
  double sT_y = orc_dot(s, y, n) / sNorm2;
  double sT_B_s = (double)s2d / sNorm2;
  double q = sT_y - sT_B_s;
  if (kind == 1) q += orc_dot(s, s, n) / sNorm2;
  q /= trA2;
  double c = q / sNorm2;
  if (kind == 0) for (int64_t i = 0; i < n; ++i) d[i] = d[i] + c * (s[i] * s[i]);
  else for (int64_t i = 0; i < n; ++i) d[i] = d[i] + (c * (s[i] * s[i]) - 1.0);
  return 0;
}

/* ================= kron  src/kron.jl:14-40 =================
 * prod : X = reshape(x,q,n);  res = α vec(B X Aᵀ) + β res      (A m×n, B p×q)  -> p×m
 * tprod: X = reshape(x,p,m);  res = α vec(Bᵀ X A) + β res                      -> q×n
 * all matrices column-major. */
void orc_kron(double *res, const double *A, int64_t m, int64_t n, const double *B, int64_t p, int64_t q,
              const double *x, double alpha, double beta, int trans) {
  int64_t r1 = trans ? q : p;   /* rows of result */
  int64_t c1 = trans ? n : m;   /* cols of result */
  int64_t xi = trans ? p : q;   /* rows of X */
  int64_t xj = trans ? m : n;   /* cols of X */
  long double *Y = (long double *)malloc((size_t)(r1 * xj) * sizeof(long double)); /* Y = op(B) X : r1×xj */
  for (int64_t j = 0; j < xj; ++j)
    for (int64_t i = 0; i < r1; ++i) {
      long double acc = 0;
      for (int64_t l = 0; l < xi; ++l) {
        double bil = trans ? B[l + i * p] : B[i + l * p];
        acc += (long double)bil * (long double)x[l + j * xi];
      }
      Y[i + j * r1] = acc;
    }
  for (int64_t j = 0; j < c1; ++j)
    for (int64_t i = 0; i < r1; ++i) {
      long double acc = 0;
      for (int64_t l = 0; l < xj; ++l) {
        /* prod: Z = Y Aᵀ -> Z[i,j] = Σ_l Y[i,l] A[j,l];  tprod: Z = Y A -> Σ_l Y[i,l] A[l,j] */
        double alj = trans ? A[l + j * m] : A[j + l * m];
        acc += Y[i + l * r1] * (long double)alj;
      }
      double z = (double)acc;
      double *r = &res[i + j * r1];
      *r = (beta == 0.0) ? alpha * z : alpha * z + beta * *r;
    }
  free(Y);
}

/* ---- LinearOperator(M) for a dense matrix: src/constructors.jl:25-27 ----
 * prod!   = mul!(res, M, v, α, β)              trans = 0   res[m] = α M v + β res
 * tprod!  = mul!(res, transpose(M), u, α, β)   trans = 1   res[n] = α Mᵀ u + β res   (ctprod! ≡ tprod! for real T)
 * M is column-major m×n with leading dimension lda.  BLAS gemv in the reference (summation order unspecified): the inner
 * product is taken in long double here; `res` is not read when β == 0 (BLAS semantics). */
void orc_gemv(double *res, const double *M, int64_t m, int64_t n, int64_t lda, const double *v, double alpha, double beta,
              int trans) {
  int64_t nout = trans ? n : m, nin = trans ? m : n;
  for (int64_t i = 0; i < nout; ++i) {
    long double acc = 0.0L;
    for (int64_t j = 0; j < nin; ++j) {
      double mij = trans ? M[j + i * lda] : M[i + j * lda];
      acc += (long double)mij * (long double)v[j];
    }
    double t = alpha * (double)acc;
    res[i] = (beta == 0.0) ? t : t + beta * res[i];
  }
}
/* Float32 storage (test/gpu/nvidia.jl:8-15 uses CUDA.rand -> Float32): same, the result rounded to Float32 once */
void orc_gemv_f32(float *res, const float *M, int64_t m, int64_t n, int64_t lda, const float *v, float alpha, float beta,
                  int trans) {
  int64_t nout = trans ? n : m, nin = trans ? m : n;
  for (int64_t i = 0; i < nout; ++i) {
    long double acc = 0.0L;
    for (int64_t j = 0; j < nin; ++j) {
      float mij = trans ? M[j + i * lda] : M[i + j * lda];
      acc += (long double)mij * (long double)v[j];
    }
    double t = (double)alpha * (double)acc;
    res[i] = (float)((beta == 0.0f) ? t : t + (double)beta * (double)res[i]);
  }
}

/* ---- LinearOperator(M) for M::SparseMatrixCSC: src/constructors.jl:25-27 -> SparseArrays' mul! ----
 * SparseArrays is a standard-library dependency (Project.toml:9, compat "1.10", Project.toml:44) and is NOT vendored under
 * /root/reference; its published algorithm (SparseArrays/src/linalg.jl, `_spmatmul!` and `_At_or_Ac_mul_B!`) is restated:
 *   prod!  (trans = 0): C = β C (or 0 when β == 0); for col = 1:n, αxj = v[col]*α; for k in nzrange(A, col): C[rowval[k]] += nzval[k]*αxj
 *   tprod! (trans = 1): for col = 1:n, tmp = Σ_{k in nzrange(A, col)} nzval[k]*u[rowval[k]]; C[col] = α*tmp + β*C[col]
 * colptr1 / rowval1 hold Julia's 1-based values.  Accurate side: the per-output sums are taken in long double (the
 * reference's summation order over a row is the storage order; nothing in its tests pins more than √eps). */
void orc_spmv_csc(double *res, int64_t m, int64_t n, const int64_t *colptr1, const int64_t *rowval1, const double *nzval,
                  const double *v, double alpha, double beta, int trans) {
  if (trans) {
    for (int64_t col = 0; col < n; ++col) {
      long double tmp = 0.0L;
      for (int64_t k = colptr1[col] - 1; k < colptr1[col + 1] - 1; ++k) tmp += (long double)nzval[k] * (long double)v[rowval1[k] - 1];
      double t = alpha * (double)tmp;
      res[col] = (beta == 0.0) ? t : t + beta * res[col];
    }
    return;
  }
  long double *acc = (long double *)calloc((size_t)(m > 0 ? m : 1), sizeof(long double));
  for (int64_t col = 0; col < n; ++col)
    for (int64_t k = colptr1[col] - 1; k < colptr1[col + 1] - 1; ++k)
      acc[rowval1[k] - 1] += (long double)nzval[k] * (long double)v[col];
  for (int64_t i = 0; i < m; ++i) {
    double t = alpha * (double)acc[i];
    res[i] = (beta == 0.0) ? t : t + beta * res[i];
  }
  free(acc);
}
void orc_spmv_csc_f32(float *res, int64_t m, int64_t n, const int64_t *colptr1, const int64_t *rowval1, const float *nzval,
                      const float *v, float alpha, float beta, int trans) {
  int64_t nout = trans ? n : m;
  long double *acc = (long double *)calloc((size_t)(nout > 0 ? nout : 1), sizeof(long double));
  for (int64_t col = 0; col < n; ++col)
    for (int64_t k = colptr1[col] - 1; k < colptr1[col + 1] - 1; ++k) {
      int64_t row = rowval1[k] - 1;
      if (trans) acc[col] += (long double)nzval[k] * (long double)v[row];
      else acc[row] += (long double)nzval[k] * (long double)v[col];
    }
  for (int64_t i = 0; i < nout; ++i) {
    double t = (double)alpha * (double)acc[i];
    res[i] = (float)((beta == 0.0f) ? t : t + (double)beta * (double)res[i]);
  }
  free(acc);
}

uint16_t orc_f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40); /* NaN */
  uint32_t r = 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)((u + r) >> 16);
}
float orc_bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
