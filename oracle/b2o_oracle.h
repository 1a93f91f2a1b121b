/*
 * b2o_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the closure bodies on the `mul!(res, op, v, α, β)`
 * hot path of JuliaSmoothOptimizers/LinearOperators.jl v2.14.2.  Only `tests/`,
 * `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
 * `bench.py` may load this library; the product (libb2o.so) never does.
 *
 * PARITY PINNING: the reference is Julia and cannot be executed in this image.
 * The oracle is pinned (tests/test_oracle_pinning.py) against every literal
 * known answer and every test predicate the reference's own suite holds for this
 * path (exact index equality test/test_linop.jl:437-467; H*B≈I, dense-BFGS /
 * dense-SR1 recurrences test/test_lbfgs.jl:13-159, test/test_lsr1.jl:7-72; dense
 * kron test/test_kron.jl:3-39).  Bit-level parity of *reductions* (dot/norm/sum
 * go to OpenBLAS in the reference, summation order unspecified) is UNPINNED by
 * construction; the oracle takes them in long double so it is the accurate side.
 *
 * Each function cites the reference file:line it follows (paths relative to the
 * reference checkout).
 */
#ifndef B2O_ORACLE_H
#define B2O_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* mode: accurate=1 -> long-double reductions (parity checker);
 *       accurate=0 -> plain double loops, `threads` OpenMP threads (timing twin). */
void orc_set_mode(int accurate, int threads);
int orc_max_threads(void);

/* deterministic counter-based U[lo,hi) generator shared bit-for-bit with the CUDA side */
void orc_fill_uniform(double *x, int64_t n, uint64_t seed, double lo, double hi);
void orc_fill_uniform_f32(float *x, int64_t n, uint64_t seed, float lo, float hi);

double orc_dot(const double *a, const double *b, int64_t n);
double orc_sum(const double *a, int64_t n);
double orc_nrm2(const double *a, int64_t n);

/* leaf closures -- src/special-operators.jl, src/linalg.jl */
void orc_eye(double *res, int64_t nres, const double *v, int64_t nv, double alpha, double beta, int64_t n_min);
void orc_ones(double *res, int64_t nres, const double *v, int64_t nv, double alpha, double beta);
void orc_zeros(double *res, int64_t nres, double alpha, double beta);
void orc_diag_square(double *res, const double *d, const double *v, int64_t n, double alpha, double beta);
void orc_diag_rect(double *res, int64_t nres, const double *d, const double *v, double alpha, double beta, int64_t n_min);
void orc_householder(double *res, const double *h, const double *v, int64_t n, double alpha, double beta);
/* ComplexF64 leaves, vectors as interleaved (re, im); alpha / beta point to 2 doubles */
void orc_cdiag(double *res, const double *d, const double *v, int64_t n, int conj_d, const double *alpha, const double *beta);
void orc_chouseholder(double *res, const double *h, const double *v, int64_t n, const double *alpha, const double *beta);
void orc_conj(double *dst, const double *src, int64_t n);
void orc_restrict(double *res, const int64_t *idx1, int64_t k, const double *v);
void orc_extend(double *res, int64_t ncol, const int64_t *idx1, int64_t k, const double *u);

/* quasi-Newton operators -- src/lbfgs.jl, src/lsr1.jl */
typedef struct orc_lbfgs orc_lbfgs;
orc_lbfgs *orc_lbfgs_create(int64_t n, int mem, int scaling, int damped, double sigma2, double sigma3, int inverse);
void orc_lbfgs_destroy(orc_lbfgs *);
/* Full-size parity runs: an apply-only operator whose columns stay where they are (the GPU) and are handed to the oracle
 * one at a time by `fetch(user, which, k0, slot)` (which: 0 s, 1 y, 2 a, 3 b; slot 0/1 = which of the caller's two host
 * buffers to fill: an apply never needs more than two columns at once).  The SAME apply code runs either way. */
typedef const double *(*orc_fetch_fn)(void *user, int which, int k0, int slot);
orc_lbfgs *orc_lbfgs_create_external(int64_t n, int mem, int scaling, int inverse, orc_fetch_fn fetch, void *user);
/* the inner products of the last apply in the order the reference takes them (forward: a_k.x, b_k.x oldest -> newest;
 * inverse: loop 1 s_k.q newest -> oldest, then loop 2 y_k.q oldest -> newest) */
const double *orc_lbfgs_last_dots(orc_lbfgs *, int *count);
void orc_lbfgs_apply(orc_lbfgs *, double *res, const double *x, double alpha, double beta);
int orc_lbfgs_push(orc_lbfgs *, const double *s, const double *y);                    /* 1 accepted, 0 rejected, <0 error */
int orc_lbfgs_push_damped_fwd(orc_lbfgs *, const double *s, const double *y, double *Bs);
int orc_lbfgs_push_damped_inv(orc_lbfgs *, const double *s, double *y, double alpha, const double *g, double *Bs);
int orc_lbfgs_diag(orc_lbfgs *, double *d);
void orc_lbfgs_reset(orc_lbfgs *);
/* state access: which = 0:s 1:y 2:a 3:b (n-vectors, slot k0 0-based); scalars below */
double *orc_lbfgs_col(orc_lbfgs *, int which, int k0);
double *orc_lbfgs_ys(orc_lbfgs *);
double orc_lbfgs_gamma(orc_lbfgs *);
void orc_lbfgs_set_gamma(orc_lbfgs *, double);
int orc_lbfgs_insert(orc_lbfgs *);                 /* 1-based, as data.insert */
void orc_lbfgs_set_insert(orc_lbfgs *, int insert1);
double orc_lbfgs_opnorm_upper_bound(orc_lbfgs *);
/* solve_shifted_system!(x, B, b, σ): (B + σI) x = b for a forward operator -- src/utilities.jl:207-248; -1 if σ < 0 or inverse */
int orc_lbfgs_solve_shifted(orc_lbfgs *, double *x, const double *b, double sigma);

typedef struct orc_lsr1 orc_lsr1;
orc_lsr1 *orc_lsr1_create(int64_t n, int mem, int scaling);
void orc_lsr1_destroy(orc_lsr1 *);
void orc_lsr1_apply(orc_lsr1 *, double *res, const double *x, double alpha, double beta);
int orc_lsr1_push(orc_lsr1 *, const double *s, const double *y);
void orc_lsr1_diag(orc_lsr1 *, double *d);
void orc_lsr1_reset(orc_lsr1 *);
double *orc_lsr1_col(orc_lsr1 *, int which, int k0); /* 0:s 1:y 2:a */
double *orc_lsr1_ys(orc_lsr1 *);
double *orc_lsr1_as(orc_lsr1 *);
double orc_lsr1_gamma(orc_lsr1 *);
void orc_lsr1_set_gamma(orc_lsr1 *, double);
int orc_lsr1_insert(orc_lsr1 *);
void orc_lsr1_set_insert(orc_lsr1 *, int insert1);
double orc_lsr1_opnorm_upper_bound(orc_lsr1 *);

/* diagonal quasi-Newton push! -- src/DiagonalHessianApproximation.jl:45-64,120-141,186-196,234-248; kind 0 PSB 1 Andrei 2 BFGS 3 Spectral
 * returns 0, or -1 when s == 0 (the reference errors) */
int orc_diagqn_push(int kind, double *d, const double *s, const double *y, int64_t n);

/* kron(A,B)*x = alpha*vec(B X A^T)+beta*res -- src/kron.jl:14-22; A m×n, B p×q, col-major;
 * trans: 0 prod, 1 tprod (B^T X A), 2 ctprod (same for real) */
void orc_kron(double *res, const double *A, int64_t m, int64_t n, const double *B, int64_t p, int64_t q,
              const double *x, double alpha, double beta, int trans);

/* LinearOperator(M), dense column-major M (m×n, leading dimension lda): mul!(res, M, v, α, β) / transpose(M) --
 * src/constructors.jl:25-27.  trans: 0 prod!, 1 tprod!/ctprod! */
void orc_gemv(double *res, const double *M, int64_t m, int64_t n, int64_t lda, const double *v, double alpha, double beta,
              int trans);
void orc_gemv_f32(float *res, const float *M, int64_t m, int64_t n, int64_t lda, const float *v, float alpha, float beta,
                  int trans);

/* LinearOperator(M::SparseMatrixCSC): SparseArrays' mul!(res, M, v, α, β) / transpose(M) restated (stdlib dependency,
 * Project.toml:9,44; not vendored).  1-based colptr / rowval as in Julia.  trans: 0 prod!, 1 tprod!/ctprod! */
void orc_spmv_csc(double *res, int64_t m, int64_t n, const int64_t *colptr1, const int64_t *rowval1, const double *nzval,
                  const double *v, double alpha, double beta, int trans);
void orc_spmv_csc_f32(float *res, int64_t m, int64_t n, const int64_t *colptr1, const int64_t *rowval1, const float *nzval,
                      const float *v, float alpha, float beta, int trans);

/* bf16 helpers used by the kron parity test (round-to-nearest-even) */
uint16_t orc_f32_to_bf16(float f);
float orc_bf16_to_f32(uint16_t h);

#ifdef __cplusplus
}
#endif
#endif
