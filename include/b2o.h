/*
 * b2o.h -- C ABI of libb2o.so: B200 (sm_100a) matrix-free operator-apply engine.
 *
 * This is the drop-in boundary for the `mul!(res, op, v, α, β)` hot path of
 * JuliaSmoothOptimizers/LinearOperators.jl v2.14.2.  The reference has NO FFI: its
 * boundary is Julia dispatch -- an `AbstractLinearOperator` whose `prod!/tprod!/ctprod!`
 * closures (src/abstract.jl:46-59) are invoked by the generic `mul!`
 * (src/operations.jl:22-32).  Each entry point below is what such a closure body
 * would `ccall`; the comment on each names the reference closure it replaces.
 * INTEGRATION.md shows the Julia-side binding (and the ctypes one the tests use).
 *
 * Conventions
 *  - plain C: raw device pointers, sizes, scalars by value.  No torch / CUDA types
 *    (a `cudaStream_t` is passed as `void*`).
 *  - every function returns a b2o_status; never throws/aborts.  The message for the
 *    last failure on the calling thread is returned by b2o_last_error().
 *  - `res_len` / `v_len` are the lengths of the caller's vectors: a mismatch with the
 *    operator shape returns B2O_ESHAPE("shape mismatch") *before* any work, like
 *    src/operations.jl:23-24.
 *  - when beta == 0 `res` is never read (src/constructors.jl:63-66): it may be
 *    uninitialised memory (NaNs do not propagate).
 *  - calls enqueue on the context's stream and return; they synchronise only when a
 *    host-visible scalar is part of the reference semantics (push! acceptance tests).
 *  - one in-flight call per context (the reference's operators share scratch the same
 *    way: src/abstract.jl:151-153, src/lbfgs.jl:127).
 *  - no allocation after handle creation (reference contract: test/test_lbfgs.jl:199-217).
 *  - row-partitioned mode: after b2o_comm_init() every vector argument is the calling
 *    rank's contiguous row slab and every inner product is all-reduced over ranks.
 */
#ifndef B2O_H
#define B2O_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B2O_OK = 0,
  B2O_ESHAPE = 1,       /* LinearOperatorException("shape mismatch")            */
  B2O_EARG = 2,         /* bad argument (null pointer, index out of range, ...) */
  B2O_ECUDA = 3,        /* CUDA runtime error                                   */
  B2O_ENCCL = 4,        /* NCCL error                                           */
  B2O_ESTATE = 5,       /* wrong push! variant for this operator (ErrorException, src/lbfgs.jl:296-298,332-334) */
  B2O_ENOMEM = 6,
  B2O_EUNSUPPORTED = 7  /* dtype / feature not built                             */
} b2o_status;

typedef enum { B2O_F64 = 0, B2O_F32 = 1, B2O_BF16 = 2 } b2o_dtype;

typedef struct b2o_ctx_s b2o_ctx;
typedef struct b2o_qn_s b2o_qn;       /* LBFGSOperator / InverseLBFGSOperator / LSR1Operator state */
typedef struct b2o_index_s b2o_index; /* opRestriction / opExtension index set */
typedef struct b2o_kron_s b2o_kron;   /* kron(A,B) operator (tcgen05 GEMM pair) */
typedef struct b2o_graph_s b2o_graph; /* static operator tree lowered to one fused launch */
typedef struct b2o_dense_s b2o_dense; /* LinearOperator(M) for a dense device matrix */
typedef struct b2o_sparse_s b2o_sparse; /* LinearOperator(M) for a sparse matrix (CSC / CSR) */

/* ---- library / context ------------------------------------------------------------------ */
int b2o_version(void);
const char *b2o_last_error(void);
/* device: CUDA ordinal; stream: cudaStream_t to enqueue on (NULL = the legacy default stream). */
int b2o_ctx_create(int device, void *stream, b2o_ctx **out);
int b2o_ctx_destroy(b2o_ctx *ctx);
int b2o_ctx_sync(b2o_ctx *ctx);
/* tuning knobs of the streaming kernels: "tile_rows" (1024|2048|4096), "stages", "grid", "threads";
 * fused trees: "graph_jit" (0|1|2, see b2o_graph_variant), "graph_blocks" (1..4), "graph_interp" (0|2); host-buffer pipeline: "host_chunks";
 * matrix leaves: "dense_scalar" (force the unvectorised dense kernels), "sparse_kernel" (0|3 pipelined rows -- default,
 * 1 plain rows, 2 TMA-staged tiles), "sparse_lanes" (-1 auto | 0..5: 2^k lanes per row);
 * index sets: "extend_form" (0 gather form through the inverse map when the set is dense enough, 1 always memset + scatter);
 * matrix right-hand sides of the two-loop inverse: "twoloop_block" (1 block recursion, 4 / 8 columns per sweep -- default; 0 column loop);
 * block apply with 5..8 right-hand sides: "multi_mma" (0 the SIMT kernel -- default; 1 the FP64 tensor-core kernel, mma.sync m8n8k4:
 * same results to a few ulp, measured slower -- 11.3 vs 10.5 ms at n = 1e8, m = 10, 8 right-hand sides);
 * memory system: "l2_fetch_granularity" (32|64|128: cudaLimitMaxL2FetchGranularity, device-wide);
 * multi-GPU: "use_mailbox" (0|1: NCCL or the in-kernel NVLink mailbox for the inner products of a connected context -- every
 * rank switches at the same point), "numa_local_host" (0|1: b2o_host_alloc prefers the GPU's NUMA node, default 1).
 * Options change which kernel runs, never the result beyond summation order (index work stays exact). */
int b2o_ctx_set_option(b2o_ctx *ctx, const char *key, int64_t value);
/* debug: raw read of workspace scalars (e.g. the kron kernel's %globaltimer timeline with option "kron_debug") */
int b2o_ctx_debug_read(b2o_ctx *ctx, int offset, int count, double *out);
/* number of libb2o kernels launched through this context since creation */
int b2o_ctx_launch_count(b2o_ctx *ctx, int64_t *out);
/* average device time (ms) of the dominant streaming kernel over the launches since the last reset,
 * measured with CUDA events on the context stream; reset=1 clears the accumulator. */
int b2o_ctx_kernel_time(b2o_ctx *ctx, int reset, double *ms_total, int64_t *launches);

/* device-memory helpers so a host without a CUDA binding (C, the ctypes tests) can drive the ABI.
 * A Julia caller passes CuArray pointers instead and never needs these. */
int b2o_malloc(b2o_ctx *ctx, size_t bytes, void **dptr);
int b2o_free(b2o_ctx *ctx, void *dptr);
int b2o_host_alloc(b2o_ctx *ctx, size_t bytes, void **hptr); /* pinned; placed on the GPU's own NUMA node when known */
int b2o_ctx_numa_node(b2o_ctx *ctx, int *node);              /* NUMA node of the context's GPU (-1 = unknown) */
int b2o_host_free(b2o_ctx *ctx, void *hptr);
int b2o_memcpy_h2d(b2o_ctx *ctx, void *dst, const void *src, size_t bytes); /* stream-ordered + sync */
int b2o_memcpy_d2h(b2o_ctx *ctx, void *dst, const void *src, size_t bytes);
int b2o_memset_zero(b2o_ctx *ctx, void *dptr, size_t bytes);
/* synthetic data: x[i] = lo + (hi-lo)*u(seed,i), the counter-based generator shared bit-for-bit
 * with oracle/b2o_oracle.c:orc_fill_uniform (bench + parity tests at sizes too big to copy). */
int b2o_fill_uniform(b2o_ctx *ctx, int dtype, void *dptr, int64_t n, uint64_t seed, double lo, double hi);
/* dot(a,b) on device, result to host (used by the tests' size-independent properties) */
int b2o_dot(b2o_ctx *ctx, int dtype, const void *a, const void *b, int64_t n, double *out);

/* ---- leaf operators (src/special-operators.jl, src/linalg.jl) ---------------------------- */
/* mulSquareOpDiagonal! :125-131 and mulOpDiagonal! :144-151.  res[0:nmin) = (alpha*d)*v (+beta*res);
 * res[nmin:nrow) = 0 regardless of beta.  `d` is borrowed per call (the reference aliases it). */
int b2o_diag_apply(b2o_ctx *ctx, int dtype, int64_t nrow, int64_t ncol, const void *d, int64_t d_len,
                   void *res, int64_t res_len, const void *v, int64_t v_len, double alpha, double beta);
/* mulOpEye! :36-44 (rectangular tail is 0 when beta==0, else the scalar beta). */
int b2o_eye_apply(b2o_ctx *ctx, int dtype, int64_t nrow, int64_t ncol, void *res, int64_t res_len,
                  const void *v, int64_t v_len, double alpha, double beta);
/* mulOpOnes! :79-85  res .= alpha*sum(v) (+beta*res) */
int b2o_ones_apply(b2o_ctx *ctx, int dtype, int64_t nrow, int64_t ncol, void *res, int64_t res_len,
                   const void *v, int64_t v_len, double alpha, double beta);
/* mulOpZeros! :102-108 */
int b2o_zeros_apply(b2o_ctx *ctx, int dtype, int64_t nrow, int64_t ncol, void *res, int64_t res_len,
                    int64_t v_len, double alpha, double beta);
/* mulHouseholder! src/linalg.jl:77-83  res = alpha*(v - 2*dot(h,v)*h) (+beta*res); one launch */
int b2o_householder_apply(b2o_ctx *ctx, int dtype, int64_t n, const void *h, void *res, int64_t res_len,
                          const void *v, int64_t v_len, double alpha, double beta);
/* ---- ComplexF64 leaves (vectors = interleaved (re, im) pairs, 16-byte aligned).  The reference is generic in the element type:
 * for complex operators the adjoint / transpose / conjugate wrappers run the conj-sandwich  conj!(res); prod!(res, conj.(v),
 * conj(α), conj(β)); conj!(res)  (src/adjtrans.jl:128-136, 196-204), opDiagonal's ctprod! uses conj.(d)
 * (src/special-operators.jl:140) and mulHouseholder!'s dot(h, v) conjugates h (src/linalg.jl:79). */
int b2o_cdiag_apply(b2o_ctx *ctx, int64_t nrow, int64_t ncol, const void *d, int64_t d_len, int conj_d, void *res, int64_t res_len,
                    const void *v, int64_t v_len, double alpha_re, double alpha_im, double beta_re, double beta_im);
int b2o_ceye_apply(b2o_ctx *ctx, int64_t nrow, int64_t ncol, void *res, int64_t res_len, const void *v, int64_t v_len,
                   double alpha_re, double alpha_im, double beta_re, double beta_im);
int b2o_czeros_apply(b2o_ctx *ctx, int64_t nrow, int64_t ncol, void *res, int64_t res_len, int64_t v_len, double beta_re,
                     double beta_im);
int b2o_chouseholder_apply(b2o_ctx *ctx, int64_t n, const void *h, void *res, int64_t res_len, const void *v, int64_t v_len,
                           double alpha_re, double alpha_im, double beta_re, double beta_im);
/* dst = conj(src), n ComplexF64 elements; dst == src is conj!(res) */
int b2o_conj(b2o_ctx *ctx, void *dst, const void *src, int64_t n);

/* opRestriction ctor :187-201: idx1 = HOST array of k 1-based indices; out-of-range -> B2O_EARG
 * ("indices should be between 1 and ncol").  Duplicates are legal. */
int b2o_index_create(b2o_ctx *ctx, const int64_t *idx1, int64_t k, int64_t ncol, b2o_index **out);
int b2o_index_destroy(b2o_index *ix);
/* mulRestrict! :167-169  res .= v[I]  (alpha, beta ignored by the reference) */
int b2o_restrict_apply(b2o_index *ix, int dtype, void *res, int64_t res_len, const void *v, int64_t v_len);
/* multRestrict! :171-174  res .= 0; res[I] = u  (duplicates: last occurrence wins) */
int b2o_extend_apply(b2o_index *ix, int dtype, void *res, int64_t res_len, const void *u, int64_t u_len);

/* diagonal quasi-Newton updates (src/DiagonalHessianApproximation.jl; their apply is b2o_diag_apply on `d`):
 * push!(B,s,y) for kind 0 DiagonalPSB :45-64, 1 DiagonalAndrei :120-141, 2 DiagonalBFGS :234-248,
 * 3 SpectralGradient :186-196 (d = 1-element device vector holding σ).  s == 0 -> B2O_ESTATE like the reference's error(). */
int b2o_diagqn_push(b2o_ctx *ctx, int kind, void *d, int64_t d_len, const void *s, const void *y, int64_t n);

/* ---- quasi-Newton operators (src/lbfgs.jl, src/lsr1.jl) ---------------------------------- */
/* dtype: B2O_F64, or B2O_F32 (T = Float32, test/test_lbfgs.jl:162-178, test/test_lsr1.jl:74-86): state columns, x and res are then
 * Float32 (4-byte aligned), every elementwise statement runs in Float32, inner products are accumulated in double and rounded to
 * Float32.  Float32 handles support every entry point below (row partitions included; apply_multi on the Float32 block kernel -- two-loop inverse handles column by column --,
 * apply_host through one staged copy) except solve_shifted and the compact modes, which return B2O_EUNSUPPORTED for them. */
/* LBFGSOperator(T,n;mem,scaling,damped,σ₂,σ₃) :168-208 (inverse=0) / InverseLBFGSOperator :112-160 (inverse=1) */
int b2o_lbfgs_create(b2o_ctx *ctx, int dtype, int64_t n, int mem, int scaling, int damped, double sigma2,
                     double sigma3, int inverse, b2o_qn **out);
/* LSR1Operator(T,n;mem,scaling) src/lsr1.jl:86-113 */
int b2o_lsr1_create(b2o_ctx *ctx, int dtype, int64_t n, int mem, int scaling, b2o_qn **out);
int b2o_qn_destroy(b2o_qn *op);
/* prod! closure: lbfgs_multiply :117-154 / :173-202, lsr1_multiply src/lsr1.jl:89-107.  One persistent launch. */
int b2o_qn_apply(b2o_qn *op, void *res, int64_t res_len, const void *x, int64_t x_len, double alpha, double beta);
/* same, HOST buffers: H2D copy of x, apply, D2H copy of res, stream sync (the end-to-end path). */
int b2o_qn_apply_host(b2o_qn *op, void *res_host, const void *x_host, int64_t len, double alpha, double beta);
/* mul!(Res::Matrix, op, X::Matrix, α, β) (src/operations.jl:34-36): Res, X are column-major n x nrhs
 * device matrices with leading dimensions ldr, ldx (>= n).  Extension: every column of Res equals b2o_qn_apply of the matching
 * column of X (to reduction-order rounding), but each state column is streamed once per 8 right-hand sides:
 * (2*ncols + 3*nrhs)*8*n algorithmic bytes per pass.  Two-loop inverse handles run the BLOCK two-loop recursion: the update and
 * dot columns of every sweep are staged once for up to 8 right-hand sides ((4*nrhs + 4)*A + nrhs vector passes instead of
 * nrhs*(8*A + 1)), results bit-identical to b2o_qn_apply per column; its nrhs work vectors are allocated on the first call.
 * Row-partitioned two-loop handles and NCCL (non-mailbox) partitions run column by column. */
int b2o_qn_apply_multi(b2o_qn *op, void *res, int64_t ldr, const void *x, int64_t ldx, int64_t len, int nrhs, double alpha,
                       double beta);
/* push!(op,s,y) src/lbfgs.jl:269-287, src/lsr1.jl:119-184.  *accepted = 0 when the pair is rejected
 * (not an error, the reference returns op silently).  Synchronises (host-side acceptance tests). */
int b2o_qn_push(b2o_qn *op, const void *s, const void *y, int64_t len, int *accepted);
/* push!(op,s,y,Bs) src/lbfgs.jl:289-321 (forward damped; Bs is caller scratch of len n) */
int b2o_lbfgs_push_damped_fwd(b2o_qn *op, const void *s, const void *y, void *Bs, int64_t len, int *accepted);
/* push!(op,s,y,α,g,Bs) src/lbfgs.jl:323-357 (inverse damped; y is overwritten with the damped y) */
int b2o_lbfgs_push_damped_inv(b2o_qn *op, const void *s, void *y, double alpha, const void *g, void *Bs,
                              int64_t len, int *accepted);
/* diag! src/lbfgs.jl:379-395 (forward only: inverse -> B2O_ESTATE), src/lsr1.jl:196-211 */
int b2o_qn_diag(b2o_qn *op, void *d, int64_t d_len);
/* reset! src/lbfgs.jl:401-427, src/lsr1.jl:217-240 */
int b2o_qn_reset(b2o_qn *op);
/* solve_shifted_system!(x, B, b, σ) src/utilities.jl:207-248 and ldiv!(x, B, b) :281-289 (σ = 0): (B + σI) x = b for a FORWARD
 * L-BFGS operator; σ < 0 -> B2O_EARG ("σ must be nonnegative", the reference's ArgumentError).  The n x 2mem work matrix
 * (data.shifted_p, src/lbfgs.jl:53) is allocated on first use. */
int b2o_lbfgs_solve_shifted(b2o_qn *op, void *x, int64_t x_len, const void *b, int64_t b_len, double sigma);
/* state access (checkpoint/resume and parity tests).  which: 0=s 1=y 2=a 3=b; k0 = 0-based ring slot.
 * scalars: insert1 (1-based data.insert), gamma (scaling_factor), opnorm_upper_bound,
 * ys[mem], aux[mem] (LBFGS inverse: α scratch; LBFGS forward: norm_b; LSR1: as). */
int b2o_qn_get_col(b2o_qn *op, int which, int k0, void *dst_device);
int b2o_qn_set_col(b2o_qn *op, int which, int k0, const void *src_device);
int b2o_qn_get_scalars(b2o_qn *op, int *insert1, double *gamma, double *opnorm_ub, double *ys, double *aux);
int b2o_qn_set_scalars(b2o_qn *op, int insert1, double gamma, double opnorm_ub, const double *ys, const double *aux);
/* options: "inverse_mode" (InverseLBFGSOperator only): 0 = the reference's two-loop recursion (default),
 * 1 = compact representation H = γI + [S γY] W [Sᵀ; γYᵀ] (Byrd-Nocedal-Schnabel): same operator, (4m+3)n instead of (8m+2)n
 * words of DRAM traffic and ONE all-reduce instead of 2m dependent ones; rounding differs from the two-loop.
 * "forward_mode" (forward LBFGSOperator): 1 = compact form (push! = O(m) dots).  "push_mode" (forward LBFGSOperator, reference
 * form): 1 (default) = each a_k of the rebuild in push! (src/lbfgs.jl:236-250) is one launch of the streaming apply kernel,
 * 0 = generic multi-dot + linear-combination passes (same statements, same rounding of the elementwise part). */
int b2o_qn_set_option(b2o_qn *op, const char *key, int64_t value);
/* algorithmic DRAM bytes of one apply (SURVEY §8d / DESIGN.md), for the roofline */
int b2o_qn_apply_bytes(b2o_qn *op, double beta, double *bytes);

/* ---- fused static operator trees (src/operations.jl:100-234 closure trees collapsed into one launch) ---- */
/* Build the tree bottom-up, compile once, apply many times.  All nodes are n x n.  Leaf vectors are aliased.
 * leaf kinds : 0 opDiagonal(d)  1 opEye  2 opZeros  3 opOnes  4 opHouseholder(h)
 * unary kinds: 12 op*x (scalar x)  13 -op  14 transpose(op)        binary kinds: 10 op1+op2  11 op1*op2
 * The lowering follows mul!'s own recursion (prod_op! :117-128, sum_prod! :187-197, x*α folding :163-177), so the
 * result equals the closure tree's, with every dot/sum taken once and no temporaries in HBM. */
int b2o_graph_create(b2o_ctx *ctx, int64_t n, b2o_graph **out);
int b2o_graph_destroy(b2o_graph *g);
int b2o_graph_leaf(b2o_graph *g, int kind, const void *vec, int *node);
int b2o_graph_unary(b2o_graph *g, int kind, int child, double x, int *node);
int b2o_graph_binary(b2o_graph *g, int kind, int a, int b, int *node);
int b2o_graph_compile(b2o_graph *g, int root);
/* mul!(res, op, v, alpha, beta) (transposed != 0: transpose(op)) in ONE cooperative launch */
int b2o_graph_apply(b2o_graph *g, int transposed, void *res, int64_t res_len, const void *v, int64_t v_len,
                    double alpha, double beta);
/* The compiled passes run either as an NVRTC-specialised instantiation of the pass template (straight-line code,
 * default when libnvrtc + the driver are present) or through the built-in interpreter kernel (ctx option
 * "graph_jit"=0).  b2o_graph_create accepts ctx == NULL for a dry graph that can be lowered and NVRTC-compiled on
 * a CPU-only box (b2o_graph_jit_check) but not applied. */
int b2o_graph_jit_source(b2o_graph *g, int transposed, double beta, char *buf, int64_t cap, int64_t *len);
int b2o_graph_jit_check(b2o_graph *g, int64_t *cubin_bytes);
int b2o_graph_uses_jit(b2o_graph *g, int transposed, double beta, int *out);
/* executor of the last apply of a variant: 0 interpreter kernel, 1 NVRTC-specialised, 2 ahead-of-time instantiation (the common
 * chains -- config 3 and its variants, H*D, D1*D2, D + c*I, H1*H2, H*D*H -- are compiled into the library, so they need no
 * libnvrtc at run time); source_hash = FNV-1a of the variant's generated source, the key of that table.  Context option
 * "graph_jit": 0 interpreter, 1 (default) table -> NVRTC -> interpreter, 2 table -> interpreter. */
int b2o_graph_variant(b2o_graph *g, int transposed, double beta, int *executor, uint64_t *source_hash);
/* passes, reductions and algorithmic DRAM bytes of one apply */
int b2o_graph_info(b2o_graph *g, int transposed, double beta, int *npasses, int *nreductions, double *alg_bytes);

/* ---- kron(A, B) (src/kron.jl:10-49): the tensor-core path ----------------------------------- */
/* A: m x n, B: p x q, COLUMN-major (Julia layout), bf16, 16-byte aligned, every dimension a multiple of 8.  The
 * matrices are borrowed (aliased) for the lifetime of the handle; row-major copies for the K-major operands are
 * made once here.  max_batch = most right-hand sides one apply may carry (sizes the L2-resident intermediate). */
int b2o_kron_create(b2o_ctx *ctx, int dtype, const void *A, int64_t m, int64_t n, const void *B, int64_t p, int64_t q,
                    int max_batch, b2o_kron **out);
int b2o_kron_destroy(b2o_kron *k);
/* trans = 0: prod!  res = alpha*vec(B X A^T) + beta*res, X = reshape(x, q, n)   (src/kron.jl:14-22)
 * trans = 1: tprod!/ctprod!  res = alpha*vec(B^T X A) + beta*res, X = reshape(x, p, m)   (:23-40)
 * x / res hold nb vectors back to back (nb = 1 is the reference call); *_len are the per-vector lengths.
 * res_dtype: B2O_BF16 (the reference's promoted element type; the final rounding alone is 2^-9 relative) or B2O_F32.
 * One plain clustered launch (a cluster of CTAs per 128-row block, no grid-wide dependency): multicast TMA -> tcgen05.mma (fp32
 * accumulate in TMEM) -> bf16 hi/lo intermediate (TMA store, stays in L2, published cluster-wide through an mbarrier) ->
 * tcgen05.mma -> TMA store of the result.  With many right-hand sides (nb * ceil(M/256) >= a third of the SM count) the launch is the
 * cta_group::2 kernel instead: CTA pairs, 256 x 256 pair tiles, 256-row units, Y handed over inside each CTA. */
int b2o_kron_apply(b2o_kron *k, int trans, void *res, int res_dtype, int64_t res_len, const void *x, int64_t x_len, int nb,
                   double alpha, double beta);
int b2o_kron_flops(b2o_kron *k, int nb, double *flops);
/* tuning overrides: "cluster" (0 auto | 1, 2, 4, 8, 16 CTAs per unit), "tile_m" (0 auto | 64, 128 rows per unit | 256 = force the
 * cta_group::2 pair kernel), "tile_n" (0 auto | 32, 64, 128 columns per CTA tile), "pair_tma_stores" (pair kernel epilogue:
 * 1 TMA stores -- default, 0 plain stores) */
int b2o_kron_set_option(b2o_kron *k, const char *key, int64_t value);
/* measurement aid: an EMPTY kernel launched with a kron configuration's grid, cluster size and dynamic shared memory, so the launch
 * overhead inside an event-timed kron figure can be stated (honours the ctx option "time_kernels") */
int b2o_kron_launch_floor(b2o_ctx *ctx, int grid, int cluster, int64_t smem_bytes);

/* ---- LinearOperator(M), dense matrix leaf (src/constructors.jl:15-29) --------------------------- */
/* M: COLUMN-major (Julia layout) m x n device matrix, leading dimension lda >= max(1,m), dtype B2O_F64 or B2O_F32, aligned to
 * its element size (16-byte alignment + lda a multiple of 16/sizeof(T) selects the vectorised kernels).  The matrix is borrowed
 * (aliased) for the lifetime of the handle, like the reference's closures capture M; the handle owns only the partial-sum
 * workspace of split products.  This is what BlockDiagonalOperator(A, B, C) of CUDA matrices runs (test/gpu/nvidia.jl:8-15). */
int b2o_dense_create(b2o_ctx *ctx, int dtype, const void *M, int64_t m, int64_t n, int64_t lda, b2o_dense **out);
int b2o_dense_destroy(b2o_dense *d);
/* trans = 0: prod!  mul!(res, M, v, α, β) :25;  trans = 1: tprod!/ctprod!  mul!(res, transpose(M), u, α, β) :26-27 (real element
 * types: adjoint(M) == transpose(M)).  res = α (M v) + β res, res never read when β == 0; vectors have the matrix's dtype. */
int b2o_dense_apply(b2o_dense *d, int trans, void *res, int64_t res_len, const void *v, int64_t v_len, double alpha, double beta);
/* algorithmic DRAM bytes of one product (matrix once + vectors), for the roofline */
int b2o_dense_apply_bytes(b2o_dense *d, int trans, double beta, double *bytes);

/* ---- LinearOperator(M), sparse matrix leaf (src/constructors.jl:15-29 with M::SparseMatrixCSC) --------------- */
/* fmt 0: CSC, Julia's SparseMatrixCSC fields colptr[n+1], rowval[nnz], nzval[nnz]; fmt 1: CSR (rowptr[m+1], colval[nnz]).
 * ptr1 / idx1: HOST arrays holding the reference's 1-based values (validated: B2O_EARG when out of range / not monotone);
 * vals: DEVICE array, B2O_F64 or B2O_F32, borrowed (aliased) for the lifetime of the handle.  The structure is transposed once
 * here (host, stable counting sort) so that both products are gather-form, atomics-free and bit-reproducible. */
int b2o_sparse_create(b2o_ctx *ctx, int dtype, int fmt, int64_t m, int64_t n, int64_t nnz, const int64_t *ptr1,
                      const int64_t *idx1, const void *vals, b2o_sparse **out);
int b2o_sparse_destroy(b2o_sparse *s);
/* trans = 0: prod!  mul!(res, M, v, α, β);  trans = 1: tprod!/ctprod!  mul!(res, transpose(M), u, α, β) (real element types).
 * res is never read when β == 0. */
int b2o_sparse_apply(b2o_sparse *s, int trans, void *res, int64_t res_len, const void *v, int64_t v_len, double alpha, double beta);
/* after nzval was changed in place: re-gather the values of the transposed copy (the given orientation is aliased) */
int b2o_sparse_refresh(b2o_sparse *s);
/* algorithmic DRAM bytes of one product: nnz*(sizeof(T)+4) + 8*(rows+1) + the vectors */
int b2o_sparse_apply_bytes(b2o_sparse *s, int trans, double beta, double *bytes);

/* ---- row-partitioned multi-GPU (one process per GPU) ------------------------------------- */
/* id: 128-byte ncclUniqueId produced on rank 0 by b2o_comm_unique_id and broadcast by the host
 * (torch.distributed / MPI / files).  After init every inner product of this context is all-reduced. */
int b2o_comm_unique_id(void *id128);
int b2o_comm_init(b2o_ctx *ctx, const void *id128, int nranks, int rank);
int b2o_comm_destroy(b2o_ctx *ctx);
/* NVLink peer mailbox: with it the quasi-Newton applies stay ONE persistent launch per GPU -- the dots are all-reduced
 * inside the kernel through peer-mapped memory (stores over NVLink + system-scope flags) instead of a kernel/NCCL pair per
 * inner product.  Each rank exports a 64-byte CUDA-IPC handle of its mailbox, the host all-gathers them (rank order),
 * every rank connects.  NCCL (b2o_comm_init) stays in use for the reductions inside push!. */
int b2o_mbox_local_handle(b2o_ctx *ctx, void *handle64);
int b2o_mbox_connect(b2o_ctx *ctx, const void *handles, int nranks, int rank);
int b2o_mbox_disconnect(b2o_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif
