// b2o_dense.cu -- LinearOperator(M) for a dense device matrix (src/constructors.jl:15-29).
//
// The reference's three closures are `mul!(res, M, v, α, β)`, `mul!(res, transpose(M), u, α, β)` and
// `mul!(res, adjoint(M), w, α, β)`: BLAS-2 gemv (LinearAlgebra -> OpenBLAS; cuBLAS for a CuMatrix, which is what
// test/gpu/nvidia.jl:8-15 exercises through BlockDiagonalOperator).  Matrix-vector products read every matrix element
// exactly once, so both kernels are HBM-bound: m*n*sizeof(T) algorithmic bytes (+ the two vectors).
//
//   dense_n_kernel : res[m] = α M v + β res.  M is COLUMN-major (Julia layout), so consecutive threads own consecutive
//                    rows (16-byte loads along a column: 2 doubles / 4 floats per thread) and walk the columns with 8
//                    independent loads in flight; grid.y splits the columns so that short-and-wide matrices still fill
//                    the machine.
//   dense_t_kernel : res[n] = α Mᵀ u + β res.  A CTA owns 8 consecutive columns and a slab of rows: the u values are
//                    loaded once per row and reused for the 8 columns (9 independent 16-byte loads in flight per
//                    thread), then one fixed-order block reduction per column; grid.y splits the rows for tall-and-skinny
//                    matrices.
//   dense_finish_kernel : when a product was split, the per-split partial sums are added in split order (deterministic,
//                    no atomics) and the α/β epilogue is applied.  With one split the epilogue happens in the first kernel.
//
// Sums are accumulated in double for both element types (Float32 results are rounded once, at the end).  When β == 0
// `res` is never read (BLAS semantics, src/constructors.jl:63-66).
#include "b2o_internal.cuh"
#include "b2o_dense_kernels.cuh"

struct b2o_dense_s {
  b2o_ctx *ctx = nullptr;
  int dtype = B2O_F64;
  int64_t m = 0, n = 0, lda = 0;
  const void *M = nullptr;       // borrowed (aliased like the reference's closure captures M)
  double *part = nullptr;        // partial sums of split products
  size_t part_elems = 0;
};

static inline size_t dense_elem(int dtype) { return dtype == B2O_F64 ? 8 : 4; }

extern "C" int b2o_dense_create(b2o_ctx *ctx, int dtype, const void *M, int64_t m, int64_t n, int64_t lda, b2o_dense **out) {
  if (!ctx || !out) B2O_FAIL(B2O_EARG, "null argument");
  if (dtype != B2O_F64 && dtype != B2O_F32) B2O_FAIL(B2O_EUNSUPPORTED, "dense: dtype %d not supported (Float64, Float32)", dtype);
  if (m < 0 || n < 0) B2O_FAIL(B2O_EARG, "negative size");
  if (lda < std::max<int64_t>(1, m)) B2O_FAIL(B2O_EARG, "dense: leading dimension %lld < max(1, nrow)", (long long)lda);
  if (m > 0 && n > 0 && !M) B2O_FAIL(B2O_EARG, "null matrix");
  if ((uintptr_t)M % dense_elem(dtype)) B2O_FAIL(B2O_EARG, "dense: matrix not aligned to its element size");
  B2O_CUDA(cudaSetDevice(ctx->device));
  b2o_dense *d = new b2o_dense_s();
  d->ctx = ctx;
  d->dtype = dtype;
  d->m = m;
  d->n = n;
  d->lda = lda;
  d->M = M;
  const size_t need = dense_workspace_elems(ctx->num_sms, (int)(16 / dense_elem(dtype)), m, n);
  if (need) {
    if (cudaMalloc(&d->part, need * sizeof(double)) != cudaSuccess) {
      cudaGetLastError();
      delete d;
      B2O_FAIL(B2O_ENOMEM, "dense: workspace allocation failed");
    }
    d->part_elems = need;
  }
  *out = d;
  return B2O_OK;
}

extern "C" int b2o_dense_destroy(b2o_dense *d) {
  if (!d) return B2O_OK;
  cudaSetDevice(d->ctx->device);
  cudaStreamSynchronize(d->ctx->stream);
  cudaFree(d->part);
  delete d;
  return B2O_OK;
}

template <typename T>
static int dense_run(b2o_dense *d, int trans, void *res, const void *v, double alpha, double beta) {
  b2o_ctx *c = d->ctx;
  return dense_run_impl<T>(c->num_sms, c->stream, &c->launches, d->M, d->m, d->n, d->lda, d->part, d->part_elems, trans, res, v,
                           alpha, beta, c->dense_scalar);
}

// trans = 0: prod!  mul!(res, M, v, α, β);  trans != 0: tprod!/ctprod!  mul!(res, transpose(M), u, α, β) (real T: adjoint ≡ transpose)
extern "C" int b2o_dense_apply(b2o_dense *d, int trans, void *res, int64_t res_len, const void *v, int64_t v_len, double alpha,
                               double beta) {
  if (!d) B2O_FAIL(B2O_EARG, "null operator");
  // a matrix leaf multiplies with the WHOLE input vector: on a row-partitioned context (b2o_comm_init) that would need an
  // all-gather of v, which is not built -- fail loudly instead of multiplying slabs.  One block per GPU (BlockDiagonalOperator)
  // uses local contexts.
  if (d->ctx->nranks > 1)
    B2O_FAIL(B2O_EUNSUPPORTED, "LinearOperator(M) is not row-partitioned: create it on a local (single-GPU) context");
  const int64_t in_len = trans ? d->m : d->n, out_len = trans ? d->n : d->m;
  if (v_len != in_len || res_len != out_len) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if ((out_len > 0 && !res) || (in_len > 0 && !v)) B2O_FAIL(B2O_EARG, "null vector");
  const size_t E = dense_elem(d->dtype);
  if (((uintptr_t)res | (uintptr_t)v) % E) B2O_FAIL(B2O_EARG, "dense: vectors not aligned to the element size");
  B2O_CUDA(cudaSetDevice(d->ctx->device));
  if (d->dtype == B2O_F64) return dense_run<double>(d, trans, res, v, alpha, beta);
  return dense_run<float>(d, trans, res, v, alpha, beta);
}

// algorithmic DRAM bytes of one product: the matrix once, the input vector once, res written (+ read when β != 0)
extern "C" int b2o_dense_apply_bytes(b2o_dense *d, int trans, double beta, double *bytes) {
  if (!d || !bytes) B2O_FAIL(B2O_EARG, "null argument");
  const double E = (double)dense_elem(d->dtype);
  const double in_len = (double)(trans ? d->m : d->n), out_len = (double)(trans ? d->n : d->m);
  *bytes = E * ((double)d->m * (double)d->n + in_len + out_len * (beta != 0.0 ? 2.0 : 1.0));
  return B2O_OK;
}
