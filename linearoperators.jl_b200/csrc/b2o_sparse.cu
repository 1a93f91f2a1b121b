// b2o_sparse.cu -- LinearOperator(M) for a sparse matrix (src/constructors.jl:15-29 with M::SparseMatrixCSC; the docstring
// :3-5 says "dense or sparse", test/test_linop.jl:41-75 runs every predicate on both).
//
// The reference's closures are `mul!(res, M, v, α, β)` and the same with transpose(M) / adjoint(M); for a SparseMatrixCSC
// they land in the SparseArrays standard library (Project.toml:9,44 -- not vendored under /root/reference): a column sweep
// with scattered `res[rowval[k]] += nzval[k] * (α v[j])` for the product, a per-column dot for the transposed product.
// Here BOTH directions are the gather form (one compressed row per lane group, fixed summation order, no atomics -> results
// are bit-reproducible): the structure is transposed ONCE at create time, on the host (stable counting sort: index work, no
// floating point), and both copies stay in HBM.  The values of the given orientation are aliased like the reference's
// closures capture M; the transposed copy holds gathered values (b2o_sparse_refresh re-gathers them after nzval changed in
// place).  Index work is exact; sums are taken in double for both element types.
//
// Algorithmic bytes per product: nnz*(E + 4) (values + 32-bit indices) + 8*(rows+1) (offsets) + the vectors; the gathers
// from x are sector-granular (32 B per random hit), so DRAM traffic can legitimately exceed that for scattered patterns.
#include "b2o_internal.cuh"
#include "b2o_sparse_kernels.cuh"
#include <vector>

struct b2o_sparse_s {
  b2o_ctx *ctx = nullptr;
  int dtype = B2O_F64, fmt = 0;
  int64_t m = 0, n = 0, nnz = 0;
  // orientation [0] serves prod! (compressed rows of M), [1] serves tprod!/ctprod! (compressed rows of Mᵀ = columns of M)
  int64_t *ptr[2] = {nullptr, nullptr};
  int32_t *idx[2] = {nullptr, nullptr};
  const void *val[2] = {nullptr, nullptr};
  int given = 0;                 // which orientation aliases the caller's values (CSR: 0, CSC: 1)
  void *tval = nullptr;          // gathered values of the other orientation (owned)
  int64_t *perm = nullptr;       // tval[k] = vals[perm[k]]
  SpTile *tiles[2] = {nullptr, nullptr};   // tile descriptors of the TMA-staged kernel, per orientation (ntiles + 1 entries)
  int64_t ntiles[2] = {0, 0};
};

static inline size_t sparse_elem(int dtype) { return dtype == B2O_F64 ? 8 : 4; }

static int sparse_regather(b2o_sparse *s) {
  if (s->nnz == 0) return B2O_OK;
  b2o_ctx *c = s->ctx;
  const int64_t grid = std::max<int64_t>(1, std::min<int64_t>((s->nnz + SP_THREADS - 1) / SP_THREADS, (int64_t)c->num_sms * 8));
  if (s->dtype == B2O_F64)
    perm_gather_kernel<double><<<(unsigned)grid, SP_THREADS, 0, c->stream>>>((double *)s->tval, (const double *)s->val[s->given], s->perm, s->nnz);
  else
    perm_gather_kernel<float><<<(unsigned)grid, SP_THREADS, 0, c->stream>>>((float *)s->tval, (const float *)s->val[s->given], s->perm, s->nnz);
  c->launches++;
  B2O_CUDA(cudaGetLastError());
  return B2O_OK;
}

static void sparse_free(b2o_sparse *s) {
  for (int o = 0; o < 2; ++o) {
    cudaFree(s->ptr[o]);
    cudaFree(s->idx[o]);
    cudaFree(s->tiles[o]);
  }
  cudaFree(s->tval);
  cudaFree(s->perm);
  delete s;
}

// fmt 0: CSC (Julia's SparseMatrixCSC: colptr[n+1], rowval[nnz], nzval[nnz]); fmt 1: CSR (rowptr[m+1], colval[nnz]).
// ptr1 / idx1 are HOST arrays with the reference's 1-based values; vals is a DEVICE array (borrowed).
extern "C" int b2o_sparse_create(b2o_ctx *ctx, int dtype, int fmt, int64_t m, int64_t n, int64_t nnz, const int64_t *ptr1,
                                 const int64_t *idx1, const void *vals, b2o_sparse **out) {
  if (!ctx || !out || !ptr1) B2O_FAIL(B2O_EARG, "null argument");
  if (dtype != B2O_F64 && dtype != B2O_F32) B2O_FAIL(B2O_EUNSUPPORTED, "sparse: dtype %d not supported (Float64, Float32)", dtype);
  if (fmt != 0 && fmt != 1) B2O_FAIL(B2O_EARG, "sparse: format must be 0 (CSC) or 1 (CSR)");
  if (m < 0 || n < 0 || nnz < 0) B2O_FAIL(B2O_EARG, "negative size");
  if (m >= 0x7fffffffLL || n >= 0x7fffffffLL) B2O_FAIL(B2O_EUNSUPPORTED, "sparse: dimensions must fit 32-bit indices");
  if (nnz > 0 && (!idx1 || !vals)) B2O_FAIL(B2O_EARG, "null index / value array");
  if ((uintptr_t)vals % sparse_elem(dtype)) B2O_FAIL(B2O_EARG, "sparse: values not aligned to the element size");
  const int64_t np = fmt == 0 ? n : m;       // compressed dimension (number of pointer intervals)
  const int64_t nd = fmt == 0 ? m : n;       // range of the stored indices
  if (ptr1[0] != 1 || ptr1[np] != nnz + 1) B2O_FAIL(B2O_EARG, "sparse: pointer array must start at 1 and end at nnz + 1");
  for (int64_t j = 0; j < np; ++j)
    if (ptr1[j + 1] < ptr1[j]) B2O_FAIL(B2O_EARG, "sparse: pointer array must be nondecreasing");
  for (int64_t k = 0; k < nnz; ++k)
    if (idx1[k] < 1 || idx1[k] > nd) B2O_FAIL(B2O_EARG, "sparse: index %lld outside 1..%lld", (long long)idx1[k], (long long)nd);
  B2O_CUDA(cudaSetDevice(ctx->device));

  // given orientation, 0-based; then its transpose by a stable counting sort over the stored indices (host, exact)
  std::vector<int64_t> gptr(np + 1), tptr(nd + 1, 0), perm(std::max<int64_t>(nnz, 1));
  std::vector<int32_t> gidx(std::max<int64_t>(nnz, 1)), tidx(std::max<int64_t>(nnz, 1));
  for (int64_t j = 0; j <= np; ++j) gptr[j] = ptr1[j] - 1;
  for (int64_t k = 0; k < nnz; ++k) {
    gidx[k] = (int32_t)(idx1[k] - 1);
    tptr[gidx[k] + 1]++;
  }
  for (int64_t i = 0; i < nd; ++i) tptr[i + 1] += tptr[i];
  {
    std::vector<int64_t> fill(tptr.begin(), tptr.end() - 1);
    for (int64_t j = 0; j < np; ++j)
      for (int64_t k = gptr[j]; k < gptr[j + 1]; ++k) {
        const int64_t dst = fill[gidx[k]]++;
        tidx[dst] = (int32_t)j;
        perm[dst] = k;
      }
  }

  // tiles of the TMA-staged kernel for both orientations (index work, host)
  std::vector<SpTile> gtiles, ttiles;
  const int64_t ngt = spmv_build_tiles(gptr.data(), np, nnz, gtiles);
  const int64_t ntt = spmv_build_tiles(tptr.data(), nd, nnz, ttiles);

  b2o_sparse *s = new b2o_sparse_s();
  s->ctx = ctx;
  s->dtype = dtype;
  s->fmt = fmt;
  s->m = m;
  s->n = n;
  s->nnz = nnz;
  s->given = fmt == 0 ? 1 : 0;               // CSC arrays are the compressed rows of Mᵀ
  const int g = s->given, t = 1 - g;
  const size_t nz = (size_t)std::max<int64_t>(nnz, 1);
  // offset arrays carry 3 spare (zeroed) entries: the bulk copies of the tile kernel read an even number of offsets
  bool ok = cudaMalloc(&s->ptr[g], sizeof(int64_t) * (np + 4)) == cudaSuccess &&
            cudaMalloc(&s->idx[g], sizeof(int32_t) * nz) == cudaSuccess &&
            cudaMalloc(&s->ptr[t], sizeof(int64_t) * (nd + 4)) == cudaSuccess &&
            cudaMalloc(&s->tiles[g], sizeof(SpTile) * gtiles.size()) == cudaSuccess &&
            cudaMalloc(&s->tiles[t], sizeof(SpTile) * ttiles.size()) == cudaSuccess &&
            cudaMalloc(&s->idx[t], sizeof(int32_t) * nz) == cudaSuccess &&
            cudaMalloc(&s->tval, sparse_elem(dtype) * nz) == cudaSuccess && cudaMalloc(&s->perm, sizeof(int64_t) * nz) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    sparse_free(s);
    B2O_FAIL(B2O_ENOMEM, "sparse: allocation failed");
  }
  s->val[g] = vals;
  s->val[t] = s->tval;
  s->ntiles[g] = ngt;
  s->ntiles[t] = ntt;
  cudaError_t e = cudaMemsetAsync(s->ptr[g], 0, sizeof(int64_t) * (np + 4), ctx->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(s->ptr[t], 0, sizeof(int64_t) * (nd + 4), ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->tiles[g], gtiles.data(), sizeof(SpTile) * gtiles.size(), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->tiles[t], ttiles.data(), sizeof(SpTile) * ttiles.size(), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->ptr[g], gptr.data(), sizeof(int64_t) * (np + 1), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->idx[g], gidx.data(), sizeof(int32_t) * nz, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->ptr[t], tptr.data(), sizeof(int64_t) * (nd + 1), cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->idx[t], tidx.data(), sizeof(int32_t) * nz, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(s->perm, perm.data(), sizeof(int64_t) * nz, cudaMemcpyHostToDevice, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);    // the host vectors die at return
  if (e != cudaSuccess) {
    sparse_free(s);
    B2O_FAIL(B2O_ECUDA, "sparse: upload failed: %s", cudaGetErrorString(e));
  }
  const int rc = sparse_regather(s);
  if (rc != B2O_OK) {
    sparse_free(s);
    return rc;
  }
  *out = s;
  return B2O_OK;
}

extern "C" int b2o_sparse_destroy(b2o_sparse *s) {
  if (!s) return B2O_OK;
  cudaSetDevice(s->ctx->device);
  cudaStreamSynchronize(s->ctx->stream);
  sparse_free(s);
  return B2O_OK;
}

// after the caller changed nzval in place (the operator aliases it): re-gather the values of the transposed copy
extern "C" int b2o_sparse_refresh(b2o_sparse *s) {
  if (!s) B2O_FAIL(B2O_EARG, "null operator");
  B2O_CUDA(cudaSetDevice(s->ctx->device));
  return sparse_regather(s);
}

// trans = 0: prod!  mul!(res, M, v, α, β);  trans = 1: tprod!/ctprod!  mul!(res, transpose(M), u, α, β)
extern "C" int b2o_sparse_apply(b2o_sparse *s, int trans, void *res, int64_t res_len, const void *v, int64_t v_len, double alpha,
                                double beta) {
  if (!s) B2O_FAIL(B2O_EARG, "null operator");
  // a matrix leaf multiplies with the WHOLE input vector: on a row-partitioned context (b2o_comm_init) that would need an
  // all-gather of v, which is not built -- fail loudly instead of multiplying slabs.  One block per GPU (BlockDiagonalOperator)
  // uses local contexts.
  if (s->ctx->nranks > 1)
    B2O_FAIL(B2O_EUNSUPPORTED, "LinearOperator(M) is not row-partitioned: create it on a local (single-GPU) context");
  const int64_t in_len = trans ? s->m : s->n, out_len = trans ? s->n : s->m;
  if (v_len != in_len || res_len != out_len) B2O_FAIL(B2O_ESHAPE, "shape mismatch");
  if ((out_len > 0 && !res) || (in_len > 0 && !v)) B2O_FAIL(B2O_EARG, "null vector");
  if (((uintptr_t)res | (uintptr_t)v) % sparse_elem(s->dtype)) B2O_FAIL(B2O_EARG, "sparse: vectors not aligned to the element size");
  b2o_ctx *c = s->ctx;
  B2O_CUDA(cudaSetDevice(c->device));
  const int o = trans ? 1 : 0;
  // sparse_kernel option: 0 / 3 the software-pipelined row kernel (default), 1 the plain row kernel (same bits), 2 the
  // TMA-staged tile kernel (needs 16-byte aligned values).  Measured on B200 (profiles/r1_sparse.jsonl): pipelined rows
  // 0.150 / 0.058 ms, plain rows 0.164 / 0.062 ms, tiles 0.167 / 0.075 ms on the two Float64 test patterns.
  const bool tiles = c->sparse_kernel == 2 && s->ntiles[o] > 0 && ((uintptr_t)s->val[o] & 15) == 0;
  if (tiles) {
    if (s->dtype == B2O_F64)
      return spmv_tiles_run_impl<double>(c->num_sms, c->stream, &c->launches, s->tiles[o], s->ntiles[o], s->ptr[o], s->idx[o],
                                         s->val[o], out_len, s->nnz, res, v, alpha, beta, c->sparse_lanes);
    return spmv_tiles_run_impl<float>(c->num_sms, c->stream, &c->launches, s->tiles[o], s->ntiles[o], s->ptr[o], s->idx[o],
                                      s->val[o], out_len, s->nnz, res, v, alpha, beta, c->sparse_lanes);
  }
  const bool pipe = c->sparse_kernel != 1;
  if (s->dtype == B2O_F64)
    return spmv_run_impl<double>(c->num_sms, c->stream, &c->launches, s->ptr[o], s->idx[o], s->val[o], out_len, s->nnz, res, v, alpha, beta,
                                 pipe, c->sparse_lanes);
  return spmv_run_impl<float>(c->num_sms, c->stream, &c->launches, s->ptr[o], s->idx[o], s->val[o], out_len, s->nnz, res, v, alpha, beta,
                              pipe, c->sparse_lanes);
}

extern "C" int b2o_sparse_apply_bytes(b2o_sparse *s, int trans, double beta, double *bytes) {
  if (!s || !bytes) B2O_FAIL(B2O_EARG, "null argument");
  const double E = (double)sparse_elem(s->dtype);
  const double in_len = (double)(trans ? s->m : s->n), out_len = (double)(trans ? s->n : s->m);
  *bytes = (double)s->nnz * (E + 4.0) + 8.0 * (out_len + 1.0) + E * (in_len + out_len * (beta != 0.0 ? 2.0 : 1.0));
  return B2O_OK;
}
