// sparse_emu.cpp -- host build of csrc/b2o_sparse_kernels.cuh under the SIMT emulator (TEST INFRASTRUCTURE).
// Mirrors b2o_sparse_create (structure transposition by a stable counting sort) + b2o_sparse_apply of csrc/b2o_sparse.cu:
// the row kernel, lane-group sizing and launch logic are the product's own code; the transposition is restated here because
// the product's sits next to its cudaMalloc/cudaMemcpy calls.  Built with hidden visibility + -Bsymbolic (see dense_emu.cpp).
#include "simt_emu.h"
namespace emu_sparse {
#include "../../linearoperators.jl_b200/csrc/b2o_sparse_kernels.cuh"
}
using namespace emu_sparse;
#define EMU_API __attribute__((visibility("default")))

extern "C" {
EMU_API const char *emu_sparse_last_error() { return emu::last_error.c_str(); }

// fmt 0 CSC / 1 CSR with 1-based host arrays, exactly the arguments of b2o_sparse_create; one product like b2o_sparse_apply
EMU_API int emu_sparse_apply(int dtype, int fmt, int64_t m, int64_t n, int64_t nnz, const int64_t *ptr1, const int64_t *idx1,
                             const void *vals, int trans, void *res, const void *v, double alpha, double beta, int num_sms,
                             int64_t *launches, int *lanes_log2, int kernel, int64_t *ntiles_out, int lanes_override) {
  const int64_t np = fmt == 0 ? n : m, nd = fmt == 0 ? m : n;
  std::vector<int64_t> gptr(np + 1), tptr(nd + 1, 0), perm(std::max<int64_t>(nnz, 1));
  std::vector<int32_t> gidx(std::max<int64_t>(nnz, 1)), tidx(std::max<int64_t>(nnz, 1));
  for (int64_t j = 0; j <= np; ++j) gptr[j] = ptr1[j] - 1;
  for (int64_t k = 0; k < nnz; ++k) {
    gidx[k] = (int32_t)(idx1[k] - 1);
    tptr[gidx[k] + 1]++;
  }
  for (int64_t i = 0; i < nd; ++i) tptr[i + 1] += tptr[i];
  std::vector<int64_t> fill(tptr.begin(), tptr.end() - 1);
  for (int64_t j = 0; j < np; ++j)
    for (int64_t k = gptr[j]; k < gptr[j + 1]; ++k) {
      const int64_t dst = fill[gidx[k]]++;
      tidx[dst] = (int32_t)j;
      perm[dst] = k;
    }
  const size_t E = dtype == B2O_F64 ? 8 : 4;
  std::vector<unsigned char> tval(std::max<int64_t>(nnz, 1) * E);
  *launches = 0;
  if (nnz > 0) {   // the product's gather kernel
    if (dtype == B2O_F64) {
      void (*k)(double *, const double *, const int64_t *, int64_t) = perm_gather_kernel<double>;
      B2O_LAUNCH(k, dim3(2), dim3(SP_THREADS), 0, nullptr, (double *)tval.data(), (const double *)vals, perm.data(), nnz);
    } else {
      void (*k)(float *, const float *, const int64_t *, int64_t) = perm_gather_kernel<float>;
      B2O_LAUNCH(k, dim3(2), dim3(SP_THREADS), 0, nullptr, (float *)tval.data(), (const float *)vals, perm.data(), nnz);
    }
  }
  const int given = fmt == 0 ? 1 : 0, o = trans ? 1 : 0;
  const int64_t *ptr = o == given ? gptr.data() : tptr.data();
  const int32_t *idx = o == given ? gidx.data() : tidx.data();
  const void *val = o == given ? vals : (const void *)tval.data();
  const int64_t out_len = trans ? n : m;
  *lanes_log2 = lanes_override >= 0 ? lanes_override : spmv_lanes_log2(out_len, nnz);
  *ntiles_out = 0;
  if (kernel == 2) {
    // the TMA-staged tile kernel: offsets padded with zeros like b2o_sparse_create, every staged array 16-byte aligned
    std::vector<SpTile> tiles;
    const int64_t ntiles = spmv_build_tiles(ptr, out_len, nnz, tiles);
    *ntiles_out = ntiles;
    auto aligned = [](const void *src, size_t bytes, size_t alloc) {
      void *p = nullptr;
      if (posix_memalign(&p, 64, std::max<size_t>(alloc, 64))) abort();
      memset(p, 0, std::max<size_t>(alloc, 64));
      if (bytes) memcpy(p, src, bytes);
      return p;
    };
    int64_t *aptr = (int64_t *)aligned(ptr, sizeof(int64_t) * (out_len + 1), sizeof(int64_t) * (out_len + 4));
    int32_t *aidx = (int32_t *)aligned(idx, sizeof(int32_t) * nnz, sizeof(int32_t) * nnz);
    void *aval = aligned(val, E * nnz, E * nnz);                  // exact size: reads past nnz would be caught by ASan builds
    int rc;
    if (dtype == B2O_F64)
      rc = spmv_tiles_run_impl<double>(num_sms, nullptr, launches, tiles.data(), ntiles, aptr, aidx, aval, out_len, nnz, res, v, alpha, beta,
                                       lanes_override);
    else
      rc = spmv_tiles_run_impl<float>(num_sms, nullptr, launches, tiles.data(), ntiles, aptr, aidx, aval, out_len, nnz, res, v, alpha, beta,
                                      lanes_override);
    free(aptr);
    free(aidx);
    free(aval);
    return rc;
  }
  const bool pipe = kernel == 3;       // the software-pipelined row kernel
  if (dtype == B2O_F64)
    return spmv_run_impl<double>(num_sms, nullptr, launches, ptr, idx, val, out_len, nnz, res, v, alpha, beta, pipe, lanes_override);
  return spmv_run_impl<float>(num_sms, nullptr, launches, ptr, idx, val, out_len, nnz, res, v, alpha, beta, pipe, lanes_override);
}

// the tile cut of a 0-based offset array (spmv_build_tiles of the product): descriptors flattened to (e0, r0, ne) triples
EMU_API int64_t emu_sparse_tiles(const int64_t *ptr0, int64_t nrows, int64_t nnz, int64_t *out, int64_t cap, int *tile_c, int *tile_rt) {
  std::vector<SpTile> tiles;
  const int64_t n = spmv_build_tiles(ptr0, nrows, nnz, tiles);
  *tile_c = ST_C;
  *tile_rt = ST_RT;
  for (int64_t i = 0; i <= n && i < cap; ++i) {
    out[3 * i] = tiles[i].e0;
    out[3 * i + 1] = tiles[i].r0;
    out[3 * i + 2] = tiles[i].ne;
  }
  return n;
}
}
