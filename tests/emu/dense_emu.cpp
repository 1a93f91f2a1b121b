// dense_emu.cpp -- host build of csrc/b2o_dense_kernels.cuh under the SIMT emulator (TEST INFRASTRUCTURE).
// tests/test_emu_dense.py compiles this with g++ and drives it through ctypes: the SAME kernel bodies, split planning and
// launch logic the product compiles with nvcc, checked against the oracle on a CPU-only box.
//
// Built with -fvisibility=hidden -Wl,-Bsymbolic and the kernels in their own namespace: libb2o.so (loaded RTLD_GLOBAL by the
// product) exports host stubs with the same template names, which must never be bound in place of the emulated bodies.
#include "simt_emu.h"
namespace emu_dense {
#include "../../linearoperators.jl_b200/csrc/b2o_dense_kernels.cuh"
}
using namespace emu_dense;
#define EMU_API __attribute__((visibility("default")))

extern "C" {
EMU_API const char *emu_last_error() { return emu::last_error.c_str(); }

// mirrors b2o_dense_create + b2o_dense_apply (csrc/b2o_dense.cu): workspace sized like create, then dense_run_impl
EMU_API int emu_dense_apply(int dtype, int trans, int64_t m, int64_t n, int64_t lda, const void *M, void *res, const void *v, double alpha,
                    double beta, int num_sms, int force_scalar, int64_t *launches) {
  const int Wv = dtype == B2O_F64 ? 2 : 4;
  std::vector<double> part(dense_workspace_elems(num_sms, Wv, m, n) + 1, std::nan(""));
  *launches = 0;
  if (dtype == B2O_F64)
    return dense_run_impl<double>(num_sms, nullptr, launches, M, m, n, lda, part.data(), part.size() - 1, trans, res, v, alpha, beta,
                                  force_scalar);
  return dense_run_impl<float>(num_sms, nullptr, launches, M, m, n, lda, part.data(), part.size() - 1, trans, res, v, alpha, beta,
                               force_scalar);
}

EMU_API void emu_dense_plan(int num_sms, int trans, int64_t m, int64_t n, int W, int64_t *gx, int64_t *chunk, int *nsplit,
                            int *narrow, int *tx_log2) {
  const DensePlan pl = dense_plan(num_sms, trans, m, n, W);
  *gx = pl.gx;
  *chunk = pl.chunk;
  *nsplit = pl.nsplit;
  *narrow = pl.narrow;
  *tx_log2 = pl.tx_log2;
}
}
