// simt_emu.h -- a minimal host-side SIMT emulator (TEST INFRASTRUCTURE, not the product).
//
// Lets tests/ compile kernel headers of csrc/ (those written without CUDA runtime calls, e.g. b2o_dense_kernels.cuh) as
// plain C++ and run them on a CPU-only box: every CUDA thread of a block is an OS thread, __syncthreads() is a block
// barrier, warp shuffles exchange through a per-warp buffer with warp barriers, `__shared__` becomes a static (blocks run
// one after the other).  It checks INDEX LOGIC and launch planning (grid shapes, ragged edges, split bookkeeping); it says
// nothing about PTX, memory ordering or performance -- the `-m gpu` parity tests on the B200 remain the proof.
#pragma once
#define B2O_SIMT_EMU 1
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include "../../include/b2o.h"

// ---- CUDA vocabulary ----------------------------------------------------------------------------------------------
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint4 {
  unsigned x, y, z, w;
};
struct double2 {
  double x, y;
};
using std::max;
using std::min;

namespace emu {
struct Block {
  unsigned nthreads;
  std::barrier<> bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<double> xchg;  // [nthreads]
  explicit Block(unsigned n) : nthreads(n), bar(n), xchg(n) {
    for (unsigned w = 0; w < (n + 31) / 32; ++w) warp_bar.emplace_back(new std::barrier<>(std::min(32u, n - 32 * w)));
  }
};
inline thread_local Block *cur_block = nullptr;
inline std::string last_error;
}  // namespace emu
inline thread_local dim3 threadIdx, blockIdx;
inline dim3 gridDim, blockDim;

inline void __syncthreads() { emu::cur_block->bar.arrive_and_wait(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::cur_block->warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
inline double __shfl_xor_sync(unsigned, double v, int lane_mask) {
  emu::Block *b = emu::cur_block;
  std::barrier<> &wb = *b->warp_bar[threadIdx.x >> 5];
  b->xchg[threadIdx.x] = v;
  wb.arrive_and_wait();
  const double r = b->xchg[threadIdx.x ^ (unsigned)lane_mask];
  wb.arrive_and_wait();
  return r;
}
template <typename T>
inline T __ldg(const T *p) {
  return *p;
}
template <typename T>
inline T __ldcg(const T *p) {
  return *p;
}
inline double __hiloint2double(int hi, int lo) {
  const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
  double d;
  memcpy(&d, &u, 8);
  return d;
}
inline float __uint_as_float(unsigned u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- the library's error / launch macros -------------------------------------------------------------------------
inline void b2o_set_error(const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  emu::last_error = buf;
}
#define B2O_FAIL(code, ...)     \
  do {                          \
    b2o_set_error(__VA_ARGS__); \
    return (code);              \
  } while (0)
#define B2O_CUDA(expr) \
  do {                 \
  } while (0)
#define B2O_STREAM_T void *

namespace emu {
// run `fn` for every thread of every block of the grid: blocks one after the other (a `__shared__` static belongs to one
// block at a time), the threads of a block concurrently.  One pool of OS threads per launch walks over the blocks.
template <typename F>
inline void launch(dim3 grid, dim3 block, F fn) {
  gridDim = grid;
  blockDim = block;
  const unsigned nthreads = block.x * block.y * block.z;
  const size_t nblocks = (size_t)grid.x * grid.y * grid.z;
  if (nblocks == 0) return;
  std::barrier<> between_blocks(nthreads);
  std::vector<std::thread> ts;
  ts.reserve(nthreads);
  for (unsigned t = 0; t < nthreads; ++t)
    ts.emplace_back([&, t] {
      threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      for (size_t b = 0; b < nblocks; ++b) {
        static Block *shared_blk = nullptr;           // published by thread 0, read by all after the barrier
        if (t == 0) shared_blk = new Block(nthreads);
        between_blocks.arrive_and_wait();
        Block *blk = shared_blk;
        cur_block = blk;
        blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((size_t)grid.x * grid.y)));
        fn();
        // a thread that returned early must not leave the others waiting at a barrier
        blk->bar.arrive_and_drop();
        blk->warp_bar[t >> 5]->arrive_and_drop();
        between_blocks.arrive_and_wait();
        if (t == 0) delete blk;
      }
    });
  for (auto &th : ts) th.join();
}
}  // namespace emu
#define B2O_LAUNCH(kern, grid, block, smem, stream, ...) emu::launch((grid), (block), [&] { kern(__VA_ARGS__); })
