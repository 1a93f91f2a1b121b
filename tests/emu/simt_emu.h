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
#include <map>
#include <memory>
#include <mutex>
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
  std::vector<unsigned char> dyn;                                  // dynamic shared memory (+ slack for 128-byte alignment)
  std::mutex named_mu;
  std::map<int, std::unique_ptr<std::barrier<>>> named;           // bar.sync id, count
  explicit Block(unsigned n, size_t smem = 0) : nthreads(n), bar(n), xchg(n), dyn(smem + 128) {
    for (unsigned w = 0; w < (n + 31) / 32; ++w) warp_bar.emplace_back(new std::barrier<>(std::min(32u, n - 32 * w)));
  }
};
inline thread_local Block *cur_block = nullptr;
inline std::string last_error;
}  // namespace emu
inline thread_local dim3 threadIdx, blockIdx;
inline dim3 gridDim, blockDim;

namespace emu {
inline unsigned char *dyn_smem() {
  unsigned char *p = cur_block->dyn.data();
  return p + ((128 - ((uintptr_t)p & 127)) & 127);
}
// bar.sync id, count: the first `count` arrivals form the barrier (all callers pass the same count for an id)
inline void named_barrier_sync(int id, int count) {
  Block *b = cur_block;
  std::barrier<> *bar;
  {
    std::lock_guard<std::mutex> g(b->named_mu);
    auto &slot = b->named[id];
    if (!slot) slot.reset(new std::barrier<>(count));
    bar = slot.get();
  }
  bar->arrive_and_wait();
}
// ---- mbarrier + 1-D bulk copy (cp.async.bulk ... mbarrier::complete_tx) restated for the host: the 8-byte barrier word
// packs pending arrivals [0,16), arrival count [16,32), outstanding transaction bytes [32,63) and the phase bit [63].
// A phase completes when the pending arrivals AND the outstanding bytes reach zero; waiting on parity P returns once the
// phase of parity P is over.  Checks the alignment rules of the real instruction (16-byte addresses and sizes).
inline std::mutex mbar_mu;
inline void mbar_check(uint64_t *b) {
  const uint64_t pend = *b & 0xffffu, init = (*b >> 16) & 0xffffu, tx = (*b >> 32) & 0x7fffffffu, ph = *b >> 63;
  if (pend == 0 && tx == 0) *b = ((ph ^ 1u) << 63) | (init << 16) | init;
}
}  // namespace emu
inline void mbar_init(uint64_t *b, uint32_t count) {
  std::lock_guard<std::mutex> g(emu::mbar_mu);
  *b = ((uint64_t)count << 16) | count;
}
inline void mbar_fence_init() {}
inline void fence_proxy_async_smem() {}
inline void mbar_expect_tx(uint64_t *b, uint32_t bytes) {      // mbarrier.arrive.expect_tx
  std::lock_guard<std::mutex> g(emu::mbar_mu);
  if ((*b & 0xffffu) == 0) {
    fprintf(stderr, "emu: mbarrier arrive with no pending arrivals\n");
    abort();
  }
  *b += ((uint64_t)bytes << 32);
  *b -= 1;
  emu::mbar_check(b);
}
inline void mbar_arrive(uint64_t *b) {
  std::lock_guard<std::mutex> g(emu::mbar_mu);
  if ((*b & 0xffffu) == 0) {
    fprintf(stderr, "emu: mbarrier arrive with no pending arrivals\n");
    abort();
  }
  *b -= 1;
  emu::mbar_check(b);
}
inline void mbar_wait(uint64_t *b, uint32_t parity) {
  for (;;) {
    {
      std::lock_guard<std::mutex> g(emu::mbar_mu);
      if ((uint32_t)(*b >> 63) != (parity & 1u)) return;
    }
    std::this_thread::yield();
  }
}
inline void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *b) {
  if (((uintptr_t)dst & 15) || ((uintptr_t)src & 15) || (bytes & 15) || bytes == 0) {
    fprintf(stderr, "emu: bulk copy breaks the 16-byte rules (dst %p src %p bytes %u)\n", dst, src, bytes);
    abort();
  }
  memcpy(dst, src, bytes);
  std::lock_guard<std::mutex> g(emu::mbar_mu);
  if (((*b >> 32) & 0x7fffffffu) < bytes) {
    fprintf(stderr, "emu: bulk copy completes more bytes than expected\n");
    abort();
  }
  *b -= ((uint64_t)bytes << 32);
  emu::mbar_check(b);
}

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
inline void launch(dim3 grid, dim3 block, size_t smem, F fn) {
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
        if (t == 0) shared_blk = new Block(nthreads, smem);
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
#define B2O_LAUNCH(kern, grid, block, smem, stream, ...) emu::launch((grid), (block), (size_t)(smem), [&] { kern(__VA_ARGS__); })
