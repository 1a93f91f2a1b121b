// b2o_shared_defs.h -- constants and plain structs shared by the kernels' headers and the host code; no CUDA includes, so the
// kernel headers can also be compiled for the host SIMT emulator of tests/emu (test infrastructure).
#pragma once
#include <stddef.h>
#include <stdint.h>

constexpr int B2O_MAX_COLS = 128;      // column streams one launch can address
constexpr int B2O_MAX_GRID = 1024;     // upper bound on persistent grid
constexpr int B2O_WS_DOTS = 1024;      // doubles reserved for reduced scalars
constexpr int B2O_WS_SWEEP = 512;      // [512, 512+129): inner products of the last fused two-loop launch, sweep order
constexpr int B2O_WS_QNDBG = 768;      // [768, 771): mailbox timing accumulators (ns waited for local CTAs, ns in the exchange, epochs)

// ---- peer mailbox layout (one per GPU): vals[2][8][128] doubles, then flags[8] u64
constexpr int MBOX_MAXV = 128, MBOX_MAXR = 8;
constexpr size_t MBOX_FLAGS_OFF = sizeof(double) * 2 * MBOX_MAXR * MBOX_MAXV;
constexpr size_t MBOX_BYTES = MBOX_FLAGS_OFF + sizeof(unsigned long long) * MBOX_MAXR;
struct MboxDev {
  double *vals[MBOX_MAXR];                 // vals region of every rank's mailbox (own entry = local memory)
  unsigned long long *flags[MBOX_MAXR];
  unsigned long long *ready;               // local "dots are published" flag for the other CTAs of this GPU
  int nranks, rank;
  unsigned long long epoch_base;           // epoch of the first all-reduce of this launch is epoch_base + 1
};
