// Stable LSD radix sort of (key, value) pairs for the backward plan (sm_100a).
//
// Two implementations behind one entry point (RECEMB_PLAN_SORT = "cub" (default) | "own"):
//   cub   cub::DeviceRadixSort::SortPairs (onesweep: one histogram pass + 8-bit passes) -- a library call,
//         like cuBLAS for a plain GEMM.  Measured on B200 at cfg 2 (16.4 M pairs, 24-bit keys): 0.39 ms.
//   own   the hand-written sort below.  Measured: 1.10 ms (hist 0.07 + 0.12, scatter 0.39 + 0.51 ms,
//         profiles/r02_sort.md) -- correct, stable, deterministic, but latency-bound: ranking 32 pairs among
//         4096 bins costs a chain of ~5 dependent shared-memory round trips per iteration, and the 12 KB of
//         private counters + tags per warp leave room for only 16 warps per SM to hide them (ncu: issue
//         slots busy 14 %, short-scoreboard stalls 52 %).  It stays selectable and tested; it is not the
//         default because the plan shares the SMs and HBM with the forward gather it is hidden under.
//
// The hand-written design:
//
// Why not a library sort: the plan's keys are table rows -- bit_width(rows) = 24 bits for the stacked
// 10 x 1M tables of cfg 2 -- and a general 8-bit-per-pass sort needs 3 passes + a histogram pass over
// them.  Here a pass takes up to 12 bits (4096 bins), so 24-bit keys need TWO passes:
//
//   hist     every CTA owns one tile of consecutive pairs and counts its digits          [C][NB] counts
//   scan     column-wise exclusive scan over the CTAs + exclusive scan of the bin totals  (tiny)
//   scatter  the CTA re-walks its tile (L2-resident by then): warp w owns a consecutive sub-range
//            and a PRIVATE row of 16-bit counters in shared memory; __match_any_sync ranks the
//            32 lanes of an iteration among equal digits in lane (= input) order.  Position =
//            bin base + CTA prefix + prefix over lower warps + earlier iterations + lane rank:
//            stable by construction, no atomics, no look-back, bit-reproducible.
//
// Stores go straight to global memory (8 useful bytes per lane): the write frontier -- NB bins x
// resident CTAs x one 32-byte sector -- is a few MB and lives in the 126 MB L2, which merges the
// partial sectors before they reach HBM.  Skewed keys (Zipf heads, k-shift collapse rows) only make
// one bin long: every CTA still owns the same number of pairs.
#pragma once

#include "common.cuh"

namespace recemb {

constexpr int kSortWarps = 16;
constexpr int kSortThreads = kSortWarps * 32;
constexpr int kSortMaxBits = 12;
constexpr int kSortMaxPasses = 4;
constexpr int kSortMaxPerWarp = 4064;  // tile <= 65024 pairs: prefixes inside a tile fit 16 bits

struct SortShape {
  bool own;           // hand-written passes (else the library sort: one call, input in a, output in b)
  int passes;
  int bits[kSortMaxPasses];
  int per_warp;       // pairs per warp sub-range (multiple of 32)
  int64_t ctas;       // tiles
  size_t temp_bytes;  // [ctas][NB_max] u32 counts + [NB_max] totals + [NB_max] bases
};

// Shape of the sort of n pairs with key_bits significant key bits on `device`.
SortShape sort_shape(int64_t n, int key_bits, int device);

// True when the unsorted pairs must be written to the (keys_b, vals_b) buffers: the passes
// ping-pong between a and b and the LAST pass always lands in b.
inline bool sort_input_in_b(const SortShape& s) { return (s.passes % 2) == 0; }

// Sorts by key (stable): input in a (or b, see sort_input_in_b), result in b.  Asynchronous on `stream`.
// Returns a recemb_status; every kernel launch is counted.
int sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int64_t n, int key_bits,
               void* temp, size_t temp_bytes, int device, cudaStream_t stream);

}  // namespace recemb
