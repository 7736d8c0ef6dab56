// Hand-written stable LSD radix sort of the backward plan's (row, slot) pairs: see sort.cuh.
#include <cub/device/device_radix_sort.cuh>

#include <cstdlib>
#include <cstring>

#include "sort.cuh"

namespace recemb {

static bool use_own_sort() {
  static int own = -1;
  if (own < 0) {
    const char* e = getenv("RECEMB_PLAN_SORT");
    own = (e && strcmp(e, "own") == 0) ? 1 : 0;
  }
  return own == 1;
}

constexpr int kSortUnroll = 8;  // independent loads in flight per lane and array

struct SortArgs {
  const uint32_t* kin;
  const uint32_t* vin;
  uint32_t* kout;
  uint32_t* vout;
  int64_t n;
  int shift;
  int bits;
  int per_warp;
  uint32_t* counts;  // [ctas][NB]: digit counts per tile, turned into exclusive prefixes over the tiles
  uint32_t* totals;  // [NB]
  uint32_t* bases;   // [NB] exclusive scan of totals
  int64_t ctas;
};

__device__ __forceinline__ uint32_t digit_of(uint32_t key, int shift, uint32_t mask) { return (key >> shift) & mask; }

// ---- warp-private digit counting -----------------------------------------------------------------
// Each warp owns a consecutive sub-range and, in shared memory, a private row of 16-bit counters plus a
// private row of 8-bit tags.  Shared-memory atomics (~2 cycles per lane) or __match_any_sync per 32 pairs
// would make the sort instruction-bound; instead the common case -- the 32 digits of an iteration are all
// different (NB = 4096 bins: 79 % of the iterations on hashed ids) -- is detected with one tag round trip:
// every lane stores its lane id at tag[digit] and reads it back; a lane that reads another id shares its
// digit with a later writer.  No collision: plain `counter[digit] += 1`, rank 0.  Collision (hot rows,
// small tables): the lanes of equal digit are found with one ballot per digit bit.
__device__ __forceinline__ uint32_t match_by_ballots(uint32_t d, int bits, uint32_t valid_mask) {
  uint32_t m = valid_mask;
  for (int b = 0; b < bits; ++b) {
    const bool bit = (d >> b) & 1u;
    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
    m &= bit ? bal : ~bal;
  }
  return m;
}

// returns true when some lanes of this iteration share a digit (warp-uniform)
__device__ __forceinline__ bool digits_collide(uint8_t* tag, uint32_t d, bool valid, int lane) {
  if (valid) tag[d] = (uint8_t)lane;
  __syncwarp();
  const bool coll = valid && tag[d] != (uint8_t)lane;
  return __ballot_sync(0xffffffffu, coll) != 0u;  // also orders the tag reads before the next iteration's writes
}

// adds the digits of [wstart, wend) to the warp's private counters `my`
__device__ __forceinline__ void count_subrange(const SortArgs& a, int64_t wstart, int64_t wend, uint16_t* my,
                                               uint8_t* tag, int lane) {
  const uint32_t nb = 1u << a.bits, mask = nb - 1u;
  for (int64_t base = wstart; base < wend; base += 32 * kSortUnroll) {
    uint32_t d[kSortUnroll];
#pragma unroll
    for (int u = 0; u < kSortUnroll; ++u) {
      const int64_t i = base + u * 32 + lane;
      d[u] = i < wend ? digit_of(a.kin[i], a.shift, mask) : nb;  // nb = "no element"
    }
#pragma unroll
    for (int u = 0; u < kSortUnroll; ++u) {
      const bool valid = d[u] < nb;
      if (!digits_collide(tag, d[u], valid, lane)) {
        if (valid) my[d[u]] = (uint16_t)(my[d[u]] + 1);
      } else {
        const uint32_t m = match_by_ballots(d[u], a.bits, __ballot_sync(0xffffffffu, valid));
        if (valid && lane == __ffs(m) - 1) my[d[u]] = (uint16_t)(my[d[u]] + __popc(m));
      }
      __syncwarp();  // the next iteration may touch the same counter from another lane
    }
  }
}

// shared memory of the counting kernels: [nb] u32 offsets, [warps][nb] u16 counters, [warps][nb] u8 tags
__device__ __forceinline__ void carve_smem(uint32_t* s_mem, uint32_t nb, int warp, uint32_t** s_off, uint16_t** s_cnt,
                                           uint16_t** my, uint8_t** tag) {
  *s_off = s_mem;
  *s_cnt = reinterpret_cast<uint16_t*>(s_mem + nb);
  *my = *s_cnt + (size_t)warp * nb;
  *tag = reinterpret_cast<uint8_t*>(*s_cnt + (size_t)kSortWarps * nb) + (size_t)warp * nb;
}
static size_t sort_smem_bytes(uint32_t nb) { return (size_t)nb * (4 + 2 * kSortWarps + kSortWarps); }

// ---- hist: digit counts of one tile ----------------------------------------------------------
__global__ void __launch_bounds__(kSortThreads, 1) sort_hist_kernel(const SortArgs a) {
  extern __shared__ uint32_t s_mem[];
  const uint32_t nb = 1u << a.bits;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* s_off;
  uint16_t *s_cnt, *my;
  uint8_t* tag;
  carve_smem(s_mem, nb, warp, &s_off, &s_cnt, &my, &tag);
  for (uint32_t i = threadIdx.x; i < nb * kSortWarps / 2; i += kSortThreads) (s_mem + nb)[i] = 0u;
  __syncthreads();
  const int64_t tile = (int64_t)a.per_warp * kSortWarps;
  const int64_t wstart = (int64_t)blockIdx.x * tile + (int64_t)warp * a.per_warp;
  count_subrange(a, wstart, min(wstart + a.per_warp, a.n), my, tag, lane);
  __syncthreads();
  uint32_t* row = a.counts + (int64_t)blockIdx.x * nb;
  for (uint32_t b = threadIdx.x; b < nb; b += kSortThreads) {
    uint32_t sum = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) sum += s_cnt[(size_t)w * nb + b];
    row[b] = sum;
  }
}

// ---- scan 1: per bin, exclusive prefix over the tiles (in place) + bin totals ------------------
// block (32 bins, 32 row groups): a warp reads 32 consecutive bins of one tile row (128 bytes)
__global__ void __launch_bounds__(1024) sort_scan_cols_kernel(const SortArgs a) {
  __shared__ uint32_t s_part[32][33];
  const uint32_t nb = 1u << a.bits;
  const uint32_t col = blockIdx.x * 32 + threadIdx.x;
  const int rg = threadIdx.y;
  const int64_t rows_per = (a.ctas + 31) / 32;
  const int64_t r0 = rg * rows_per, r1 = min(r0 + rows_per, a.ctas);
  uint32_t sum = 0;
  if (col < nb)
    for (int64_t r = r0; r < r1; ++r) sum += a.counts[r * nb + col];
  s_part[rg][threadIdx.x] = sum;
  __syncthreads();
  uint32_t run = 0;
  for (int r = 0; r < rg; ++r) run += s_part[r][threadIdx.x];
  if (col < nb) {
    if (rg == 31) a.totals[col] = run + sum;
    for (int64_t r = r0; r < r1; ++r) {
      const uint32_t v = a.counts[r * nb + col];
      a.counts[r * nb + col] = run;
      run += v;
    }
  }
}

// ---- scan 2: exclusive scan of the bin totals (one CTA, nb <= 4096) ----------------------------
__global__ void __launch_bounds__(1024) sort_scan_bins_kernel(const SortArgs a) {
  __shared__ uint32_t s_warp[32];
  const uint32_t nb = 1u << a.bits;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int PER = (1 << kSortMaxBits) / 1024;  // 4 consecutive bins per thread
  uint32_t v[PER], sum = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const uint32_t b = threadIdx.x * PER + j;
    v[j] = b < nb ? a.totals[b] : 0u;
    sum += v[j];
  }
  uint32_t inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = s_warp[lane], winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    s_warp[lane] = winc - w;
  }
  __syncthreads();
  uint32_t run = s_warp[warp] + inc - sum;
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const uint32_t b = threadIdx.x * PER + j;
    if (b < nb) a.bases[b] = run;
    run += v[j];
  }
}

// ---- scatter: stable placement of one tile ---------------------------------------------------
template <bool LAST>
__global__ void __launch_bounds__(kSortThreads, 1) sort_scatter_kernel(const SortArgs a) {
  extern __shared__ uint32_t s_mem[];
  const uint32_t nb = 1u << a.bits, mask = nb - 1u;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t* s_off;   // [nb] global position of the tile's first pair per bin
  uint16_t *s_cnt, *my;
  uint8_t* tag;
  carve_smem(s_mem, nb, warp, &s_off, &s_cnt, &my, &tag);
  {
    for (uint32_t i = threadIdx.x; i < nb * kSortWarps / 2; i += kSortThreads) (s_mem + nb)[i] = 0u;
    const uint32_t* row = a.counts + (int64_t)blockIdx.x * nb;
    for (uint32_t b = threadIdx.x; b < nb; b += kSortThreads) s_off[b] = a.bases[b] + row[b];
  }
  __syncthreads();

  const int64_t tile = (int64_t)a.per_warp * kSortWarps;
  const int64_t wstart = (int64_t)blockIdx.x * tile + (int64_t)warp * a.per_warp;
  const int64_t wend = min(wstart + a.per_warp, a.n);

  // phase 1: digit counts of this warp's sub-range
  count_subrange(a, wstart, wend, my, tag, lane);
  __syncthreads();

  // phase 2: per bin, exclusive prefix over the warps (warp order = input order)
  for (uint32_t b = threadIdx.x; b < nb; b += kSortThreads) {
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = s_cnt[(size_t)w * nb + b];
      s_cnt[(size_t)w * nb + b] = (uint16_t)run;
      run += c;
    }
  }
  __syncthreads();

  // phase 3: rank and store; 32 consecutive pairs per iteration, lane order = input order
  for (int64_t base = wstart; base < wend; base += 32 * kSortUnroll) {
    uint32_t k[kSortUnroll], v[kSortUnroll];
#pragma unroll
    for (int u = 0; u < kSortUnroll; ++u) {
      const int64_t i = base + u * 32 + lane;
      const bool in = i < wend;
      k[u] = in ? a.kin[i] : 0u;
      v[u] = in ? a.vin[i] : 0u;
    }
#pragma unroll
    for (int u = 0; u < kSortUnroll; ++u) {
      const bool in = base + u * 32 + lane < wend;
      const uint32_t d = in ? digit_of(k[u], a.shift, mask) : nb;
      uint32_t pos = 0;
      if (!digits_collide(tag, d, in, lane)) {
        if (in) {
          const uint32_t c = my[d];
          my[d] = (uint16_t)(c + 1);
          pos = s_off[d] + c;
        }
      } else {
        const uint32_t m = match_by_ballots(d, a.bits, __ballot_sync(0xffffffffu, in));
        uint32_t c = 0;
        if (in) {
          c = my[d];
          pos = s_off[d] + c + (uint32_t)__popc(m & lt);
        }
        __syncwarp();  // every lane has read my[d] before the leader advances it
        if (in && lane == __ffs(m) - 1) my[d] = (uint16_t)(c + __popc(m));
      }
      __syncwarp();
      if (in) {
        if (LAST) {  // write-once results: keep them out of the way of the data the next kernels need
          __stcs(a.kout + pos, k[u]);
          __stcs(a.vout + pos, v[u]);
        } else {
          a.kout[pos] = k[u];
          a.vout[pos] = v[u];
        }
      }
    }
  }
}

static int round_up32(int64_t x) { return (int)((x + 31) / 32 * 32); }

SortShape sort_shape(int64_t n, int key_bits, int device) {
  SortShape s{};
  if (key_bits < 1) key_bits = 1;
  if (key_bits > 32) key_bits = 32;
  s.own = use_own_sort();
  if (!s.own) {
    s.passes = 1;  // one library call: a -> b
    s.bits[0] = key_bits;
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, n > 0 ? n : 1, 0, key_bits);
    s.temp_bytes = align_up(temp, 256);
    return s;
  }
  const int sms = sm_count(device);
  // tiles: a whole number of waves of one CTA per SM when the problem is large enough
  int64_t per_warp = round_up32((n + (int64_t)kSortWarps * sms - 1) / ((int64_t)kSortWarps * sms));
  if (per_warp > kSortMaxPerWarp) {
    const int64_t k = (per_warp + kSortMaxPerWarp - 1) / kSortMaxPerWarp;
    per_warp = round_up32((n + (int64_t)kSortWarps * sms * k - 1) / ((int64_t)kSortWarps * sms * k));
  }
  if (per_warp < 32 * kSortUnroll) per_warp = 32 * kSortUnroll;
  s.per_warp = (int)per_warp;
  s.ctas = (n + per_warp * kSortWarps - 1) / (per_warp * kSortWarps);
  if (s.ctas < 1) s.ctas = 1;
  // bits per pass: as few passes as the count matrix allows (<= 256 MB)
  int max_bits = kSortMaxBits;
  while (max_bits > 4 && (size_t)s.ctas * ((size_t)1 << max_bits) * 4 > ((size_t)256 << 20)) --max_bits;
  s.passes = (key_bits + max_bits - 1) / max_bits;
  if (s.passes > kSortMaxPasses) s.passes = kSortMaxPasses;  // 4 x 8 bits always suffices
  const int base = key_bits / s.passes, rem = key_bits % s.passes;
  int nb_max = 1;
  for (int p = 0; p < s.passes; ++p) {
    s.bits[p] = base + (p < rem ? 1 : 0);
    if ((1 << s.bits[p]) > nb_max) nb_max = 1 << s.bits[p];
  }
  s.temp_bytes = align_up((size_t)s.ctas * nb_max * 4, 256) + 2 * align_up((size_t)nb_max * 4, 256);
  return s;
}

int sort_pairs(uint32_t* keys_a, uint32_t* vals_a, uint32_t* keys_b, uint32_t* vals_b, int64_t n, int key_bits,
               void* temp, size_t temp_bytes, int device, cudaStream_t stream) {
  if (n <= 0) return RECEMB_OK;
  const SortShape sh = sort_shape(n, key_bits, device);
  if (temp_bytes < sh.temp_bytes) {
    set_error("sort workspace %zu < required %zu", temp_bytes, sh.temp_bytes);
    return RECEMB_ERR_WORKSPACE;
  }
  RECEMB_CHECK_ARG(((uintptr_t)temp % 256) == 0, "sort workspace must be 256-byte aligned");
  if (!sh.own) {
    size_t tb = temp_bytes;
    RECEMB_CUDA(cub::DeviceRadixSort::SortPairs(temp, tb, (const uint32_t*)keys_a, keys_b, (const uint32_t*)vals_a,
                                                vals_b, n, 0, key_bits, stream));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);  // >= 1 library kernels, counted once
    return RECEMB_OK;
  }
  int nb_max = 1;
  for (int p = 0; p < sh.passes; ++p) nb_max = nb_max > (1 << sh.bits[p]) ? nb_max : (1 << sh.bits[p]);
  static bool attr_set[64] = {};
  const int d = (device >= 0 && device < 64) ? device : 0;
  if (!attr_set[d]) {
    const int max_smem = (int)sort_smem_bytes(1u << kSortMaxBits);
    RECEMB_CUDA(cudaFuncSetAttribute(sort_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    RECEMB_CUDA(cudaFuncSetAttribute(sort_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    RECEMB_CUDA(cudaFuncSetAttribute(sort_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    attr_set[d] = true;
  }
  SortArgs a;
  a.n = n;
  a.per_warp = sh.per_warp;
  a.ctas = sh.ctas;
  a.counts = (uint32_t*)temp;
  a.totals = (uint32_t*)((char*)temp + align_up((size_t)sh.ctas * nb_max * 4, 256));
  a.bases = (uint32_t*)((char*)a.totals + align_up((size_t)nb_max * 4, 256));
  bool in_b = sort_input_in_b(sh);
  int shift = 0;
  for (int p = 0; p < sh.passes; ++p) {
    a.kin = in_b ? keys_b : keys_a;
    a.vin = in_b ? vals_b : vals_a;
    a.kout = in_b ? keys_a : keys_b;
    a.vout = in_b ? vals_a : vals_b;
    a.shift = shift;
    a.bits = sh.bits[p];
    const uint32_t nb = 1u << a.bits;
    const size_t smem = sort_smem_bytes(nb);
    sort_hist_kernel<<<(unsigned)sh.ctas, kSortThreads, smem, stream>>>(a);
    RECEMB_LAUNCHED();
    sort_scan_cols_kernel<<<(nb + 31) / 32, dim3(32, 32), 0, stream>>>(a);
    RECEMB_LAUNCHED();
    sort_scan_bins_kernel<<<1, 1024, 0, stream>>>(a);
    RECEMB_LAUNCHED();
    if (p == sh.passes - 1) sort_scatter_kernel<true><<<(unsigned)sh.ctas, kSortThreads, smem, stream>>>(a);
    else sort_scatter_kernel<false><<<(unsigned)sh.ctas, kSortThreads, smem, stream>>>(a);
    RECEMB_LAUNCHED();
    shift += a.bits;
    in_b = !in_b;
  }
  return RECEMB_OK;
}

}  // namespace recemb
