// Id production on the GPU: the feeder of the embedding path.
//   recemb_xxh64_ids      strings -> signed 64-bit ids, bit-exact xxhash XXH64(str, seed) - 2^63
//                         (commons/feature_utils.py:40-46 hash_string_to_long; per-value Python
//                         call at :141-146 in the reference)
//   recemb_pad_histories  ragged id lists -> [B, L] truncated / right-padded with 0, optionally
//                         dropping the row's own target id (commons/feature_utils.py:21-25 pad_array,
//                         :149-179 handle_categorical_history_feature; a per-row Python loop there)
#include "common.cuh"

namespace recemb {

constexpr uint64_t P1 = 11400714785074694791ull, P2 = 14029467366897019727ull, P3 = 1609587929392839161ull,
                   P4 = 9650029242287828579ull, P5 = 2870177450012600261ull;

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
__device__ __forceinline__ uint64_t xxh_round(uint64_t acc, uint64_t in) { return rotl64(acc + in * P2, 31) * P1; }
__device__ __forceinline__ uint64_t xxh_merge(uint64_t acc, uint64_t v) { return (acc ^ xxh_round(0, v)) * P1 + P4; }

// little-endian reads from an arbitrarily aligned byte string, with optional ASCII lower-casing
struct ByteReader {
  const uint8_t* p;
  bool lower;
  __device__ __forceinline__ uint64_t byte(int64_t i) const {
    uint8_t c = p[i];
    if (lower && c >= 'A' && c <= 'Z') c += 32;
    return c;
  }
  __device__ __forceinline__ uint64_t u64(int64_t i) const {
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v |= byte(i + k) << (8 * k);
    return v;
  }
  __device__ __forceinline__ uint64_t u32(int64_t i) const {
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) v |= byte(i + k) << (8 * k);
    return v;
  }
};

__device__ uint64_t xxh64(const ByteReader& r, int64_t len, uint64_t seed) {
  int64_t i = 0;
  uint64_t h;
  if (len >= 32) {
    uint64_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    for (; i + 32 <= len; i += 32) {
      v1 = xxh_round(v1, r.u64(i));
      v2 = xxh_round(v2, r.u64(i + 8));
      v3 = xxh_round(v3, r.u64(i + 16));
      v4 = xxh_round(v4, r.u64(i + 24));
    }
    h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
    h = xxh_merge(h, v1);
    h = xxh_merge(h, v2);
    h = xxh_merge(h, v3);
    h = xxh_merge(h, v4);
  } else {
    h = seed + P5;
  }
  h += (uint64_t)len;
  for (; i + 8 <= len; i += 8) {
    h ^= xxh_round(0, r.u64(i));
    h = rotl64(h, 27) * P1 + P4;
  }
  if (i + 4 <= len) {
    h ^= r.u32(i) * P1;
    h = rotl64(h, 23) * P2 + P3;
    i += 4;
  }
  for (; i < len; ++i) {
    h ^= r.byte(i) * P5;
    h = rotl64(h, 11) * P1;
  }
  h ^= h >> 33;
  h *= P2;
  h ^= h >> 29;
  h *= P3;
  h ^= h >> 32;
  return h;
}

__global__ void __launch_bounds__(256) xxh64_ids_kernel(const uint8_t* __restrict__ bytes,
                                                       const int64_t* __restrict__ offsets, int64_t n,
                                                       uint64_t seed, int to_lower, int64_t* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * 256;
  for (; i < n; i += stride) {
    const int64_t b = offsets[i], e = offsets[i + 1];
    ByteReader r{bytes + b, to_lower != 0};
    // xxh64 in [0, 2^64)  minus 2^63  ==  the digest with its top bit flipped, read as int64
    out[i] = (int64_t)(xxh64(r, e - b, seed) ^ 0x8000000000000000ull);
  }
}

// one warp per row: keep the first L elements of the row that differ from the row's target id
__global__ void __launch_bounds__(256) pad_histories_kernel(const int64_t* __restrict__ values,
                                                           const int64_t* __restrict__ offsets,
                                                           const int64_t* __restrict__ remove_ids, int64_t rows,
                                                           int32_t L, int64_t pad_token,
                                                           int64_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  int64_t row = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int64_t stride = ((int64_t)gridDim.x * 256) >> 5;
  for (; row < rows; row += stride) {
    const int64_t b = offsets[row], e = offsets[row + 1];
    const bool filt = remove_ids != nullptr;
    const int64_t target = filt ? remove_ids[row] : 0;
    int kept = 0;
    for (int64_t base = b; base < e && kept < L; base += 32) {
      const int64_t i = base + lane;
      const int64_t v = i < e ? values[i] : 0;
      const bool keep = i < e && !(filt && v == target);
      const uint32_t m = __ballot_sync(0xffffffffu, keep);
      const int pos = kept + __popc(m & ((1u << lane) - 1u));
      if (keep && pos < L) out[row * L + pos] = v;
      kept += __popc(m);
    }
    if (kept > L) kept = L;
    for (int p = kept + lane; p < L; p += 32) out[row * L + p] = pad_token;
  }
}

}  // namespace recemb

using namespace recemb;

extern "C" int recemb_xxh64_ids(const uint8_t* bytes, const int64_t* offsets, int64_t n, uint64_t seed,
                                int to_lower, int64_t* ids_out, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n >= 0, "n < 0");
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(offsets && ids_out, "null pointer");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  int64_t grid = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count(device) * 16;
  if (grid > cap) grid = cap;
  xxh64_ids_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(bytes, offsets, n, seed, to_lower, ids_out);
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

extern "C" int recemb_pad_histories(const int64_t* values, const int64_t* offsets, const int64_t* remove_ids,
                                    int64_t rows, int32_t history_length, int64_t pad_token, int64_t* out,
                                    int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(rows >= 0 && history_length >= 1, "bad rows / history_length");
  if (rows == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(offsets && out, "null pointer");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  int64_t grid = (rows * 32 + 255) / 256;
  const int64_t cap = (int64_t)sm_count(device) * 16;
  if (grid > cap) grid = cap;
  pad_histories_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(values, offsets, remove_ids, rows,
                                                                        history_length, pad_token, out);
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}
