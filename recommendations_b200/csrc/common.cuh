// Shared device/host helpers for the recemb_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/recemb_b200.h"

namespace recemb {

// ---------------------------------------------------------------- errors ----
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;

#define RECEMB_CHECK_ARG(cond, ...)         \
  do {                                      \
    if (!(cond)) {                          \
      ::recemb::set_error(__VA_ARGS__);     \
      return RECEMB_ERR_INVALID;            \
    }                                       \
  } while (0)

#define RECEMB_UNSUPPORTED(cond, ...)       \
  do {                                      \
    if (!(cond)) {                          \
      ::recemb::set_error(__VA_ARGS__);     \
      return RECEMB_ERR_UNSUPPORTED;        \
    }                                       \
  } while (0)

#define RECEMB_CUDA(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::recemb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr,                 \
                          cudaGetErrorString(_e));                                     \
      return RECEMB_ERR_CUDA;                                                          \
    }                                                                                  \
  } while (0)

// Every kernel launch goes through this so recemb_launch_count() is honest.
#define RECEMB_LAUNCHED()                                        \
  do {                                                           \
    ::recemb::g_launch_count.fetch_add(1, std::memory_order_relaxed); \
    RECEMB_CUDA(cudaGetLastError());                             \
  } while (0)

// Sets the device for the calling thread for the lifetime of the guard (the
// backward pass arrives on the autograd engine thread with arbitrary state).
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int device) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != device) {
      err = cudaSetDevice(device);
      changed = (err == cudaSuccess);
    }
  }
  ~DeviceGuard() {
    if (changed) cudaSetDevice(prev);
  }
};

int sm_count(int device);

// --------------------------------------------------- exact 64-bit modulo ----
// Barrett reduction by a runtime-constant divisor n (1 <= n < 2^63):
// magic = floor((2^64-1)/n); q_est = mulhi(u, magic) is in [q-2, q].
struct ModN {
  uint64_t n;
  uint64_t magic;
};

inline ModN make_modn(uint64_t n) {
  ModN m;
  m.n = n;
  m.magic = n ? (~0ull) / n : 0ull;
  return m;
}

__device__ __forceinline__ uint64_t udivmod(uint64_t u, const ModN& m, uint64_t* q_out) {
  uint64_t q = __umul64hi(u, m.magic);
  uint64_t r = u - q * m.n;
  while (r >= m.n) {  // at most twice
    r -= m.n;
    ++q;
  }
  if (q_out) *q_out = q;
  return r;
}

// Python / torch.remainder semantics: result in [0, n) for any signed x.
__device__ __forceinline__ int64_t floor_mod(int64_t x, const ModN& m) {
  if (x >= 0) return (int64_t)udivmod((uint64_t)x, m, nullptr);
  // x < 0: u = -x-1 >= 0, x = -(q+1) n + (n-1-s) with s = u mod n
  uint64_t s = udivmod(~(uint64_t)x, m, nullptr);
  return (int64_t)(m.n - 1 - s);
}

struct HashSpec {
  int mode;      // recemb_hash
  int shift;     // ROTL_FLOORMOD: rotation c
  ModN mod_rows; // FLOORMOD / ROTL: num_rows ; QR: d
  ModN mod_sq;   // QR: d*d
  // table-batched lookups: lookup i belongs to table t = i / ids_per_table and its row is
  // offset by t * rows_per_table into one stacked [T * rows_per_table, dim] table (0 = off)
  uint32_t ids_per_table;
  uint32_t num_tables;    // > 0: table index wraps modulo num_tables
  int64_t rows_per_table; // LOCAL rows of one table
  // row-wise sharding: global row r lives on rank r % shard_world at local row r / shard_world;
  // a lookup whose row belongs to another rank is dropped (shard_world <= 1 = off)
  uint32_t shard_world;
  uint32_t shard_rank;
  ModN mod_world;
  uint32_t flip_len;  // > 0: outputs / gradient rows mirrored inside sequences of flip_len lookups
  // sequence window (query_tower.py:73-86, the batch-wide trim): sequences of win_len lookups of which only
  // *win_keep (DEVICE scalar, read at kernel start) positions are looked up -- the first (win_side 0) or the
  // last (win_side 1) ones -- and written compactly as [n / win_len, *win_keep, dim]
  uint32_t win_len;
  uint32_t win_side;
  const int32_t* win_keep;
  // pooled bags written feature-interleaved into [bags_per_table, out_feats, dim] (0 = bag order)
  uint32_t out_feats;
  uint32_t out_feat_off;
  uint32_t out_bpt;  // bags per table
  uint32_t partition;  // peer bucket: 1 = table-wise (table t whole on rank t % shard_world)
};

// output / gradient row of pooled bag g
__device__ __forceinline__ int64_t bag_out_row(int64_t g, const HashSpec& h) {
  if (!h.out_feats) return g;
  const int64_t t = g / h.out_bpt;
  return (g - t * h.out_bpt) * h.out_feats + t + h.out_feat_off;
}

// position i mirrored inside its sequence of L lookups (L == 0: unchanged)
__device__ __forceinline__ int64_t flip_index(int64_t i, uint32_t L) {
  if (!L) return i;
  const int64_t b = i / L;
  return b * L + (L - 1 - (i - b * L));
}

__device__ __forceinline__ uint32_t window_keep(const HashSpec& h) {
  if (!h.win_len) return 0;
  const int32_t k = *h.win_keep;
  return (uint32_t)min(max(k, 0), (int32_t)h.win_len);
}

// output / gradient row of lookup i: mirrored inside its sequence (flip_len), compacted to the kept
// window (win_len; keep from window_keep()); -1 = outside the window, neither read nor written
__device__ __forceinline__ int64_t out_row(int64_t i, const HashSpec& h, uint32_t keep) {
  if (!h.win_len) return flip_index(i, h.flip_len);
  const uint32_t L = h.win_len;
  const int64_t b = i / L;
  uint32_t p = (uint32_t)(i - b * L);
  if (h.win_side) {
    if (p < L - keep) return -1;
    p -= L - keep;
  } else if (p >= keep) {
    return -1;
  }
  if (h.flip_len) p = keep - 1 - p;
  return b * keep + p;
}

int make_hash_spec(int hash_mode, int64_t num_rows, int64_t hash_arg, HashSpec* out,
                   const recemb_layout* layout = nullptr);
int64_t layout_local_rows(int64_t num_rows, const recemb_layout* layout);
int64_t layout_tables(const recemb_layout* layout, int64_t n_ids);

// global row -> local row on this shard, or -1 when another rank owns it
__device__ __forceinline__ int64_t shard_local_row(int64_t row, const HashSpec& h) {
  if (h.shard_world <= 1 || row < 0) return row;
  uint64_t q;
  const uint64_t r = udivmod((uint64_t)row, h.mod_world, &q);
  return r == h.shard_rank ? (int64_t)q : -1;
}

__device__ __forceinline__ int64_t table_offset(int64_t i, const HashSpec& h) {
  if (!h.ids_per_table) return 0;
  uint32_t t = (uint32_t)i / h.ids_per_table;
  if (h.num_tables) t %= h.num_tables;
  return (int64_t)t * h.rows_per_table;
}

// (x << c) | (x >> (64-c)) with wrapping << and ARITHMETIC >> on signed int64:
// a true rotate only for x >= 0 (commons/layers.py:182).
__device__ __forceinline__ int64_t signed_rotl(int64_t x, int c) {
  if (c == 0) return x;
  return (int64_t)(((uint64_t)x << c) | (uint64_t)(x >> (64 - c)));
}

__device__ __forceinline__ int64_t row_of(int64_t id, const HashSpec& h) {
  switch (h.mode) {
    case RECEMB_HASH_IDENTITY:
      // ids ARE rows: anything outside [0, num_rows) is dropped (-1) by every kernel instead of
      // reading / updating out of bounds (nn.EmbeddingBag raises there; the host layer checks)
      return ((uint64_t)id < h.mod_rows.n) ? id : -1;
    case RECEMB_HASH_FLOORMOD:
      return floor_mod(id, h.mod_rows);
    case RECEMB_HASH_ROTL_FLOORMOD:
      return floor_mod(signed_rotl(id, h.shift), h.mod_rows);
    case RECEMB_HASH_QR_QUOTIENT: {
      uint64_t x = (uint64_t)floor_mod(id, h.mod_sq);
      uint64_t q;
      udivmod(x, h.mod_rows, &q);  // floor(x / d), already in [0, d)
      return (int64_t)q;
    }
    case RECEMB_HASH_DIV_FLOORMOD: {
      // floor_mod(floor_divide(id, div), num_rows), div = mod_sq.n  (PatternFromTimelocal)
      uint64_t q;
      int64_t fq;
      if (id >= 0) {
        udivmod((uint64_t)id, h.mod_sq, &q);
        fq = (int64_t)q;
      } else {  // floor(x / d) = -floor((-x - 1) / d) - 1 for x < 0
        udivmod(~(uint64_t)id, h.mod_sq, &q);
        fq = -(int64_t)q - 1;
      }
      return floor_mod(fq, h.mod_rows);
    }
    default: {  // RECEMB_HASH_QR_REMAINDER
      uint64_t x = (uint64_t)floor_mod(id, h.mod_sq);
      return (int64_t)udivmod(x, h.mod_rows, nullptr);
    }
  }
}

__device__ __forceinline__ int64_t kshift_row(int64_t id, int c, const ModN& m) {
  return floor_mod(signed_rotl(id, c), m);
}

// ------------------------------------------------------ peer flag helpers ----
__device__ __forceinline__ void st_release_sys_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ uint64_t ld_relaxed_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Fused gradient push of the sharded backward (bwd.cu: the first `push_ctas` CTAs of the level-0 launch are
// the all-gather push, table by table; the others are the segmented reduction, every chunk gated on the
// arrival count of the last table it touches).  Gate region of an arena: step u64 | done u32[kGateTables] |
// arrival counts u64[kGateTables] (+ reserve) (recemb_peer_arena.off_gate).
constexpr int kGateTables = 64;
constexpr int64_t kGateOffDone = 64;
constexpr int64_t kGateOffFlags = kGateOffDone + kGateTables * 4;
constexpr int64_t kGateBytes = kGateOffFlags + (int64_t)kGateTables * RECEMB_MAX_PEERS * 8;

struct PeerGate {
  int32_t push_ctas;  // 0 = no fused push: plain segmented reduction
  int32_t world, rank;
  int32_t tables;
  char* arena[RECEMB_MAX_PEERS];
  int64_t off_grads;
  int64_t off_gate;
  const uint4* src;          // my pooled gradients [tables][bags_per_table][row]
  int64_t vecs_per_table;    // bags_per_table * row_vecs
  int64_t sender_vecs;       // bags_total * row_vecs: stride between two senders' slices
  uint32_t rows_per_table;   // local rows of one table in my stacked shard
  long long timeout_cycles;
  uint32_t* status;
  int32_t partition;     // 1 = table-wise: table t goes to rank t % world only, gating its local table t / world
  int32_t local_tables;  // tables in MY shard
  int debug;  // RECEMB_GATE_DEBUG (timing experiments only): 1 = reduction does not wait, 2 = pushers do not copy
};

// ------------------------------------------------------- memory helpers ----
// 128-bit read-only load that does not allocate in L1 (table rows are touched
// once per CTA; L2 keeps whatever reuse exists).
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// 128-bit read-only load that MAY stay in L1: for rows a whole batch keeps hitting (the k-shift
// collapse rows [N - 2^(c-1), N), Zipf heads) -- served from every SM's L1 instead of all SMs
// queueing on the one or two L2 slices that hold the row
__device__ __forceinline__ uint4 ldg_nc_l1_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// streaming (evict-first) 128-bit store for write-once outputs
__device__ __forceinline__ void stg_cs_v4(void* p, const uint4& v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 ldg_v4(const void* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void stg_v4(void* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }

// A 16-byte vector viewed as fp32 lanes for a given storage dtype.
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int kElems = 4;
  __device__ __forceinline__ static void unpack(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x);
    f[1] = __uint_as_float(v.y);
    f[2] = __uint_as_float(v.z);
    f[3] = __uint_as_float(v.w);
  }
  __device__ __forceinline__ static uint4 pack(const float* f) {
    return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]),
                      __float_as_uint(f[3]));
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int kElems = 8;
  __device__ __forceinline__ static void unpack(const uint4& v, float* f) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ static uint4 pack(const float* f) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 p = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&p);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
  }
};

// sum over the G (power of two, <= 32) consecutive lanes that share a row
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------ mbarrier / bulk copy ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA engine 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes,
                                         uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// shared -> global bulk store (bulk async-group completion)
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace recemb
