// Backward of the embedding hot path (sm_100a):
//   plan   : lookup slots -> (row, slot) pairs sorted by row, stable in slot
//            (sort-based dedup; replaces the sort inside ATen's
//            embedding_dense_backward that autograd runs for
//            embedding_module_gen.py:113, :152)
//   apply  : chunked segmented reduction over the sorted pairs + fused
//            optimizer update of the touched rows only (replaces the dense
//            [N, D] gradient + torch.optim.Adagrad full-table pass,
//            embedding_module_gen.py:97, :137, :153)
//
// The segmented reduction is load-balanced by construction: every group of G
// lanes owns a chunk of kChunk consecutive sorted entries whatever the run
// lengths are (the k-shift collapse puts ~50 % of a shift's lookups on a
// handful of rows).  Runs closed inside a chunk are applied
// directly; runs that cross a chunk boundary leave (row, partial sum) records
// that the next level reduces with the same kernel, until one chunk is left.
// Summation order is fixed (sorted order, then chunk order): deterministic, no
// atomics.
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace recemb {

constexpr int kBwdThreads = 256;
constexpr int kChunk = 32;
constexpr uint32_t kNoKey = 0xffffffffu;
constexpr size_t kCounterBytes = 256;

// --------------------------------------------------------------- layout ----
struct PlanLayout {
  int64_t n;
  int key_bits;
  size_t off_keys_in, off_vals_in, off_keys_out, off_vals_out, off_temp, temp_bytes, total;
};

static int bit_width_u64(uint64_t x) {
  int b = 0;
  while (x) {
    ++b;
    x >>= 1;
  }
  return b;
}

static cudaError_t plan_layout(int64_t n, int64_t num_rows, PlanLayout* L) {
  L->n = n;
  L->key_bits = bit_width_u64((uint64_t)num_rows);  // the sentinel key == num_rows must sort last
  if (L->key_bits < 1) L->key_bits = 1;
  size_t temp = 0;
  cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, temp, (const uint32_t*)nullptr,
                                                  (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                                  (uint32_t*)nullptr, (int64_t)n, 0, L->key_bits);
  if (e != cudaSuccess) return e;
  const size_t arr = align_up((size_t)n * 4, 256);
  size_t off = kCounterBytes;
  L->off_keys_in = off;
  off += arr;
  L->off_vals_in = off;
  off += arr;
  L->off_keys_out = off;
  off += arr;
  L->off_vals_out = off;
  off += arr;
  L->off_temp = off;
  L->temp_bytes = temp;
  off += align_up(temp, 256);
  L->total = off;
  return cudaSuccess;
}

// ------------------------------------------------------------- plan keys ----
struct PlanKeyArgs {
  const int64_t* ids;
  int64_t n_slots;
  int32_t slots_per_id;
  HashSpec h;
  int zero_pad;
  int64_t pad_id;
  int64_t pad_row;
  int32_t bag_size;
  const int32_t* lengths;
  int32_t last_n;
  uint32_t sentinel;
  uint32_t* keys;
  uint32_t* vals;
};

__global__ void __launch_bounds__(kBwdThreads) plan_keys_kernel(const PlanKeyArgs a) {
  int64_t s = (int64_t)blockIdx.x * kBwdThreads + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * kBwdThreads;
  for (; s < a.n_slots; s += stride) {
    int64_t id_idx = s;
    int c = 0;
    if (a.slots_per_id > 1) {
      id_idx = s / a.slots_per_id;
      c = (int)(s - id_idx * a.slots_per_id);
    }
    const int64_t id = a.ids[id_idx];
    bool ok = !(a.zero_pad && id == a.pad_id);
    if (ok && a.bag_size > 0) {
      const int64_t bag = id_idx / a.bag_size;
      const int p = (int)(id_idx - bag * a.bag_size);
      int hi = a.bag_size;
      if (a.lengths) hi = min(max(a.lengths[bag], 0), a.bag_size);
      const int lo = a.last_n > 0 ? max(0, hi - a.last_n) : 0;
      ok = p >= lo && p < hi;
    }
    uint32_t key = a.sentinel;
    if (ok) {
      const int64_t row =
          a.slots_per_id > 1 ? kshift_row(id, c, a.h.mod_rows) : row_of(id, a.h);
      if (row != a.pad_row) key = (uint32_t)row;
    }
    a.keys[s] = key;
    a.vals[s] = (uint32_t)s;
  }
}

// counters[0] = valid slots, counters[1] = distinct rows
__global__ void __launch_bounds__(kBwdThreads) plan_count_kernel(const uint32_t* __restrict__ keys,
                                                                 int64_t n, uint32_t sentinel,
                                                                 unsigned long long* counters) {
  __shared__ unsigned int s_valid, s_heads;
  if (threadIdx.x == 0) s_valid = s_heads = 0;
  __syncthreads();
  unsigned int valid = 0, heads = 0;
  int64_t i = (int64_t)blockIdx.x * kBwdThreads + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * kBwdThreads;
  for (; i < n; i += stride) {
    const uint32_t k = keys[i];
    if (k < sentinel) {
      ++valid;
      if (i == 0 || keys[i - 1] != k) ++heads;
    }
  }
  valid = __reduce_add_sync(0xffffffffu, valid);
  heads = __reduce_add_sync(0xffffffffu, heads);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_valid, valid);
    atomicAdd(&s_heads, heads);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicAdd(&counters[0], (unsigned long long)s_valid);
    atomicAdd(&counters[1], (unsigned long long)s_heads);
  }
}

// ---------------------------------------------------- segmented reduction ----
// Rows are handled as "quads" of 4 consecutive elements (16 B of fp32 / 8 B of
// bf16): G lanes x V quads per lane cover a row.
template <typename T>
__device__ __forceinline__ void load_quad(const T* p, float* f);
template <>
__device__ __forceinline__ void load_quad<float>(const float* p, float* f) {
  const uint4 v = ldg_v4(p);
  Vec16<float>::unpack(v, f);
}
template <>
__device__ __forceinline__ void load_quad<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
}
// gradients / partial sums are streamed exactly once: keep them out of L1
template <typename T>
__device__ __forceinline__ void load_quad_stream(const T* p, float* f);
template <>
__device__ __forceinline__ void load_quad_stream<float>(const float* p, float* f) {
  const uint4 v = ldg_nc_v4(p);
  Vec16<float>::unpack(v, f);
}
template <>
__device__ __forceinline__ void load_quad_stream<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
}
template <typename T>
__device__ __forceinline__ void store_quad(T* p, const float* f);
template <>
__device__ __forceinline__ void store_quad<float>(float* p, const float* f) {
  stg_v4(p, Vec16<float>::pack(f));
}
template <>
__device__ __forceinline__ void store_quad<__nv_bfloat16>(__nv_bfloat16* p, const float* f) {
  __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
  uint2 v;
  v.x = *reinterpret_cast<uint32_t*>(&a);
  v.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = v;
}

struct SegArgs {
  const uint32_t* keys;   // sorted rows (level 0) or record rows (level >= 1)
  const uint32_t* slots;  // level 0 only
  int64_t n;              // entries at this level
  const void* grad;       // level 0: [grad_rows, dim] GT ; level >= 1: fp32 partials [n, dim]
  int32_t dim;
  int32_t quads;          // dim / 4
  int32_t slots_per_grad_row;
  const float* slot_weight;
  const float* grad_row_scale;
  uint32_t sentinel;      // keys >= sentinel carry nothing
  int update;
  void* table;
  float* state1;
  float* state2;
  recemb_optim_params hp;
  uint32_t* out_keys;     // [2 * chunks]
  float* out_partials;    // [2 * chunks, dim]
};

template <int G>
__device__ __forceinline__ float masked_group_sum(float v, uint32_t mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

template <int G, int V, typename WT>
__device__ __forceinline__ void apply_row(const SegArgs& a, uint32_t row, float (&g)[V][4], int lig,
                                          uint32_t gmask) {
  WT* wrow = reinterpret_cast<WT*>(a.table) + (int64_t)row * a.dim;
  const recemb_optim_params& hp = a.hp;
  if (a.update == RECEMB_UPD_DENSE_GRAD) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int q = j * G + lig;
      if (q < a.quads) store_quad<WT>(wrow + q * 4, g[j]);
    }
    return;
  }
  float w[V][4];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int q = j * G + lig;
#pragma unroll
    for (int e = 0; e < 4; ++e) w[j][e] = 0.f;
    if (q < a.quads) load_quad<WT>(wrow + q * 4, w[j]);
  }
  const bool l2_decay = hp.weight_decay != 0.f && a.update != RECEMB_UPD_ADAMW;
  if (l2_decay) {
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) g[j][e] += hp.weight_decay * w[j][e];
  }
  if (a.update == RECEMB_UPD_SGD) {
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) w[j][e] -= hp.lr * g[j][e];
  } else if (a.update == RECEMB_UPD_ADAGRAD) {
    float* srow = a.state1 + (int64_t)row * a.dim;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int q = j * G + lig;
      if (q < a.quads) {
        float s[4];
        load_quad<float>(srow + q * 4, s);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          s[e] += g[j][e] * g[j][e];
          w[j][e] += (-hp.lr * g[j][e]) / (sqrtf(s[e]) + hp.eps);
        }
        store_quad<float>(srow + q * 4, s);
      }
    }
  } else if (a.update == RECEMB_UPD_ROWWISE_ADAGRAD) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) ss += g[j][e] * g[j][e];  // lanes past the row hold zeros
    ss = masked_group_sum<G>(ss, gmask) / (float)a.dim;
    const float s_new = a.state1[row] + ss;
    const float denom = sqrtf(s_new) + hp.eps;
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) w[j][e] += (-hp.lr * g[j][e]) / denom;
    __syncwarp(gmask);  // every lane has read state1[row] before lane 0 overwrites it
    if (lig == 0) a.state1[row] = s_new;
  } else {  // ADAM / ADAMW, lazy: only touched rows move
    float* mrow = a.state1 + (int64_t)row * a.dim;
    float* vrow = a.state2 + (int64_t)row * a.dim;
    const float step_size = hp.lr / hp.bias_correction1;
    const float bc2_sqrt = sqrtf(hp.bias_correction2);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int q = j * G + lig;
      if (q < a.quads) {
        float m[4], v[4];
        load_quad<float>(mrow + q * 4, m);
        load_quad<float>(vrow + q * 4, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (a.update == RECEMB_UPD_ADAMW) w[j][e] *= (1.f - hp.lr * hp.weight_decay);
          m[e] = hp.beta1 * m[e] + (1.f - hp.beta1) * g[j][e];
          v[e] = hp.beta2 * v[e] + (1.f - hp.beta2) * g[j][e] * g[j][e];
          const float denom = sqrtf(v[e]) / bc2_sqrt + hp.eps;
          w[j][e] -= step_size * (m[e] / denom);
        }
        store_quad<float>(mrow + q * 4, m);
        store_quad<float>(vrow + q * 4, v);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j) {
    const int q = j * G + lig;
    if (q < a.quads) store_quad<WT>(wrow + q * 4, w[j]);
  }
}

template <int G, int V, typename GT, typename WT, bool L0>
__global__ void __launch_bounds__(kBwdThreads, (V == 1) ? 3 : (V == 2 ? 2 : 1)) seg_kernel(const SegArgs a) {
  constexpr int BATCH = (V == 1) ? 8 : (V == 2 ? 4 : 2);
  constexpr int GROUPS = kBwdThreads / G;
  const int lane = threadIdx.x & 31;
  const int lig = lane % G;
  const int gi_warp = lane / G;
  const uint32_t gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (gi_warp * G));
  const int64_t chunk = (int64_t)blockIdx.x * GROUPS + threadIdx.x / G;
  const int64_t start = chunk * kChunk;
  if (start >= a.n) return;
  const int64_t end = min(start + (int64_t)kChunk, a.n);

  if (lig == 0) {
    a.out_keys[2 * chunk] = kNoKey;
    a.out_keys[2 * chunk + 1] = kNoKey;
  }

  uint32_t cur = a.keys[start];
  const bool left_open = start > 0 && a.keys[start - 1] == cur;
  bool first = true;
  float acc[V][4];
#pragma unroll
  for (int j = 0; j < V; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;

  auto flush = [&](uint32_t key, bool leading, bool trailing) {
    if (key >= a.sentinel) return;
    if (!leading && !trailing) {
      apply_row<G, V, WT>(a, key, acc, lig, gmask);
      return;
    }
    const int64_t rec = 2 * chunk + (leading ? 0 : 1);
    float* dst = a.out_partials + rec * a.dim;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int q = j * G + lig;
      if (q < a.quads) store_quad<float>(dst + q * 4, acc[j]);
    }
    if (lig == 0) a.out_keys[rec] = key;
    if (leading && trailing) {  // the whole chunk is one run open on both sides
      float z[4] = {0.f, 0.f, 0.f, 0.f};
      float* dst2 = a.out_partials + (rec + 1) * a.dim;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const int q = j * G + lig;
        if (q < a.quads) store_quad<float>(dst2 + q * 4, z);
      }
      if (lig == 0) a.out_keys[rec + 1] = key;
    }
  };

  for (int64_t e0 = start; e0 < end; e0 += BATCH) {
    uint32_t k[BATCH];
    float g[BATCH][V][4];
    float wt[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int64_t i = e0 + u;
      k[u] = (i < end) ? a.keys[i] : kNoKey;
      wt[u] = 1.f;
#pragma unroll
      for (int j = 0; j < V; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) g[u][j][e] = 0.f;
      if (k[u] < a.sentinel) {
        int64_t grow;
        if (L0) {
          const uint32_t slot = a.slots[i];
          grow = (a.slots_per_grad_row > 1) ? (int64_t)(slot / (uint32_t)a.slots_per_grad_row)
                                            : (int64_t)slot;
          if (a.slot_weight) wt[u] = a.slot_weight[slot];
          if (a.grad_row_scale) wt[u] *= a.grad_row_scale[grow];
        } else {
          grow = i;
        }
        const GT* src = reinterpret_cast<const GT*>(a.grad) + grow * a.dim;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int q = j * G + lig;
          if (q < a.quads) load_quad_stream<GT>(src + q * 4, g[u][j]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      if (e0 + u < end) {
        if (k[u] != cur) {
          flush(cur, first && left_open, false);
          first = false;
          cur = k[u];
#pragma unroll
          for (int j = 0; j < V; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
        }
        if (k[u] < a.sentinel) {
          if (L0 && (a.slot_weight || a.grad_row_scale)) {
#pragma unroll
            for (int j = 0; j < V; ++j)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[j][e] += wt[u] * g[u][j][e];
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[j][e] += g[u][j][e];
          }
        }
      }
    }
  }
  const bool right_open = end < a.n && a.keys[end] == cur;
  flush(cur, first && left_open, right_open);
}

struct QuadShape {
  int G, V;
};
static bool pick_quads(int quads, QuadShape* s) {
  int g = 1;
  while (g < 32 && g < quads) g <<= 1;
  int v = (quads + g - 1) / g;
  int vp = 1;
  while (vp < v) vp <<= 1;
  if (vp > 8) return false;
  s->G = g;
  s->V = vp;
  return true;
}

#define SEG_GV(G_, V_)                                                                  \
  if (shape.G == G_ && shape.V == V_) {                                                 \
    const int64_t chunks = (a.n + kChunk - 1) / kChunk;                                 \
    const int groups = kBwdThreads / G_;                                                \
    const int64_t grid = (chunks + groups - 1) / groups;                                \
    seg_kernel<G_, V_, GT, WT, L0><<<(unsigned)grid, kBwdThreads, 0, s>>>(a);           \
    launched = true;                                                                    \
  }

template <typename GT, typename WT, bool L0>
static int launch_seg(const SegArgs& a, QuadShape shape, cudaStream_t s) {
  bool launched = false;
  SEG_GV(1, 1) SEG_GV(2, 1) SEG_GV(4, 1) SEG_GV(8, 1) SEG_GV(16, 1) SEG_GV(32, 1) SEG_GV(32, 2)
  SEG_GV(32, 4) SEG_GV(32, 8)
  if (!launched) {
    set_error("bwd_apply: no kernel for G=%d V=%d", shape.G, shape.V);
    return RECEMB_ERR_UNSUPPORTED;
  }
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

// level sizes: n_0 = n_slots, n_{l+1} = 2 * ceil(n_l / kChunk), until n_l <= kChunk
static int level_sizes(int64_t n0, int64_t* sizes, int max_levels) {
  int L = 0;
  int64_t n = n0;
  sizes[L++] = n;
  while (n > kChunk && L < max_levels) {
    n = 2 * ((n + kChunk - 1) / kChunk);
    sizes[L++] = n;
  }
  return L;
}
constexpr int kMaxLevels = 16;

// measurement hook: events recorded around the level-0 launch of the next bwd_apply
static thread_local cudaEvent_t t_ev_start = nullptr, t_ev_stop = nullptr;

// ------------------------------------------------------- epilogue backward ----
template <typename T>
__global__ void __launch_bounds__(kBwdThreads)
    epilogue_bwd_kernel(const T* __restrict__ grad_out, const T* __restrict__ out,
                        const float* __restrict__ inv_norm, int64_t n, int32_t dim, int epilogue,
                        float sqrt_k, float* __restrict__ dx) {
  // one warp per row; fp32 math
  const int lane = threadIdx.x & 31;
  int64_t row = ((int64_t)blockIdx.x * kBwdThreads + threadIdx.x) >> 5;
  const int64_t stride = ((int64_t)gridDim.x * kBwdThreads) >> 5;
  for (; row < n; row += stride) {
    const T* g = grad_out + row * dim;
    float* d = dx + row * dim;
    if (epilogue == RECEMB_EPI_L2NORM) {
      const T* y = out + row * dim;
      float dot = 0.f;
      for (int e = lane; e < dim; e += 32) dot += (float)y[e] * (float)g[e];
      dot = group_sum<32>(dot);
      const float inv = inv_norm[row];
      for (int e = lane; e < dim; e += 32) d[e] = ((float)g[e] - (float)y[e] * dot) * inv;
    } else if (epilogue == RECEMB_EPI_RSQRT_K) {
      for (int e = lane; e < dim; e += 32) d[e] = (float)g[e] / sqrt_k;
    } else {
      for (int e = lane; e < dim; e += 32) d[e] = (float)g[e];
    }
  }
}

}  // namespace recemb

using namespace recemb;

extern "C" size_t recemb_bwd_plan_bytes(int64_t n_slots, int64_t num_rows) {
  if (n_slots <= 0 || num_rows <= 0) return kCounterBytes;
  PlanLayout L;
  cudaError_t e = plan_layout(n_slots, num_rows, &L);
  if (e != cudaSuccess) {
    set_error("plan_layout: %s", cudaGetErrorString(e));
    return 0;
  }
  return L.total;
}

extern "C" int recemb_bwd_plan(const int64_t* ids, int64_t n_ids, int32_t slots_per_id,
                               int hash_mode, int64_t num_rows, int64_t hash_arg, int zero_pad,
                               int64_t pad_id, int64_t pad_row, int32_t bag_size,
                               const int32_t* lengths, int32_t last_n, void* plan,
                               size_t plan_bytes, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n_ids >= 0 && slots_per_id >= 1, "bad n_ids / slots_per_id");
  RECEMB_CHECK_ARG(plan != nullptr && plan_bytes >= kCounterBytes, "plan buffer missing");
  RECEMB_CHECK_ARG((uintptr_t)plan % 256 == 0, "plan buffer must be 256-byte aligned");
  RECEMB_CHECK_ARG(num_rows >= 1, "num_rows < 1");
  RECEMB_UNSUPPORTED(num_rows < 0xfffffff0ll, "num_rows %lld does not fit 32-bit sort keys",
                     (long long)num_rows);
  const int64_t n = n_ids * slots_per_id;
  RECEMB_UNSUPPORTED(n < 0x7fffffffll, "%lld slots do not fit 32-bit slot ids", (long long)n);
  RECEMB_CHECK_ARG(slots_per_id == 1 || hash_mode == RECEMB_HASH_ROTL_FLOORMOD,
                   "slots_per_id > 1 requires ROTL_FLOORMOD");
  RECEMB_CHECK_ARG(bag_size == 0 || slots_per_id == 1, "bags and k-shift cannot be combined");
  RECEMB_CHECK_ARG(bag_size == 0 || n_ids % bag_size == 0, "n_ids not a multiple of bag_size");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  cudaStream_t s = (cudaStream_t)stream;
  char* base = (char*)plan;
  RECEMB_CUDA(cudaMemsetAsync(base, 0, kCounterBytes, s));
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(ids != nullptr, "ids null");
  PlanLayout L;
  RECEMB_CUDA(plan_layout(n, num_rows, &L));
  if (plan_bytes < L.total) {
    set_error("plan buffer %zu < required %zu", plan_bytes, L.total);
    return RECEMB_ERR_WORKSPACE;
  }
  PlanKeyArgs a;
  a.ids = ids;
  a.n_slots = n;
  a.slots_per_id = slots_per_id;
  int rc = make_hash_spec(hash_mode, num_rows, slots_per_id > 1 ? 0 : hash_arg, &a.h);
  if (rc) return rc;
  a.zero_pad = zero_pad;
  a.pad_id = pad_id;
  a.pad_row = pad_row;
  a.bag_size = bag_size;
  a.lengths = lengths;
  a.last_n = last_n;
  a.sentinel = (uint32_t)num_rows;
  a.keys = (uint32_t*)(base + L.off_keys_in);
  a.vals = (uint32_t*)(base + L.off_vals_in);
  const int sms = sm_count(device);
  int64_t grid = (n + kBwdThreads - 1) / kBwdThreads;
  if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
  plan_keys_kernel<<<(unsigned)grid, kBwdThreads, 0, s>>>(a);
  RECEMB_LAUNCHED();
  size_t temp = L.temp_bytes;
  RECEMB_CUDA(cub::DeviceRadixSort::SortPairs(
      base + L.off_temp, temp, (const uint32_t*)a.keys, (uint32_t*)(base + L.off_keys_out),
      (const uint32_t*)a.vals, (uint32_t*)(base + L.off_vals_out), (int64_t)n, 0, L.key_bits, s));
  g_launch_count.fetch_add(1, std::memory_order_relaxed);  // the sort is >= 1 launch; counted once
  plan_count_kernel<<<(unsigned)grid, kBwdThreads, 0, s>>>(
      (const uint32_t*)(base + L.off_keys_out), n, a.sentinel, (unsigned long long*)base);
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

// n_slots / num_rows are recomputed from plan_bytes? No: the caller tells us.
extern "C" int recemb_plan_views(const void* plan, size_t plan_bytes, const uint32_t** sorted_rows,
                                 const uint32_t** sorted_slots, const int64_t** counters,
                                 int64_t* n_slots_host) {
  // the layout depends on (n_slots, num_rows); n_slots_host carries n_slots in, and
  // on return holds it unchanged.  num_rows only affects the CUB temp size which
  // sits after the arrays, so any num_rows gives the same array offsets.
  RECEMB_CHECK_ARG(plan && n_slots_host, "null pointer");
  const int64_t n = *n_slots_host;
  const size_t arr = align_up((size_t)(n > 0 ? n : 0) * 4, 256);
  RECEMB_CHECK_ARG(plan_bytes >= kCounterBytes + 4 * arr, "plan buffer too small for n_slots");
  const char* base = (const char*)plan;
  if (counters) *counters = (const int64_t*)base;
  if (sorted_rows) *sorted_rows = (const uint32_t*)(base + kCounterBytes + 2 * arr);
  if (sorted_slots) *sorted_slots = (const uint32_t*)(base + kCounterBytes + 3 * arr);
  return RECEMB_OK;
}

extern "C" size_t recemb_bwd_apply_workspace_bytes(int64_t n_slots, int32_t dim) {
  if (n_slots <= 0 || dim <= 0) return 256;
  int64_t sizes[kMaxLevels];
  const int L = level_sizes(n_slots, sizes, kMaxLevels);
  size_t total = 256;
  for (int l = 1; l < L; ++l) {
    total += align_up((size_t)sizes[l] * 4, 256);
    total += align_up((size_t)sizes[l] * dim * 4, 256);
  }
  // the last level also writes (never-read) records
  const int64_t last = 2 * ((sizes[L - 1] + kChunk - 1) / kChunk);
  total += align_up((size_t)last * 4, 256) + align_up((size_t)last * dim * 4, 256);
  return total;
}

extern "C" int recemb_bwd_apply(const void* plan, size_t plan_bytes, const void* grad,
                                int grad_dtype, int64_t grad_rows, int32_t dim,
                                int32_t slots_per_grad_row, const float* slot_weight,
                                const float* grad_row_scale, int update, void* table, int dtype,
                                int64_t num_rows, void* state1, void* state2,
                                const recemb_optim_params* hp_host, void* workspace,
                                size_t workspace_bytes, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(plan && table, "null plan/table");
  RECEMB_CHECK_ARG(grad_rows >= 0 && slots_per_grad_row >= 1, "bad grad_rows / slots_per_grad_row");
  RECEMB_CHECK_ARG(update >= RECEMB_UPD_DENSE_GRAD && update <= RECEMB_UPD_ADAMW, "bad update %d", update);
  RECEMB_CHECK_ARG(update == RECEMB_UPD_DENSE_GRAD || hp_host != nullptr, "optimizer params missing");
  RECEMB_CHECK_ARG((dtype == RECEMB_F32 || dtype == RECEMB_BF16) &&
                       (grad_dtype == RECEMB_F32 || grad_dtype == RECEMB_BF16),
                   "bad dtype");
  RECEMB_UNSUPPORTED(!(grad_dtype == RECEMB_BF16 && dtype == RECEMB_F32),
                     "bf16 gradients into an fp32 table are not supported");
  RECEMB_UNSUPPORTED(dim > 0 && dim % 4 == 0, "dim %d is not a multiple of 4", dim);
  if ((update == RECEMB_UPD_ADAGRAD || update == RECEMB_UPD_ROWWISE_ADAGRAD) && !state1) {
    set_error("adagrad needs state1");
    return RECEMB_ERR_INVALID;
  }
  if ((update == RECEMB_UPD_ADAM || update == RECEMB_UPD_ADAMW) && (!state1 || !state2)) {
    set_error("adam needs state1 and state2");
    return RECEMB_ERR_INVALID;
  }
  const int64_t n = grad_rows * slots_per_grad_row;
  RECEMB_UNSUPPORTED(n < 0x7fffffffll, "too many slots");
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(grad != nullptr && workspace != nullptr, "null grad/workspace");
  RECEMB_CHECK_ARG(((uintptr_t)grad | (uintptr_t)table | (uintptr_t)workspace) % 16 == 0,
                   "grad/table/workspace must be 16-byte aligned");
  QuadShape shape;
  RECEMB_UNSUPPORTED(pick_quads(dim / 4, &shape), "dim %d too large", dim);
  const size_t need = recemb_bwd_apply_workspace_bytes(n, dim);
  if (workspace_bytes < need) {
    set_error("workspace %zu < required %zu", workspace_bytes, need);
    return RECEMB_ERR_WORKSPACE;
  }
  const size_t arr = align_up((size_t)n * 4, 256);
  RECEMB_CHECK_ARG(plan_bytes >= kCounterBytes + 4 * arr, "plan buffer too small for these slots");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  cudaStream_t s = (cudaStream_t)stream;

  int64_t sizes[kMaxLevels];
  const int L = level_sizes(n, sizes, kMaxLevels);
  const char* pbase = (const char*)plan;
  char* w = (char*)workspace;
  size_t woff = 256;

  SegArgs a;
  a.dim = dim;
  a.quads = dim / 4;
  a.slots_per_grad_row = slots_per_grad_row;
  a.slot_weight = slot_weight;
  a.grad_row_scale = grad_row_scale;
  a.sentinel = (uint32_t)num_rows;
  a.update = update;
  a.table = table;
  a.state1 = (float*)state1;
  a.state2 = (float*)state2;
  if (hp_host) a.hp = *hp_host;
  else a.hp = recemb_optim_params{};

  const uint32_t* in_keys = (const uint32_t*)(pbase + kCounterBytes + 2 * arr);
  const uint32_t* in_slots = (const uint32_t*)(pbase + kCounterBytes + 3 * arr);
  const void* in_grad = grad;
  for (int l = 0; l < L; ++l) {
    const int64_t recs = 2 * ((sizes[l] + kChunk - 1) / kChunk);
    uint32_t* out_keys = (uint32_t*)(w + woff);
    woff += align_up((size_t)recs * 4, 256);
    float* out_part = (float*)(w + woff);
    woff += align_up((size_t)recs * dim * 4, 256);
    a.keys = in_keys;
    a.slots = in_slots;
    a.n = sizes[l];
    a.grad = in_grad;
    a.out_keys = out_keys;
    a.out_partials = out_part;
    int rc;
    if (l == 0 && t_ev_start) cudaEventRecord(t_ev_start, s);
    if (l == 0) {
      if (grad_dtype == RECEMB_F32 && dtype == RECEMB_F32)
        rc = launch_seg<float, float, true>(a, shape, s);
      else if (grad_dtype == RECEMB_BF16)
        rc = launch_seg<__nv_bfloat16, __nv_bfloat16, true>(a, shape, s);
      else
        rc = launch_seg<float, __nv_bfloat16, true>(a, shape, s);
    } else {
      if (dtype == RECEMB_F32) rc = launch_seg<float, float, false>(a, shape, s);
      else rc = launch_seg<float, __nv_bfloat16, false>(a, shape, s);
    }
    if (l == 0 && t_ev_stop) cudaEventRecord(t_ev_stop, s);
    if (l == 0) t_ev_start = t_ev_stop = nullptr;
    if (rc) return rc;
    in_keys = out_keys;
    in_slots = nullptr;
    in_grad = out_part;
  }
  return RECEMB_OK;
}

extern "C" int recemb_time_next_apply(void* start_event, void* stop_event) {
  t_ev_start = (cudaEvent_t)start_event;
  t_ev_stop = (cudaEvent_t)stop_event;
  return RECEMB_OK;
}

extern "C" int recemb_epilogue_bwd(const void* grad_out, const void* out, int dtype,
                                   const float* inv_norm, int64_t n, int32_t dim, int epilogue,
                                   int32_t num_shifts, float* dx, int device,
                                   recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n >= 0 && dim > 0, "bad shape");
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(grad_out && dx, "null pointer");
  RECEMB_CHECK_ARG(epilogue != RECEMB_EPI_L2NORM || (out && inv_norm), "L2NORM needs out and inv_norm");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  const int sms = sm_count(device);
  int64_t grid = (n * 32 + kBwdThreads - 1) / kBwdThreads;
  if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
  const float sqrt_k = (float)sqrt((double)(num_shifts > 0 ? num_shifts : 1));
  if (dtype == RECEMB_F32)
    epilogue_bwd_kernel<float><<<(unsigned)grid, kBwdThreads, 0, (cudaStream_t)stream>>>(
        (const float*)grad_out, (const float*)out, inv_norm, n, dim, epilogue, sqrt_k, dx);
  else if (dtype == RECEMB_BF16)
    epilogue_bwd_kernel<__nv_bfloat16><<<(unsigned)grid, kBwdThreads, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)grad_out, (const __nv_bfloat16*)out, inv_norm, n, dim, epilogue,
        sqrt_k, dx);
  else {
    set_error("bad dtype");
    return RECEMB_ERR_INVALID;
  }
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}
