// Backward of the embedding hot path (sm_100a):
//   plan   : lookup slots -> (row, slot) pairs sorted by row, stable in slot
//            (sort-based dedup with the hand-written radix sort of sort.cu;
//            replaces the sort inside ATen's embedding_dense_backward that
//            autograd runs for embedding_module_gen.py:113, :152)
//   apply  : chunked segmented reduction over the sorted pairs + fused
//            optimizer update of the touched rows only (replaces the dense
//            [N, D] gradient + torch.optim.Adagrad full-table pass,
//            embedding_module_gen.py:97, :137, :153)
//
// The segmented reduction is load-balanced by construction: every group of G
// lanes owns a chunk of kChunk0 consecutive sorted entries whatever the run
// lengths are (the k-shift collapse puts ~50 % of a shift's lookups on a
// handful of rows).  Runs closed inside a chunk are applied
// directly; runs that cross a chunk boundary leave (row, partial sum) records
// that the next level reduces with the same kernel, until one chunk is left.
// Summation order is fixed (sorted order, then chunk order): deterministic, no
// atomics.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "sort.cuh"

namespace recemb {

constexpr int kBwdThreads = 256;
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
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const size_t temp = sort_shape(n, L->key_bits, dev).temp_bytes;
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

// Slot indices fit 32 bits (the plan stores them as uint32), so the per-slot divisions (slot -> id, id -> bag)
// are 32-bit; every thread handles kPlanUnroll slots per round with their id loads issued together (one
// dependent 8-byte load per round was all the memory parallelism this kernel had: 2.4 TB/s).
constexpr int kPlanUnroll = 4;

__global__ void __launch_bounds__(kBwdThreads) plan_keys_kernel(const PlanKeyArgs a) {
  const uint32_t keep = window_keep(a.h);
  const uint32_t n = (uint32_t)a.n_slots;
  const uint32_t spi = (uint32_t)a.slots_per_id, bsz = (uint32_t)a.bag_size;
  const bool remap = (a.h.flip_len | a.h.win_len) != 0;
  const uint64_t round = (uint64_t)gridDim.x * kBwdThreads * kPlanUnroll;
  for (uint64_t s0 = (uint64_t)blockIdx.x * kBwdThreads * kPlanUnroll + threadIdx.x; s0 < n; s0 += round) {
    uint32_t idx[kPlanUnroll];
    int64_t idv[kPlanUnroll];
#pragma unroll
    for (int u = 0; u < kPlanUnroll; ++u) {
      const uint64_t s = s0 + (uint64_t)u * kBwdThreads;
      idx[u] = 0;
      idv[u] = 0;
      if (s < n) {
        idx[u] = spi > 1 ? (uint32_t)s / spi : (uint32_t)s;
        idv[u] = a.ids[idx[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < kPlanUnroll; ++u) {
      const uint64_t s64 = s0 + (uint64_t)u * kBwdThreads;
      if (s64 >= n) continue;
      const uint32_t s = (uint32_t)s64;
      const uint32_t id_idx = idx[u];
      const int c = spi > 1 ? (int)(s - id_idx * spi) : 0;
      const int64_t id = idv[u];
      bool ok = !(a.zero_pad && id == a.pad_id);
      uint32_t bag = 0, p = 0;
      if (bsz > 0) {
        bag = id_idx / bsz;
        p = id_idx - bag * bsz;
        if (ok) {
          int hi = (int)bsz;
          if (a.lengths) hi = min(max(a.lengths[bag], 0), (int)bsz);
          const int lo = a.last_n > 0 ? max(0, hi - a.last_n) : 0;
          ok = (int)p >= lo && (int)p < hi;
        }
      }
      // the slot's gradient row: mirrored inside its sequence when the forward wrote flipped outputs, compacted
      // to the sequence window (slots outside it carry nothing)
      const int64_t orow = remap ? out_row((int64_t)id_idx, a.h, keep) : (int64_t)id_idx;
      ok = ok && orow >= 0;
      uint32_t key = a.sentinel;
      if (ok) {
        int64_t row = spi > 1 ? kshift_row(id, c, a.h.mod_rows) : row_of(id, a.h);
        if (row != a.pad_row) {
          row = shard_local_row(row, a.h);  // -1: another rank owns this row
          if (row >= 0) key = (uint32_t)(row + table_offset((int64_t)id_idx, a.h));
        }
      }
      a.keys[s] = key;
      if (a.h.out_feats && bsz > 0) {
        // pooled bags whose gradient arrives feature-interleaved ([bags_per_table, F, dim], the interaction's
        // backward): slot -> (gradient row of its bag) * bag_size + position
        a.vals[s] = (uint32_t)(bag_out_row((int64_t)bag, a.h) * bsz + p);
      } else {
        a.vals[s] = remap ? (uint32_t)(max(orow, (int64_t)0) * spi + c) : s;
      }
    }
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
// A row is covered by G lanes x V vectors per lane; a lane-vector is E consecutive elements:
// E = 4 (16 B of fp32, 8 B of bf16) in general, E = 8 when gradients and table are both bf16
// (one 16-byte load per lane: half the instructions per row of the E = 4 mapping).
template <typename T>
__device__ __forceinline__ void unpack16(const uint4& v, float* f) { Vec16<T>::unpack(v, f); }
__device__ __forceinline__ void unpack_bf16x4(const uint2& v, float* f) {
  f[0] = __uint_as_float(v.x << 16);
  f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16);
  f[3] = __uint_as_float(v.y & 0xffff0000u);
}

// STREAM: data read exactly once (gradients, partial records) -> ld.global.nc.L1::no_allocate
template <typename T, int E, bool STREAM>
__device__ __forceinline__ void load_vec(const T* p, float* f) {
  constexpr int BYTES = E * (int)sizeof(T);
  static_assert(BYTES == 8 || BYTES == 16 || BYTES == 32, "lane vector must be 8, 16 or 32 bytes");
  if constexpr (BYTES == 8) {  // 4 bf16
    uint2 v;
    if constexpr (STREAM)
      asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    else
      v = *reinterpret_cast<const uint2*>(p);
    unpack_bf16x4(v, f);
  } else if constexpr (BYTES == 16) {  // 4 fp32 or 8 bf16
    const uint4 v = STREAM ? ldg_nc_v4(p) : ldg_v4(p);
    unpack16<T>(v, f);
  } else {  // 8 fp32
    const uint4 v0 = STREAM ? ldg_nc_v4(p) : ldg_v4(p);
    const uint4 v1 = STREAM ? ldg_nc_v4(p + 4) : ldg_v4(p + 4);
    unpack16<T>(v0, f);
    unpack16<T>(v1, f + 4);
  }
}
template <typename T, int E>
__device__ __forceinline__ void store_vec(T* p, const float* f) {
  constexpr int BYTES = E * (int)sizeof(T);
  if constexpr (BYTES == 8) {
    __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(f[2], f[3]);
    uint2 v;
    v.x = *reinterpret_cast<uint32_t*>(&a);
    v.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = v;
  } else if constexpr (BYTES == 16) {
    stg_v4(p, Vec16<T>::pack(f));
  } else {
    stg_v4(p, Vec16<T>::pack(f));
    stg_v4(p + 4, Vec16<T>::pack(f + 4));
  }
}
template <typename T, int E>
__device__ __forceinline__ void prefetch_vec(const T* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  if constexpr (E * (int)sizeof(T) == 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 4));
}

struct SegArgs {
  const uint32_t* keys;   // sorted rows (level 0) or record rows (level >= 1)
  const uint32_t* slots;  // level 0 only
  int32_t n;              // entries at this level (< 2^31)
  const void* grad;       // level 0: [grad_rows, dim] GT ; level >= 1: fp32 partials [n, dim]
  uint32_t dim;
  int32_t vecs;           // dim / E lane-vectors per row
  uint32_t spg;           // slots per gradient row (k shifts / bag size), >= 1
  uint32_t spg_magic;     // floor(2^32 / spg): slot / spg without an integer divide
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
  const uint32_t* flag_in;  // level >= 1: non-zero iff the previous level emitted a record
  uint32_t* flag_out;
  const uint32_t* guard;  // optional: a non-zero word means "do not touch the table" (every CTA returns)
  int32_t chunk;          // consecutive entries per group (level 0: wave-fitted, >= kChunk0)
  int32_t pf_bulk;        // level 0 fast path: rows prefetched with one bulk L2 prefetch per row
  PeerGate gate;          // seg_pre_kernel, sharded backward: fused gradient push + per-table gating
  int32_t* chunk_used;    // HOST pointer (never read on the device): the level-0 chunk length the launch chose
};

// ---- fused gradient push (sharded backward) ---------------------------------------------------
// Pusher CTAs (blockIdx.x < gate.push_ctas; dispatched first, they never wait): my pooled gradients go to
// slice `rank` of every rank's gradient buffer, table by table -- one 16-byte load, `world` stores per
// vector, peers visited in rotated order so the ranks never all store into the same peer.  When the last
// pusher CTA is done with table t it adds one to the arrival count of table t on every rank (system
// scope): the peers' reduction groups that need table t are waiting for exactly that.
__device__ __forceinline__ void gate_push_role(const PeerGate& g) {
  __shared__ uint4* s_dst[RECEMB_MAX_PEERS];
  char* mine = g.arena[g.rank];
  if ((int)threadIdx.x < g.world)
    s_dst[threadIdx.x] = (uint4*)(g.arena[threadIdx.x] + g.off_grads) + (int64_t)g.rank * g.sender_vecs;
  __syncthreads();
  uint32_t* done = (uint32_t*)(mine + g.off_gate + kGateOffDone);
  // one table, copied by CTAs `sub` of `cnt`, then signalled
  auto push_table = [&](int t, int sub, int cnt) {
    const uint4* src = g.src + (int64_t)t * g.vecs_per_table;
    const int64_t dst0 = (int64_t)t * g.vecs_per_table;
    // table-wise partitioning: table t has ONE owner, t % world -- its gradients cross NVLink once
    const int q_lo = g.partition ? ((t % g.world) - g.rank + g.world - 1) % g.world + 1 : 1;
    const int q_hi = g.partition ? q_lo : g.world;
    const int64_t stride = (int64_t)cnt * kBwdThreads * 4;
    for (int64_t i0 = (int64_t)sub * kBwdThreads * 4 + threadIdx.x; i0 < g.vecs_per_table && !(g.debug & 2);
         i0 += stride) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + u * kBwdThreads;
        if (i < g.vecs_per_table) v[u] = ldg_nc_v4(src + i);
      }
      for (int q = q_lo; q <= q_hi; ++q) {
        int p = g.rank + q;
        if (p >= g.world) p -= g.world;
        uint4* dst = s_dst[p] + dst0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t i = i0 + u * kBwdThreads;
          if (i < g.vecs_per_table) stg_v4(dst + i, v[u]);
        }
      }
    }
    // ONE system-scope fence per CTA and table, by one thread after the CTA barrier (the barrier orders the
    // other threads' stores before it).  Only warp 0 signals; the other warps go straight on to the next table.
    __syncthreads();
    if (threadIdx.x < 32) {
      int last = 0;
      if (threadIdx.x == 0) {
        asm volatile("fence.acq_rel.sys;" ::: "memory");
        // wraps back to 0 with the last pusher CTA of the table: nothing to reset
        last = atomicInc(done + t, (unsigned)cnt - 1u) == (unsigned)cnt - 1u;
      }
      last = __shfl_sync(0xffffffffu, last, 0);
      // one more sender's table t has landed: its arrival count goes up by one on the ranks that wait for it
      // (remote atomic over NVLink); the count only ever grows, step s is complete at s * world.  Every pusher
      // CTA fenced before its increment of `done`, so all of the table is visible system-wide here.
      if (g.partition) {  // only the owner waits for table t, as its local table t / world
        if (last && (int)threadIdx.x == t % g.world)
          atomicAdd_system((unsigned long long*)(g.arena[threadIdx.x] + g.off_gate + kGateOffFlags) + t / g.world, 1ull);
      } else if (last && (int)threadIdx.x < g.world) {
        atomicAdd_system((unsigned long long*)(g.arena[threadIdx.x] + g.off_gate + kGateOffFlags) + t, 1ull);
      }
    }
  };
  if (g.partition && g.push_ctas % g.tables == 0) {
    // table-wise: every owner waits for exactly its own tables from everybody -- push the tables side by side
    // (CTA c copies table c % T) so that no owner sits behind the other owners' tables
    push_table((int)blockIdx.x % g.tables, (int)blockIdx.x / g.tables, g.push_ctas / g.tables);
  } else {
    // row-wise: the reduction consumes the tables in order -- push them one after the other, all CTAs on each
    for (int t = 0; t < g.tables; ++t) push_table(t, (int)blockIdx.x, g.push_ctas);
  }
}

// Reduction side: ONE thread per CTA waits until every rank's gradients of table `t_need` (the last table the
// CTA's chunks touch) have landed here, i.e. until the table's arrival count reaches step * world; the CTA
// barrier releases the other groups.  One poller per CTA, relaxed loads, back-off from 0.5 to 8 us: thousands
// of groups wait at the start of a step, and a spin from all of them on one word saturates its L2 bank -- the
// pushers' own atomics then queue behind the polls (measured: +0.13 ms on a 0.25 ms update).  False on a
// timeout: status bit 2 is set and the CTA's chunks are skipped -- a late peer may lose a step, never corrupt it.
__device__ __forceinline__ bool gate_wait(const PeerGate& g, int t_need) {
  char* mine = g.arena[g.rank];
  const uint64_t want = *(const volatile uint64_t*)(mine + g.off_gate) * (uint64_t)g.world;
  const uint64_t* count = (const uint64_t*)(mine + g.off_gate + kGateOffFlags) + t_need;
  const long long t0 = clock64();
  unsigned ns = 500;
  bool ok = true;
  while (ld_relaxed_sys_u64(count) < want) {
    if (clock64() - t0 > g.timeout_cycles) {
      atomicOr(g.status, 2u);
      ok = false;
      break;
    }
    __nanosleep(ns);
    if (ns < 8000) ns *= 2;
  }
  asm volatile("fence.acq_rel.sys;" ::: "memory");
  return ok;
}

__global__ void gate_advance_kernel(uint64_t* step) { *step += 1; }


template <int G>
__device__ __forceinline__ float masked_group_sum(float v, uint32_t mask) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

// The optimizer math is a few flops per 16 bytes moved, but at HBM speed the SM has only
// ~96 issue slots per 512-byte warp access: IEEE sqrt/div sequences (~20 instructions per
// element) would make the update instruction-bound.  One MUFU each for sqrt and reciprocal
// (~2 ulp) is far inside the 1e-5 parity tolerance on updated weights.
__device__ __forceinline__ float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// whole row (16-byte aligned, multiple of 16 bytes) into L2 through the TMA engine: one
// instruction per row instead of one per lane, and it translates like a real access (a
// per-lane prefetch.global.L2 that misses the TLB may be dropped: tables of tens of GB)
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// UPD >= 0: the update kind is a compile-time constant (dead variants disappear);
// UPD < 0: read a.update at run time.  EXACT: the row is exactly G*V quads.
template <int G, int V, int E, typename WT, int UPD, bool EXACT>
__device__ __forceinline__ void apply_row(const SegArgs& a, uint32_t row, float (&g)[V][E], int lig,
                                          uint32_t gmask) {
  const int upd = UPD >= 0 ? UPD : a.update;
  const size_t off = (size_t)row * a.dim + lig * E;
  WT* wrow = reinterpret_cast<WT*>(a.table) + off;
  const recemb_optim_params& hp = a.hp;
  auto live = [&](int j) { return EXACT || (j * G + lig) < a.vecs; };
  if (upd == RECEMB_UPD_DENSE_GRAD) {
#pragma unroll
    for (int j = 0; j < V; ++j)
      if (live(j)) store_vec<WT, E>(wrow + j * G * E, g[j]);
    return;
  }
  float w[V][E];
#pragma unroll
  for (int j = 0; j < V; ++j) {
#pragma unroll
    for (int e = 0; e < E; ++e) w[j][e] = 0.f;
    if (live(j)) load_vec<WT, E, false>(wrow + j * G * E, w[j]);
  }
  if (hp.weight_decay != 0.f && upd != RECEMB_UPD_ADAMW) {
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int e = 0; e < E; ++e) g[j][e] += hp.weight_decay * w[j][e];
  }
  if (upd == RECEMB_UPD_SGD) {
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int e = 0; e < E; ++e) w[j][e] -= hp.lr * g[j][e];
  } else if (upd == RECEMB_UPD_ADAGRAD) {
    float* srow = a.state1 + off;
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (live(j)) {
        float s[E];
        load_vec<float, E, false>(srow + j * G * E, s);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          s[e] += g[j][e] * g[j][e];
          w[j][e] += (-hp.lr * g[j][e]) * fast_rcp(fast_sqrt(s[e]) + hp.eps);
        }
        store_vec<float, E>(srow + j * G * E, s);
      }
    }
  } else if (upd == RECEMB_UPD_ROWWISE_ADAGRAD) {
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int e = 0; e < E; ++e) ss += g[j][e] * g[j][e];  // lanes past the row hold zeros
    ss = masked_group_sum<G>(ss, gmask) / (float)a.dim;
    const float s_new = a.state1[row] + ss;
    const float inv = -hp.lr * fast_rcp(fast_sqrt(s_new) + hp.eps);
#pragma unroll
    for (int j = 0; j < V; ++j)
#pragma unroll
      for (int e = 0; e < E; ++e) w[j][e] += g[j][e] * inv;
    __syncwarp(gmask);  // every lane has read state1[row] before lane 0 overwrites it
    if (lig == 0) a.state1[row] = s_new;
  } else {  // ADAM / ADAMW, lazy: only touched rows move
    float* mrow = a.state1 + off;
    float* vrow = a.state2 + off;
    const float step_size = hp.lr / hp.bias_correction1;
    const float inv_bc2_sqrt = 1.f / sqrtf(hp.bias_correction2);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      if (live(j)) {
        float m[E], v[E];
        load_vec<float, E, false>(mrow + j * G * E, m);
        load_vec<float, E, false>(vrow + j * G * E, v);
#pragma unroll
        for (int e = 0; e < E; ++e) {
          if (upd == RECEMB_UPD_ADAMW) w[j][e] *= (1.f - hp.lr * hp.weight_decay);
          m[e] = hp.beta1 * m[e] + (1.f - hp.beta1) * g[j][e];
          v[e] = hp.beta2 * v[e] + (1.f - hp.beta2) * g[j][e] * g[j][e];
          const float denom = fast_sqrt(v[e]) * inv_bc2_sqrt + hp.eps;
          w[j][e] -= step_size * m[e] * fast_rcp(denom);
        }
        store_vec<float, E>(mrow + j * G * E, m);
        store_vec<float, E>(vrow + j * G * E, v);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < V; ++j)
    if (live(j)) store_vec<WT, E>(wrow + j * G * E, w[j]);
}

// Compile-time shape of one instantiation of the walk.
//   B     entries per batch            PD   prefetch distance in batches
//   UPD   update kind or -1 (runtime)  PLAIN  level 0, one slot per gradient row, no scales
//   EXACT row is exactly G*V quads     MINB   min CTAs / SM (register cap)
template <int B_, int PD_, int MINB_, int UPD_, bool PLAIN_, bool EXACT_>
struct SegCfg {
  static constexpr int B = B_, PD = PD_, MINB = MINB_, UPD = UPD_;
  static constexpr bool PLAIN = PLAIN_, EXACT = EXACT_;
};

// One group of G lanes walks CH consecutive sorted entries B at a time.  While batch b is
// reduced, the rows batch b+PD will touch (its gradient rows and, for every run that ends
// inside it, the table / optimizer-state rows) are pulled into L2 with prefetch.global.L2, so
// the dependent loads of the walk hit L2 instead of paying a DRAM (+TLB) round trip each.
//
// Records: chunk c leaves at most two (row, partial) records for the next level --
// leading(c) at index 2c-1 (its first run continues from chunk c-1) and trailing(c) at 2c
// (its last run continues into chunk c+1).  The two halves of a run cut by ONE boundary are
// therefore the aligned pair (2c, 2c+1): the next level closes them inside one chunk, and
// only runs longer than a chunk reach the levels above (which exit on an empty-level flag).
template <int G, int V, int E, typename GT, typename WT, bool L0, int CH, typename Cfg>
__global__ void __launch_bounds__(kBwdThreads, Cfg::MINB) seg_kernel(const SegArgs a) {
  if (!L0 && *a.flag_in == 0) return;  // the previous level emitted nothing
  if (a.guard && *reinterpret_cast<const volatile uint32_t*>(a.guard) != 0u) return;
  constexpr int B = Cfg::B, PD = Cfg::PD;
  constexpr bool PLAIN = Cfg::PLAIN && L0, EXACT = Cfg::EXACT;
  constexpr int GROUPS = kBwdThreads / G;
  const int lane = threadIdx.x & 31;
  const int lig = lane % G;
  const int gi_warp = lane / G;
  const uint32_t gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (gi_warp * G));
  const int chunk = blockIdx.x * GROUPS + threadIdx.x / G;
  const int CHR = L0 ? a.chunk : CH;  // level 0: run-time chunk (wave-fitted by the host)
  const int num_chunks = (a.n + CHR - 1) / CHR;
  if (chunk >= num_chunks) return;
  const int start = chunk * CHR;
  const int end = min(start + CHR, a.n);

  if (lig == 0) {
    if (chunk > 0) a.out_keys[2 * chunk - 1] = kNoKey;
    a.out_keys[2 * chunk] = kNoKey;
    if (chunk == num_chunks - 1) a.out_keys[2 * chunk + 1] = kNoKey;
  }

  // dropped slots (pad ids, positions outside a bag's window, unused inbox capacity) carry the
  // sentinel key and sort last: a chunk that starts with one holds nothing else
  if (L0 && a.keys[start] >= a.sentinel) return;
  const float grad_mul = a.hp.grad_div > 0.f ? 1.f / a.hp.grad_div : 1.f;
  const int upd = Cfg::UPD >= 0 ? Cfg::UPD : a.update;
  const GT* gbase = reinterpret_cast<const GT*>(a.grad) + lig * E;
  const WT* tbase = reinterpret_cast<const WT*>(a.table) + lig * E;
  const bool pf_w = upd != RECEMB_UPD_DENSE_GRAD;
  const bool pf_s1 = upd == RECEMB_UPD_ADAGRAD || upd >= RECEMB_UPD_ADAM;
  const bool pf_s2 = upd >= RECEMB_UPD_ADAM;
  auto live = [&](int j) { return EXACT || (j * G + lig) < a.vecs; };

  auto grad_row_of = [&](int i, uint32_t slot) -> uint32_t {
    if (!L0) return (uint32_t)i;
    if (PLAIN || a.spg == 1) return slot;
    uint32_t q = __umulhi(slot, a.spg_magic);  // q in {floor(slot/spg) - 1, floor(slot/spg)}
    if (slot - q * a.spg >= a.spg) ++q;
    return q;
  };
  // keys of entries [e0, e0+B] (one look-ahead past the batch) and slots of [e0, e0+B)
  auto load_meta = [&](int e0, uint32_t(&kk)[B + 1], uint32_t(&ss)[B]) {
#pragma unroll
    for (int u = 0; u <= B; ++u) {
      const int i = e0 + u;
      const bool in = (u < B) ? (i < end) : (i < a.n);
      kk[u] = in ? a.keys[i] : kNoKey;
      if (u < B) ss[u] = (L0 && in) ? a.slots[i] : 0u;
    }
  };
  // L2 prefetch of what a batch will touch, from its already loaded keys / slots
  auto prefetch_from = [&](int e0, const uint32_t(&kk)[B + 1], const uint32_t(&ss)[B]) {
    if constexpr (L0 && EXACT) {
      if (a.pf_bulk) {
        if (lig == 0) {
#pragma unroll
          for (int u = 0; u < B; ++u) {
            if (kk[u] < a.sentinel) {
              prefetch_l2_bulk(gbase + (size_t)grad_row_of(e0 + u, ss[u]) * a.dim, a.dim * (uint32_t)sizeof(GT));
              if (pf_w && kk[u + 1] != kk[u]) {
                const size_t off = (size_t)kk[u] * a.dim;
                prefetch_l2_bulk(tbase + off, a.dim * (uint32_t)sizeof(WT));
                if (pf_s1) prefetch_l2_bulk(a.state1 + off, a.dim * 4u);
                if (pf_s2) prefetch_l2_bulk(a.state2 + off, a.dim * 4u);
                if (upd == RECEMB_UPD_ROWWISE_ADAGRAD) prefetch_l2(a.state1 + kk[u]);
              }
            }
          }
        }
        return;
      }
    }
#pragma unroll
    for (int u = 0; u < B; ++u) {
      if (kk[u] < a.sentinel) {
        const GT* src = gbase + (size_t)grad_row_of(e0 + u, ss[u]) * a.dim;
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (live(j)) prefetch_vec<GT, E>(src + j * G * E);
        if (pf_w && kk[u + 1] != kk[u]) {  // this run ends here: its row will be updated
          const size_t off = (size_t)kk[u] * a.dim;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            if (live(j)) {
              prefetch_vec<WT, E>(tbase + off + j * G * E);
              if (pf_s1) prefetch_vec<float, E>(a.state1 + off + lig * E + j * G * E);
              if (pf_s2) prefetch_vec<float, E>(a.state2 + off + lig * E + j * G * E);
            }
          }
          if (upd == RECEMB_UPD_ROWWISE_ADAGRAD && lig == 0) prefetch_l2(a.state1 + kk[u]);
        }
      }
    }
  };
  auto prefetch_batch = [&](int e0) {
    if (e0 >= end) return;
    uint32_t kk[B + 1], ss[B];
    load_meta(e0, kk, ss);
    prefetch_from(e0, kk, ss);
  };

  // PD == 1: the keys / slots of batch b+1 are loaded while batch b is processed and carried in
  // registers -- they drive the L2 prefetch and then the walk itself (no dependent key load at the
  // top of an iteration; the non-plain walks are issue-bound, not DRAM-bound)
  uint32_t k[B + 1], sl[B];
  load_meta(start, k, sl);
  if (PD == 1) {
    prefetch_from(start, k, sl);
  } else {
#pragma unroll
    for (int d = 0; d < PD; ++d) prefetch_batch(start + d * B);
  }

  uint32_t cur = k[0];
  const bool left_open = start > 0 && a.keys[start - 1] == cur;
  bool first = true;
  float acc[V][E];
#pragma unroll
  for (int j = 0; j < V; ++j)
#pragma unroll
    for (int e = 0; e < E; ++e) acc[j][e] = 0.f;

  auto flush = [&](uint32_t key, bool leading, bool trailing) {
    if (key >= a.sentinel) return;
    if (!leading && !trailing) {
      apply_row<G, V, E, WT, Cfg::UPD, EXACT>(a, key, acc, lig, gmask);
      return;
    }
    const int rec = leading ? 2 * chunk - 1 : 2 * chunk;
    float* dst = a.out_partials + (size_t)rec * a.dim + lig * E;
#pragma unroll
    for (int j = 0; j < V; ++j)
      if (live(j)) store_vec<float, E>(dst + j * G * E, acc[j]);
    if (lig == 0) {
      a.out_keys[rec] = key;
      *a.flag_out = 1u;
    }
    if (leading && trailing) {  // the whole chunk is one run open on both sides
      float z[E] = {};
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (live(j)) store_vec<float, E>(dst + a.dim + j * G * E, z);
      if (lig == 0) a.out_keys[rec + 1] = key;
    }
  };

  for (int e0 = start; e0 < end; e0 += B) {
    uint32_t nk[B + 1] = {}, ns[B] = {};
    if (PD == 1) {
      if (e0 + B < end) {
        load_meta(e0 + B, nk, ns);
        prefetch_from(e0 + B, nk, ns);
      }
    } else {
      prefetch_batch(e0 + PD * B);
      if (e0 > start) load_meta(e0, k, sl);
    }
    float g[B][V][E];
    float wt[B];
#pragma unroll
    for (int u = 0; u < B; ++u) {
      wt[u] = 1.f;
#pragma unroll
      for (int j = 0; j < V; ++j)
#pragma unroll
        for (int e = 0; e < E; ++e) g[u][j][e] = 0.f;
      if (k[u] < a.sentinel) {
        const uint32_t grow = grad_row_of(e0 + u, sl[u]);
        if (L0 && !PLAIN) {
          if (a.slot_weight) wt[u] = a.slot_weight[sl[u]];
          if (a.grad_row_scale) wt[u] *= a.grad_row_scale[grow];
        }
        const GT* src = gbase + (size_t)grow * a.dim;
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (live(j)) load_vec<GT, E, true>(src + j * G * E, g[u][j]);
      }
    }
#pragma unroll
    for (int u = 0; u < B; ++u) {
      if (e0 + u < end) {
        if (k[u] != cur) {
          flush(cur, first && left_open, false);
          first = false;
          cur = k[u];
#pragma unroll
          for (int j = 0; j < V; ++j)
#pragma unroll
            for (int e = 0; e < E; ++e) acc[j][e] = 0.f;
        }
        if (k[u] < a.sentinel) {
          if (L0 && !PLAIN && (a.slot_weight || a.grad_row_scale)) {
#pragma unroll
            for (int j = 0; j < V; ++j)
#pragma unroll
              for (int e = 0; e < E; ++e) acc[j][e] += wt[u] * g[u][j][e];
          } else if (L0 && !PLAIN && a.hp.grad_div > 0.f) {
            // one FMUL per element: an IEEE division costs ~8 instructions each and this walk is
            // issue-bound (74 % issue-active at 37 % DRAM); <= 1 ulp from the reference's g / sqrt(k)
#pragma unroll
            for (int j = 0; j < V; ++j)
#pragma unroll
              for (int e = 0; e < E; ++e) acc[j][e] += g[u][j][e] * grad_mul;
          } else {
#pragma unroll
            for (int j = 0; j < V; ++j)
#pragma unroll
              for (int e = 0; e < E; ++e) acc[j][e] += g[u][j][e];
          }
        }
      }
    }
    if (PD == 1) {
#pragma unroll
      for (int u = 0; u <= B; ++u) k[u] = nk[u];
#pragma unroll
      for (int u = 0; u < B; ++u) sl[u] = ns[u];
    }
  }
  const bool right_open = end < a.n && a.keys[end] == cur;
  flush(cur, first && left_open, right_open);
}

// ---- level 0, mostly-unique rows: table rows loaded together with the gradients --------------
// seg_kernel closes a run when the NEXT key arrives and only then loads the row it updates: with
// few duplicates (pooled bags over large vocabularies: every entry its own run) a group has one
// dependent row access in flight at a time and the walk is bound by memory latency, not
// bandwidth.  Here a run is closed at its LAST entry (k[u+1] != k[u], one look-ahead key), so
// the table rows (and row-wise Adagrad states) of every run that ends inside a batch are
// requested at the top of the batch next to its gradient rows: 2B row accesses in flight per
// group instead of ~1.  Rows of 16 lanes x 16 bytes (or 32 x 16), update kind fixed at compile
// time: row-wise Adagrad or SGD (the update kinds whose only per-row state is a scalar).
template <int G, int E, typename T, typename Cfg>
__global__ void __launch_bounds__(kBwdThreads, Cfg::MINB) seg_pre_kernel(const SegArgs a) {
  constexpr int B = Cfg::B, PD = Cfg::PD, UPD = Cfg::UPD;
  constexpr bool PLAIN = Cfg::PLAIN;
  constexpr int GROUPS = kBwdThreads / G;
  static_assert(E * sizeof(T) == 16, "one 16-byte vector per lane");
  static_assert(UPD == RECEMB_UPD_ROWWISE_ADAGRAD || UPD == RECEMB_UPD_SGD, "scalar-state updates only");
  if (a.gate.push_ctas > 0 && (int)blockIdx.x < a.gate.push_ctas) {  // the peers need my gradients whatever
    gate_push_role(a.gate);                                           // my own status says
    return;
  }
  if (a.guard && *reinterpret_cast<const volatile uint32_t*>(a.guard) != 0u) return;
  const int lane = threadIdx.x & 31;
  const int lig = lane % G;
  const int gi_warp = lane / G;
  const uint32_t gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << (gi_warp * G));
  const int cta = (int)blockIdx.x - a.gate.push_ctas;
  if (a.gate.push_ctas > 0) {
    // sorted keys are table-major: the CTA's chunks need the gradients of the tables up to its last valid key's
    __shared__ int s_go;
    if (threadIdx.x == 0) {
      const int64_t first = (int64_t)cta * GROUPS * a.chunk;
      const int64_t last = min(first + (int64_t)GROUPS * a.chunk, (int64_t)a.n) - 1;
      int go = 1;
      if (first < a.n && a.keys[first] < a.sentinel) {
        const uint32_t klast = a.keys[last];
        const int t_need = klast >= a.sentinel ? a.gate.local_tables - 1
                                               : min((int)(klast / a.gate.rows_per_table), a.gate.local_tables - 1);
        if (!(a.gate.debug & 1)) go = gate_wait(a.gate, t_need);
      }
      s_go = go;
    }
    __syncthreads();
    if (!s_go) return;
  }
  const int chunk = cta * GROUPS + threadIdx.x / G;
  const int num_chunks = (a.n + a.chunk - 1) / a.chunk;
  if (chunk >= num_chunks) return;
  const int start = chunk * a.chunk;
  const int end = min(start + a.chunk, a.n);
  if (lig == 0) {
    if (chunk > 0) a.out_keys[2 * chunk - 1] = kNoKey;
    a.out_keys[2 * chunk] = kNoKey;
    if (chunk == num_chunks - 1) a.out_keys[2 * chunk + 1] = kNoKey;
  }
  if (a.keys[start] >= a.sentinel) return;  // sentinel keys sort last: nothing but dropped slots here
  const float grad_mul = a.hp.grad_div > 0.f ? 1.f / a.hp.grad_div : 1.f;
  const T* gbase = reinterpret_cast<const T*>(a.grad) + lig * E;
  T* tbase = reinterpret_cast<T*>(a.table) + lig * E;
  const recemb_optim_params& hp = a.hp;

  auto grad_row_of = [&](uint32_t slot) -> uint32_t {
    if (PLAIN || a.spg == 1) return slot;
    uint32_t q = __umulhi(slot, a.spg_magic);
    if (slot - q * a.spg >= a.spg) ++q;
    return q;
  };
  // true keys of [e0, e0+B] (kNoKey past the list) and slots of [e0, e0+B)
  auto load_meta = [&](int e0, uint32_t(&kk)[B + 1], uint32_t(&ss)[B]) {
#pragma unroll
    for (int u = 0; u <= B; ++u) {
      const int i = e0 + u;
      kk[u] = i < a.n ? a.keys[i] : kNoKey;
      if (u < B) ss[u] = i < a.n ? a.slots[i] : 0u;
    }
  };
  // L2 prefetch of what a batch will touch, from its (already loaded) meta data
  auto prefetch_meta = [&](int e0, const uint32_t(&kk)[B + 1], const uint32_t(&gr)[B]) {
#pragma unroll
    for (int u = 0; u < B; ++u) {
      if (e0 + u < end && kk[u] < a.sentinel) {
        prefetch_vec<T, E>(gbase + (size_t)gr[u] * a.dim);
        if (kk[u + 1] != kk[u]) {
          prefetch_vec<T, E>(tbase + (size_t)kk[u] * a.dim);
          if (UPD == RECEMB_UPD_ROWWISE_ADAGRAD && lig == 0) prefetch_l2(a.state1 + kk[u]);
        }
      }
    }
  };
  auto load_batch = [&](int e0, uint32_t(&kk)[B + 1], uint32_t(&ss)[B], uint32_t(&gr)[B]) {
    load_meta(e0, kk, ss);
#pragma unroll
    for (int u = 0; u < B; ++u) gr[u] = grad_row_of(ss[u]);
  };
  auto prefetch_batch = [&](int e0) {
    if (e0 >= end) return;
    uint32_t kk[B + 1], ss[B], gr[B];
    load_batch(e0, kk, ss, gr);
    prefetch_meta(e0, kk, gr);
  };

  const bool left_open = start > 0 && a.keys[start - 1] == a.keys[start];
  bool first = true;
  float acc[E];
#pragma unroll
  for (int e = 0; e < E; ++e) acc[e] = 0.f;

  // The meta data (keys, slots, gradient rows) of batch b+1 is loaded while batch b is processed
  // and carried in registers (PD == 1: the same data drives the L2 prefetch): no dependent key
  // load at the top of an iteration and no second hash of the same slots -- the walk is
  // issue-bound (61-75 % issue-active), not DRAM-bound.
  uint32_t k[B + 1], sl[B], gr[B];
  load_batch(start, k, sl, gr);
  if (PD == 1) {
    prefetch_meta(start, k, gr);
  } else {
#pragma unroll
    for (int d = 0; d < PD; ++d) prefetch_batch(start + d * B);
  }

  for (int e0 = start; e0 < end; e0 += B) {
    uint32_t nk[B + 1] = {}, ns[B] = {}, ngr[B] = {};
    if (PD == 1) {
      if (e0 + B < end) {
        load_batch(e0 + B, nk, ns, ngr);
        prefetch_meta(e0 + B, nk, ngr);
      }
    } else {
      if (PD > 0) prefetch_batch(e0 + PD * B);
      if (e0 > start) load_batch(e0, k, sl, gr);
    }
    uint4 gv[B], wv[B];
    float sv[B], wt[B];
    bool closes[B];
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const bool live = e0 + u < end && k[u] < a.sentinel;
      gv[u] = make_uint4(0, 0, 0, 0);
      wv[u] = make_uint4(0, 0, 0, 0);
      sv[u] = 0.f;
      wt[u] = 1.f;
      closes[u] = live && k[u + 1] != k[u];
      if (live) {
        const uint32_t grow = gr[u];
        gv[u] = ldg_nc_v4(gbase + (size_t)grow * a.dim);
        if (!PLAIN) {
          if (a.slot_weight) wt[u] = a.slot_weight[sl[u]];
          if (a.grad_row_scale) wt[u] *= a.grad_row_scale[grow];
        }
      }
      if (closes[u]) {
        wv[u] = ldg_v4(tbase + (size_t)k[u] * a.dim);
        if (UPD == RECEMB_UPD_ROWWISE_ADAGRAD) sv[u] = a.state1[k[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < B; ++u) {
      const int i = e0 + u;
      if (i < end) {
        const bool live = k[u] < a.sentinel;
        if (live) {
          float f[E];
          unpack16<T>(gv[u], f);
          if (!PLAIN && (a.slot_weight || a.grad_row_scale)) {
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] += wt[u] * f[e];
          } else if (!PLAIN && hp.grad_div > 0.f) {
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] += f[e] * grad_mul;
          } else {
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] += f[e];
          }
        }
        const bool chunk_last = i == end - 1;
        if (live && (closes[u] || chunk_last)) {
          const bool leading = first && left_open;
          const bool trailing = !closes[u];  // the run goes on in the next chunk
          if (!leading && !trailing) {
            float w[E];
            unpack16<T>(wv[u], w);
            if (hp.weight_decay != 0.f) {
#pragma unroll
              for (int e = 0; e < E; ++e) acc[e] += hp.weight_decay * w[e];
            }
            if (UPD == RECEMB_UPD_SGD) {
#pragma unroll
              for (int e = 0; e < E; ++e) w[e] -= hp.lr * acc[e];
            } else {
              float ss = 0.f;
#pragma unroll
              for (int e = 0; e < E; ++e) ss += acc[e] * acc[e];
              ss = masked_group_sum<G>(ss, gmask) / (float)a.dim;
              const float s_new = sv[u] + ss;
              const float inv = -hp.lr * fast_rcp(fast_sqrt(s_new) + hp.eps);
#pragma unroll
              for (int e = 0; e < E; ++e) w[e] += acc[e] * inv;
              if (lig == 0) a.state1[k[u]] = s_new;
            }
            stg_v4(tbase + (size_t)k[u] * a.dim, Vec16<T>::pack(w));
          } else {
            const int rec = leading ? 2 * chunk - 1 : 2 * chunk;
            float* dst = a.out_partials + (size_t)rec * a.dim + lig * E;
            store_vec<float, E>(dst, acc);
            if (lig == 0) {
              a.out_keys[rec] = k[u];
              *a.flag_out = 1u;
            }
            if (leading && trailing) {  // the whole chunk is one run open on both sides
              float z[E] = {};
              store_vec<float, E>(dst + a.dim, z);
              if (lig == 0) a.out_keys[rec + 1] = k[u];
            }
          }
        }
        if (closes[u] || chunk_last || !live) {
          if (live || closes[u]) first = false;
#pragma unroll
          for (int e = 0; e < E; ++e) acc[e] = 0.f;
        }
      }
    }
    if (PD == 1) {
#pragma unroll
      for (int u = 0; u <= B; ++u) k[u] = nk[u];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        sl[u] = ns[u];
        gr[u] = ngr[u];
      }
    }
  }
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

#ifndef RECEMB_CHUNK0
#define RECEMB_CHUNK0 64
#endif
constexpr int kChunk0 = RECEMB_CHUNK0;  // sorted entries per group at level 0
constexpr int kChunkN = 32;             // records per group at levels >= 1

static int env_flag(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <auto kernel, int G, bool L0>
static void launch_kernel(const SegArgs& a_in, cudaStream_t s) {
  constexpr int CH = L0 ? kChunk0 : kChunkN;
  constexpr int groups = kBwdThreads / G;
  SegArgs a = a_in;
  a.chunk = CH;
  if constexpr (L0) {
    // Wave fit: every group does the same amount of work, so a grid of 2.2 waves takes as long
    // as 3.  Stretch the chunk (never below kChunk0: the workspace is sized for that) until the
    // chunks fill a whole number of waves of resident groups.
    static int occ_of[64];  // per instantiation (the kernel is a template argument) and per device
    static const int fit = env_flag("RECEMB_SEG_FIT", 1);
    int dev = 0;
    cudaGetDevice(&dev);
    int& occ = occ_of[(dev >= 0 && dev < 64) ? dev : 0];
    if (occ == 0) {
      int v = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kBwdThreads, 0) != cudaSuccess || v < 1) v = 1;
      occ = v;
    }
    const int64_t resident = (int64_t)occ * sm_count(dev) * groups;
    const int64_t chunks64 = ((int64_t)a.n + CH - 1) / CH;
    if (fit && chunks64 > resident) {
      // the pusher CTAs of a fused launch hold slots of the first wave: without them in the count the
      // reduction spills into one more (almost empty) wave, +0.13 ms on cfg 5
      const int64_t waves = chunks64 / resident;
      const int64_t slots = waves * resident - (int64_t)a.gate.push_ctas * groups;
      const int64_t c = slots > 0 ? ((int64_t)a.n + slots - 1) / slots : 0;
      if (c > CH && c <= 2 * CH) a.chunk = (int32_t)c;
    }
    if (a.chunk_used) *a.chunk_used = a.chunk;  // the host sizes level 1 from it
  }
  const int chunks = (a.n + a.chunk - 1) / a.chunk;
  kernel<<<(unsigned)((chunks + groups - 1) / groups + a.gate.push_ctas), kBwdThreads, 0, s>>>(a);
}

template <int G, int V, int E, typename GT, typename WT, bool L0, typename Cfg>
static void launch_one(const SegArgs& a, cudaStream_t s) {
  launch_kernel<seg_kernel<G, V, E, GT, WT, L0, (L0 ? kChunk0 : kChunkN), Cfg>, G, L0>(a, s);
}

// generic instantiations: any shape / update / scale combination
#define SEG_GV(G_, V_)                                                                          \
  if (shape.G == G_ && shape.V == V_) {                                                         \
    using Cfg = SegCfg<(V_ == 1) ? 2 : 1, 2, (V_ == 1) ? 4 : (V_ == 2 ? 2 : 1), -1, false, false>; \
    launch_one<G_, V_, 4, GT, WT, L0, Cfg>(a, s);                                                  \
    launched = true;                                                                            \
  }

// specialised level-0 instantiations for the rows the BASELINE configs use (256-byte rows:
// D = 64 fp32 -> G = 16, D = 128 bf16 -> G = 32), update kind fixed at compile time
template <int G, typename GT, typename WT, bool PLAIN, int B, int PD, int MINB>
static bool launch_fast(const SegArgs& a, cudaStream_t s) {
  constexpr int E = 16 / (int)sizeof(WT);  // fast path: one 16-byte vector per lane
  switch (a.update) {
    case RECEMB_UPD_ADAGRAD:
      launch_one<G, 1, E, GT, WT, true, SegCfg<B, PD, MINB, RECEMB_UPD_ADAGRAD, PLAIN, true>>(a, s);
      return true;
    case RECEMB_UPD_ROWWISE_ADAGRAD:
      launch_one<G, 1, E, GT, WT, true, SegCfg<B, PD, MINB, RECEMB_UPD_ROWWISE_ADAGRAD, PLAIN, true>>(a, s);
      return true;
    case RECEMB_UPD_SGD:
      launch_one<G, 1, E, GT, WT, true, SegCfg<B, PD, MINB, RECEMB_UPD_SGD, PLAIN, true>>(a, s);
      return true;
    case RECEMB_UPD_DENSE_GRAD:
      launch_one<G, 1, E, GT, WT, true, SegCfg<B, PD, MINB, RECEMB_UPD_DENSE_GRAD, PLAIN, true>>(a, s);
      return true;
    default:
      return false;
  }
}

// mostly-unique fast path (seg_pre_kernel): row-wise Adagrad / SGD on rows of 16 or 32 lane-vectors
template <int G, typename T, bool PLAIN, int B, int PD, int MINB>
static bool launch_pre(const SegArgs& a, cudaStream_t s) {
  constexpr int E = 16 / (int)sizeof(T);
  switch (a.update) {
    case RECEMB_UPD_ROWWISE_ADAGRAD:
      launch_kernel<seg_pre_kernel<G, E, T, SegCfg<B, PD, MINB, RECEMB_UPD_ROWWISE_ADAGRAD, PLAIN, true>>, G, true>(a, s);
      return true;
    case RECEMB_UPD_SGD:
      launch_kernel<seg_pre_kernel<G, E, T, SegCfg<B, PD, MINB, RECEMB_UPD_SGD, PLAIN, true>>, G, true>(a, s);
      return true;
    default:
      return false;
  }
}

// RECEMB_SEG_TUNE="B,PD,MINB": tuning aid for the fp32 G = 16 fast path (one GPU session can
// compare instantiations).  Unset = default.
static void tune_params(int* b, int* pd, int* minb) {
  static int tb = -1, tp = 0, tm = 0;
  if (tb == -1) {
    tb = 0;
    const char* e = getenv("RECEMB_SEG_TUNE");
    if (e && sscanf(e, "%d,%d,%d", &tb, &tp, &tm) != 3) tb = 0;
  }
  *b = tb;
  *pd = tp;
  *minb = tm;
}

template <typename GT, typename WT, bool L0>
static int launch_seg(const SegArgs& a, QuadShape shape, cudaStream_t s) {
  bool launched = false;
  // fast path: gradients and table of one dtype, rows of exactly 16 or 32 lane-vectors of 16 bytes
  const int vecs16 = (int)(a.dim * sizeof(WT) / 16);
  if constexpr (L0 && std::is_same<GT, WT>::value)
  if ((a.dim * sizeof(WT)) % 16 == 0 && (vecs16 == 16 || vecs16 == 32)) {
    const bool plain = a.spg == 1 && !a.slot_weight && !a.grad_row_scale && !(a.hp.grad_div > 0.f);
    SegArgs f = a;
    f.vecs = vecs16;
    const SegArgs& a = f;
    static const int pre = env_flag("RECEMB_SEG_PRE", 1);
    if (pre && (a.update == RECEMB_UPD_ROWWISE_ADAGRAD || a.update == RECEMB_UPD_SGD)) {
      int tb, tp, tm;
      tune_params(&tb, &tp, &tm);
      if (vecs16 == 16) {
        if (plain && tb == 4) launched = launch_pre<16, WT, true, 4, 1, 3>(a, s);
        else if (plain && tb == 2 && tm == 3) launched = launch_pre<16, WT, true, 2, 1, 3>(a, s);
        else if (plain && tb == 2 && tp == 0) launched = launch_pre<16, WT, true, 2, 0, 4>(a, s);
        else if (plain && tb == 2 && tp == 2) launched = launch_pre<16, WT, true, 2, 2, 4>(a, s);
        else if (plain) launched = launch_pre<16, WT, true, 2, 1, 4>(a, s);
        else if (tb == 2 && tp == 0) launched = launch_pre<16, WT, false, 2, 0, 4>(a, s);
        else if (tb == 4) launched = launch_pre<16, WT, false, 4, 0, 3>(a, s);
        else launched = launch_pre<16, WT, false, 2, 1, 4>(a, s);
      } else {
        if (plain) launched = launch_pre<32, WT, true, 2, 2, 4>(a, s);
        else launched = launch_pre<32, WT, false, 2, 2, 4>(a, s);
      }
      if (launched) {
        RECEMB_LAUNCHED();
        return RECEMB_OK;
      }
    }
    if (vecs16 == 16) {
      int tb, tp, tm;
      tune_params(&tb, &tp, &tm);
      if (plain && tb == 2 && tp == 1 && tm == 4) launched = launch_fast<16, GT, WT, true, 2, 1, 4>(a, s);
      else if (plain && tb == 2 && tp == 3 && tm == 4) launched = launch_fast<16, GT, WT, true, 2, 3, 4>(a, s);
      else if (plain && tb == 2 && tp == 2 && tm == 5) launched = launch_fast<16, GT, WT, true, 2, 2, 5>(a, s);
      else if (plain && tb == 4 && tp == 1 && tm == 4) launched = launch_fast<16, GT, WT, true, 4, 1, 4>(a, s);
      else if (plain && tb == 1 && tp == 4 && tm == 5) launched = launch_fast<16, GT, WT, true, 1, 4, 5>(a, s);
      else if (plain && tb == 2 && tp == 4 && tm == 5) launched = launch_fast<16, GT, WT, true, 2, 4, 5>(a, s);
      else if (plain && tb == 2 && tp == 2 && tm == 6) launched = launch_fast<16, GT, WT, true, 2, 2, 6>(a, s);
      else if (plain && tb == 2 && tp == 2 && tm == 4) launched = launch_fast<16, GT, WT, true, 2, 2, 4>(a, s);
      else if (plain && tb == 2 && tp == 4 && tm == 4) launched = launch_fast<16, GT, WT, true, 2, 4, 4>(a, s);
      else if (plain && tb == 2 && tp == 8 && tm == 4) launched = launch_fast<16, GT, WT, true, 2, 8, 4>(a, s);
      else if (plain) launched = launch_fast<16, GT, WT, true, 2, 1, 4>(a, s);
      else launched = launch_fast<16, GT, WT, false, 2, 1, 4>(a, s);
    } else {
      // one group per warp: half the rows in flight per warp -> prefetch further ahead
      if (plain) launched = launch_fast<32, GT, WT, true, 2, 4, 4>(a, s);
      else launched = launch_fast<32, GT, WT, false, 2, 4, 4>(a, s);
    }
    if (launched) {
      RECEMB_LAUNCHED();
      return RECEMB_OK;
    }
  }
  if (!launched) {
    SEG_GV(1, 1) SEG_GV(2, 1) SEG_GV(4, 1) SEG_GV(8, 1) SEG_GV(16, 1) SEG_GV(32, 1) SEG_GV(32, 2)
    SEG_GV(32, 4) SEG_GV(32, 8)
  }
  if (!launched) {
    set_error("bwd_apply: no kernel for G=%d V=%d", shape.G, shape.V);
    return RECEMB_ERR_UNSUPPORTED;
  }
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

// level sizes: n_0 = n_slots, n_1 = 2*ceil(n_0/kChunk0), n_{l+1} = 2*ceil(n_l/kChunkN);
// the last level is a single chunk (no boundary => nothing can stay open).
static int level_sizes(int64_t n0, int64_t* sizes, int max_levels) {
  int L = 0;
  int64_t n = n0;
  sizes[L++] = n;
  while (n > (L == 1 ? kChunk0 : kChunkN) && L < max_levels) {
    const int ch = (L == 1) ? kChunk0 : kChunkN;
    n = 2 * ((n + ch - 1) / ch);
    sizes[L++] = n;
  }
  return L;
}
static int64_t records_of(int64_t n, int level) {
  const int ch = level == 0 ? kChunk0 : kChunkN;
  return 2 * ((n + ch - 1) / ch);
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

extern "C" int recemb_bwd_plan(const int64_t* ids, int64_t n_ids, const recemb_layout* layout,
                               int32_t slots_per_id, int hash_mode, int64_t num_rows, int64_t hash_arg, int zero_pad,
                               int64_t pad_id, int64_t pad_row, int32_t bag_size,
                               const int32_t* lengths, int32_t last_n, void* plan,
                               size_t plan_bytes, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n_ids >= 0 && slots_per_id >= 1, "bad n_ids / slots_per_id");
  RECEMB_CHECK_ARG(plan != nullptr && plan_bytes >= kCounterBytes, "plan buffer missing");
  RECEMB_CHECK_ARG((uintptr_t)plan % 256 == 0, "plan buffer must be 256-byte aligned");
  RECEMB_CHECK_ARG(num_rows >= 1, "num_rows < 1");
  // keys are rows of the stacked (table-batched) and / or local (sharded) table
  const int64_t total_rows = layout_local_rows(num_rows, layout) * layout_tables(layout, n_ids);
  RECEMB_CHECK_ARG(!(layout && layout->ids_per_table > 0) || slots_per_id == 1,
                   "table batching and k-shift cannot be combined");
  RECEMB_UNSUPPORTED(total_rows < 0xfffffff0ll, "%lld rows do not fit 32-bit sort keys",
                     (long long)total_rows);
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
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(ids != nullptr, "ids null");
  PlanLayout L;
  RECEMB_CUDA(plan_layout(n, total_rows, &L));
  if (plan_bytes < L.total) {
    set_error("plan buffer %zu < required %zu", plan_bytes, L.total);
    return RECEMB_ERR_WORKSPACE;
  }
  PlanKeyArgs a;
  a.ids = ids;
  a.n_slots = n;
  a.slots_per_id = slots_per_id;
  int rc = make_hash_spec(hash_mode, num_rows, slots_per_id > 1 ? 0 : hash_arg, &a.h, layout);
  if (rc) return rc;
  a.zero_pad = zero_pad;
  a.pad_id = pad_id;
  a.pad_row = pad_row;
  a.bag_size = bag_size;
  a.lengths = lengths;
  a.last_n = last_n;
  a.sentinel = (uint32_t)total_rows;
  if (a.h.out_feats) {
    RECEMB_CHECK_ARG(bag_size > 0 && slots_per_id == 1 && layout->ids_per_table % bag_size == 0,
                     "out_features needs pooled bags (bag_size > 0) with ids_per_table a multiple of bag_size");
    a.h.out_bpt = (uint32_t)(layout->ids_per_table / bag_size);
    RECEMB_UNSUPPORTED((int64_t)a.h.out_bpt * a.h.out_feats * bag_size < 0xffffffffll, "too many gradient slots");
  }
  // the sort ping-pongs between the two pair buffers and always ends in the "out" one
  const bool start_in_out = sort_input_in_b(sort_shape(n, L.key_bits, device));
  a.keys = (uint32_t*)(base + (start_in_out ? L.off_keys_out : L.off_keys_in));
  a.vals = (uint32_t*)(base + (start_in_out ? L.off_vals_out : L.off_vals_in));
  const int sms = sm_count(device);
  int64_t grid = (n + (int64_t)kBwdThreads * kPlanUnroll - 1) / ((int64_t)kBwdThreads * kPlanUnroll);
  if (grid > (int64_t)sms * 16) grid = (int64_t)sms * 16;
  plan_keys_kernel<<<(unsigned)grid, kBwdThreads, 0, s>>>(a);
  RECEMB_LAUNCHED();
  return sort_pairs((uint32_t*)(base + L.off_keys_in), (uint32_t*)(base + L.off_vals_in),
                    (uint32_t*)(base + L.off_keys_out), (uint32_t*)(base + L.off_vals_out), n, L.key_bits,
                    base + L.off_temp, L.temp_bytes, device, s);
}

extern "C" int recemb_plan_count(void* plan, size_t plan_bytes, int64_t n_slots, int64_t num_rows,
                                 int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(plan != nullptr && n_slots >= 0, "bad plan / n_slots");
  const size_t arr = align_up((size_t)n_slots * 4, 256);
  RECEMB_CHECK_ARG(plan_bytes >= kCounterBytes + 4 * arr, "plan buffer too small for n_slots");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  cudaStream_t s = (cudaStream_t)stream;
  char* base = (char*)plan;
  RECEMB_CUDA(cudaMemsetAsync(base, 0, kCounterBytes, s));
  if (n_slots == 0) return RECEMB_OK;
  int64_t grid = (n_slots + kBwdThreads - 1) / kBwdThreads;
  const int64_t cap = (int64_t)sm_count(device) * 16;
  if (grid > cap) grid = cap;
  plan_count_kernel<<<(unsigned)grid, kBwdThreads, 0, s>>>(
      (const uint32_t*)(base + kCounterBytes + 2 * arr), n_slots, (uint32_t)num_rows,
      (unsigned long long*)base);
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
  const int64_t last = records_of(sizes[L - 1], L - 1);
  total += align_up((size_t)last * 4, 256) + align_up((size_t)last * dim * 4, 256);
  return total;
}

extern "C" int recemb_bwd_apply(const void* plan, size_t plan_bytes, int64_t n_slots, const void* grad,
                                int grad_dtype, int64_t grad_rows, int32_t dim,
                                int32_t slots_per_grad_row, const float* slot_weight,
                                const float* grad_row_scale, int update, void* table, int dtype,
                                int64_t num_rows, void* state1, void* state2,
                                const recemb_optim_params* hp_host, void* workspace,
                                size_t workspace_bytes, int device, recemb_stream_t stream) {
  return recemb_bwd_apply_guarded(plan, plan_bytes, n_slots, grad, grad_dtype, grad_rows, dim, slots_per_grad_row,
                                  slot_weight, grad_row_scale, update, table, dtype, num_rows, state1, state2, hp_host,
                                  workspace, workspace_bytes, nullptr, device, stream);
}

static int apply_impl(const void* plan, size_t plan_bytes, int64_t n_slots, const void* grad, int grad_dtype,
                      int64_t grad_rows, int32_t dim, int32_t slots_per_grad_row, const float* slot_weight,
                      const float* grad_row_scale, int update, void* table, int dtype, int64_t num_rows,
                      void* state1, void* state2, const recemb_optim_params* hp_host, void* workspace,
                      size_t workspace_bytes, const uint32_t* skip_if_nonzero, int device, recemb_stream_t stream,
                      const PeerGate* gate);

extern "C" int recemb_bwd_apply_guarded(const void* plan, size_t plan_bytes, int64_t n_slots, const void* grad,
                                        int grad_dtype, int64_t grad_rows, int32_t dim,
                                        int32_t slots_per_grad_row, const float* slot_weight,
                                        const float* grad_row_scale, int update, void* table, int dtype,
                                        int64_t num_rows, void* state1, void* state2,
                                        const recemb_optim_params* hp_host, void* workspace,
                                        size_t workspace_bytes, const uint32_t* skip_if_nonzero, int device,
                                        recemb_stream_t stream) {
  return apply_impl(plan, plan_bytes, n_slots, grad, grad_dtype, grad_rows, dim, slots_per_grad_row, slot_weight,
                    grad_row_scale, update, table, dtype, num_rows, state1, state2, hp_host, workspace,
                    workspace_bytes, skip_if_nonzero, device, stream, nullptr);
}

static int peer_bwd_fused_impl(const recemb_peer_group* group, const recemb_peer_arena* arena, const void* plan,
                               size_t plan_bytes, const void* my_grad, int32_t tables, int64_t bags_per_table,
                               int32_t dim, int dtype, int update, void* table, int64_t total_rows,
                               int64_t rows_per_table, void* state1, const recemb_optim_params* hp, void* workspace,
                               size_t workspace_bytes, int32_t push_ctas, int partition, int device,
                               recemb_stream_t stream);

extern "C" int recemb_peer_bwd_apply_fused(const recemb_peer_group* group, const recemb_peer_arena* arena,
                                           const void* plan, size_t plan_bytes, const void* my_grad, int32_t tables,
                                           int64_t bags_per_table, int32_t dim, int dtype, int update, void* table,
                                           int64_t total_rows, int64_t rows_per_table, void* state1,
                                           const recemb_optim_params* hp, void* workspace, size_t workspace_bytes,
                                           int32_t push_ctas, int device, recemb_stream_t stream) {
  return peer_bwd_fused_impl(group, arena, plan, plan_bytes, my_grad, tables, bags_per_table, dim, dtype, update, table,
                             total_rows, rows_per_table, state1, hp, workspace, workspace_bytes, push_ctas, 0, device,
                             stream);
}

extern "C" int recemb_peer_bwd_apply_fused_tablewise(const recemb_peer_group* group, const recemb_peer_arena* arena,
                                                     const void* plan, size_t plan_bytes, const void* my_grad,
                                                     int32_t tables, int64_t bags_per_table, int32_t dim, int dtype,
                                                     int update, void* table, int64_t total_rows,
                                                     int64_t rows_per_table, void* state1,
                                                     const recemb_optim_params* hp, void* workspace,
                                                     size_t workspace_bytes, int32_t push_ctas, int device,
                                                     recemb_stream_t stream) {
  return peer_bwd_fused_impl(group, arena, plan, plan_bytes, my_grad, tables, bags_per_table, dim, dtype, update, table,
                             total_rows, rows_per_table, state1, hp, workspace, workspace_bytes, push_ctas, 1, device,
                             stream);
}

static int peer_bwd_fused_impl(const recemb_peer_group* group, const recemb_peer_arena* arena, const void* plan,
                               size_t plan_bytes, const void* my_grad, int32_t tables, int64_t bags_per_table,
                               int32_t dim, int dtype, int update, void* table, int64_t total_rows,
                               int64_t rows_per_table, void* state1, const recemb_optim_params* hp, void* workspace,
                               size_t workspace_bytes, int32_t push_ctas, int partition, int device,
                               recemb_stream_t stream) {
  RECEMB_CHECK_ARG(group && arena && my_grad, "null peer group / arena / gradients");
  RECEMB_CHECK_ARG(group->world >= 1 && group->world <= RECEMB_MAX_PEERS && group->rank >= 0 &&
                       group->rank < group->world, "peer group world / rank out of range");
  RECEMB_CHECK_ARG(tables >= 1 && tables <= kGateTables, "tables %d outside [1, %d]", tables, kGateTables);
  RECEMB_CHECK_ARG(bags_per_table >= 1 && (int64_t)tables * bags_per_table == arena->bags_total,
                   "tables x bags_per_table != arena bags_total");
  const int local_tables = partition ? (tables - group->rank + group->world - 1) / group->world : tables;
  RECEMB_CHECK_ARG(rows_per_table >= 1 && rows_per_table < 0xffffffffll && rows_per_table * local_tables == total_rows,
                   "rows_per_table x (local) tables != total_rows");
  RECEMB_CHECK_ARG(push_ctas >= 1 && push_ctas <= 1024, "push_ctas %d outside [1, 1024]", push_ctas);
  RECEMB_CHECK_ARG(dtype == RECEMB_F32 || dtype == RECEMB_BF16, "bad dtype");
  RECEMB_CHECK_ARG((uintptr_t)my_grad % 16 == 0, "gradients must be 16-byte aligned");
  const int64_t row_bytes = (int64_t)dim * (dtype == RECEMB_F32 ? 4 : 2);
  RECEMB_UNSUPPORTED((row_bytes == 256 || row_bytes == 512) &&
                         (update == RECEMB_UPD_ROWWISE_ADAGRAD || update == RECEMB_UPD_SGD) &&
                         env_flag("RECEMB_SEG_PRE", 1) != 0,
                     "fused push: rows of 256 / 512 bytes with row-wise Adagrad or SGD only");
  if (partition && push_ctas % tables != 0) push_ctas = (push_ctas / tables + 1) * tables;  // tables side by side
  PeerGate g;
  g.push_ctas = push_ctas;
  g.world = group->world;
  g.rank = group->rank;
  g.tables = tables;
  for (int i = 0; i < RECEMB_MAX_PEERS; ++i) g.arena[i] = nullptr;
  for (int i = 0; i < group->world; ++i) {
    RECEMB_CHECK_ARG(group->arena[i], "peer arena %d not mapped", i);
    g.arena[i] = (char*)group->arena[i];
  }
  g.off_grads = arena->off_grads;
  g.off_gate = arena->off_gate;
  g.src = (const uint4*)my_grad;
  g.vecs_per_table = bags_per_table * (row_bytes / 16);
  g.sender_vecs = arena->bags_total * (row_bytes / 16);
  g.rows_per_table = (uint32_t)rows_per_table;
  {
    const char* e = getenv("RECEMB_PEER_BARRIER_TIMEOUT_S");
    double sec = e ? atof(e) : 600.0;
    if (!(sec > 0.0)) sec = 600.0;
    g.timeout_cycles = (long long)(sec * 2.0e9);
  }
  char* mine = g.arena[g.rank];
  g.status = (uint32_t*)(mine + arena->off_status);
  g.partition = partition;
  g.local_tables = local_tables;
  g.debug = env_flag("RECEMB_GATE_DEBUG", 0);
  DeviceGuard dg(device);
  RECEMB_CUDA(dg.err);
  gate_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((uint64_t*)(mine + arena->off_gate));
  RECEMB_LAUNCHED();
  const int64_t n = (int64_t)group->world * arena->cap;
  return apply_impl(plan, plan_bytes, n, mine + arena->off_grads, dtype, (int64_t)group->world * arena->bags_total, dim,
                    1, nullptr, nullptr, update, table, dtype, total_rows, state1, nullptr, hp, workspace,
                    workspace_bytes, g.status, device, stream, &g);
}

static int apply_impl(const void* plan, size_t plan_bytes, int64_t n_slots, const void* grad, int grad_dtype,
                      int64_t grad_rows, int32_t dim, int32_t slots_per_grad_row, const float* slot_weight,
                      const float* grad_row_scale, int update, void* table, int dtype, int64_t num_rows,
                      void* state1, void* state2, const recemb_optim_params* hp_host, void* workspace,
                      size_t workspace_bytes, const uint32_t* skip_if_nonzero, int device, recemb_stream_t stream,
                      const PeerGate* gate) {
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
  RECEMB_CHECK_ARG(n_slots >= 0, "n_slots < 0");
  const int64_t n = n_slots;  // entries of the plan (== grad_rows * slots_per_grad_row for plans
                              // built from ids; free for plans built from routed entries)
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
  int L = level_sizes(n, sizes, kMaxLevels);
  RECEMB_CUDA(cudaMemsetAsync(workspace, 0, 256, s));  // per-level "records emitted" flags
  const char* pbase = (const char*)plan;
  char* w = (char*)workspace;
  size_t woff = 256;

  SegArgs a;
  a.dim = (uint32_t)dim;
  a.vecs = dim / 4;  // generic kernels: E = 4 (the fast path overrides it)
  a.spg = (uint32_t)slots_per_grad_row;
  a.spg_magic = (uint32_t)((1ull << 32) / (uint64_t)slots_per_grad_row);
  a.slot_weight = slot_weight;
  a.grad_row_scale = grad_row_scale;
  a.sentinel = (uint32_t)num_rows;
  a.update = update;
  a.table = table;
  a.state1 = (float*)state1;
  a.state2 = (float*)state2;
  if (hp_host) a.hp = *hp_host;
  else a.hp = recemb_optim_params{};
  a.chunk = kChunk0;
  a.guard = skip_if_nonzero;
  a.gate.push_ctas = 0;
  int32_t chunk_used = kChunk0;  // per call, no state carried between calls or threads
  a.chunk_used = &chunk_used;
  {
    static const int pf = env_flag("RECEMB_SEG_PF_BULK", 0);
    a.pf_bulk = pf;
  }

  const uint32_t* in_keys = (const uint32_t*)(pbase + kCounterBytes + 2 * arr);
  const uint32_t* in_slots = (const uint32_t*)(pbase + kCounterBytes + 3 * arr);
  const void* in_grad = grad;
  for (int l = 0; l < L; ++l) {
    const int64_t recs = records_of(sizes[l], l);
    uint32_t* out_keys = (uint32_t*)(w + woff);
    woff += align_up((size_t)recs * 4, 256);
    float* out_part = (float*)(w + woff);
    woff += align_up((size_t)recs * dim * 4, 256);
    a.keys = in_keys;
    a.slots = in_slots;
    a.n = (int32_t)sizes[l];
    a.grad = in_grad;
    a.out_keys = out_keys;
    a.out_partials = out_part;
    a.flag_in = (const uint32_t*)w + l;  // flags live in the first 256 bytes of the workspace
    a.flag_out = (uint32_t*)w + l + 1;
    int rc;
    if (l == 0 && t_ev_start) cudaEventRecord(t_ev_start, s);
    if (gate) {  // level 0 only: the first CTAs push my gradients, the chunks are gated per table
      if (l == 0) a.gate = *gate;
      else a.gate.push_ctas = 0;
    }
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
    if (l == 0 && chunk_used > kChunk0) {
      // level 0 used longer chunks: fewer records than the (upper-bound) layout reserves;
      // the levels above shrink accordingly (offsets stay inside the reserved areas)
      sizes[1] = 2 * ((n + chunk_used - 1) / chunk_used);
      int64_t m = sizes[1];
      int LL = 2;
      while (m > kChunkN && LL < kMaxLevels) {
        m = 2 * ((m + kChunkN - 1) / kChunkN);
        sizes[LL++] = m;
      }
      if (L > 1) L = LL;
    }
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
