// Forward kernels of the embedding hot path (sm_100a):
//   row_index_kernel  - id -> row hashing                     (a2; commons/layers.py:174-185)
//   gather_kernel     - sequence gather, optional 2nd table   (a1, a4, a6; commons/layers.py:56-61, :115-123)
//   kshift_kernel     - fused k-shift bag                     (a3; commons/layers.py:152-172)
//   pool_kernel       - pooled multi-hot bag sum/mean/last-N  (a5, a11; commons/transformers/layers.py:457-469)
//
// All of them are HBM-bound byte movers: a row is moved as 16-byte vectors by a
// group of G consecutive lanes (G*V vectors per row), id tiles are staged into
// shared memory by the TMA engine's 1-D bulk copy (cp.async.bulk -> UBLKCP)
// double-buffered against the gather, table rows are read with
// ld.global.nc.L1::no_allocate and write-once outputs leave with st.global.cs.
#include <cstdlib>

#include "common.cuh"

namespace recemb {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTileIds = 1024;  // ids per staged tile (8 KB)

// ------------------------------------------------------------ id staging ----
// Double-buffered id tiles.  Thread 0 issues the bulk copy of the NEXT tile
// while the CTA gathers the current one.  If the id pointer is not 16-byte
// aligned the CTA falls back to cooperative plain loads (same results).
struct IdStager {
  int64_t* buf[2];
  uint64_t* bar;  // [2]
  uint32_t phase[2];
  bool bulk;

  __device__ __forceinline__ void init(int64_t* b0, int64_t* b1, uint64_t* bars, bool use_bulk) {
    buf[0] = b0;
    buf[1] = b1;
    bar = bars;
    phase[0] = phase[1] = 0;
    bulk = use_bulk;
    if (bulk && threadIdx.x == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      mbar_fence_init();
    }
    __syncthreads();
  }
  // thread 0 only (bulk mode); `count` ids starting at `src`
  __device__ __forceinline__ void issue(int b, const int64_t* src, int count) {
    if (!bulk || threadIdx.x != 0) return;
    fence_proxy_async();  // earlier generic-proxy accesses to buf[b] are ordered before the async write
    const int even = count & ~1;
    if (count & 1) buf[b][count - 1] = src[count - 1];
    mbar_expect_tx(&bar[b], (uint32_t)even * 8u);
    if (even) bulk_g2s(buf[b], src, (uint32_t)even * 8u, &bar[b]);
  }
  // all threads
  __device__ __forceinline__ void acquire(int b, const int64_t* src, int count) {
    if (bulk) {
      mbar_wait(&bar[b], phase[b]);
      phase[b] ^= 1;
    } else {
      for (int i = threadIdx.x; i < count; i += kThreads) buf[b][i] = src[i];
      __syncthreads();
    }
  }
};

// ------------------------------------------------------------- row index ----
__global__ void __launch_bounds__(kThreads) row_index_kernel(const int64_t* __restrict__ ids,
                                                             int64_t n, HashSpec h,
                                                             int64_t* __restrict__ rows) {
  int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (; i < n; i += stride) rows[i] = row_of(ids[i], h);
}

// ----------------------------------------------------------------- gather ----
struct GatherArgs {
  const uint4* table;
  const uint4* table2;
  const int64_t* ids;
  uint4* out;
  float* inv_norm;
  int64_t n;
  int32_t row_vecs;
  HashSpec h1, h2;
  int zero_pad;
  int64_t pad_id;
  int bulk_ok;
  int prefetch_iters;  // L2 prefetch distance in warp iterations (0 = off)
  int l1_rows;         // rows may stay in L1 (skewed ids: hot rows are re-read by every SM)
};

// MAP: the output position is remapped (flipped sequences and / or a sequence window); the plain lookup
// (MAP = false, the cfg 2 headline) carries none of that code or its registers
template <int G, int V, typename T, int EPI, bool TWO, bool MAP = true>
__global__ void __launch_bounds__(kThreads) gather_kernel(const GatherArgs a) {
  constexpr int RPW = 32 / G;                  // rows per warp pass
  constexpr int UNROLL = (V >= 4) ? 2 : (V == 2 ? 2 : 4);
  constexpr int E = Vec16<T>::kElems;
  __shared__ alignas(16) int64_t s_ids[2][kTileIds];
  __shared__ alignas(16) int64_t s_rows2[TWO ? kTileIds : 1];
  __shared__ alignas(8) uint64_t s_bar[2];

  const int64_t num_tiles = (a.n + kTileIds - 1) / kTileIds;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lig = lane % G, gi = lane / G;

  IdStager st;
  st.init(s_ids[0], s_ids[1], s_bar, a.bulk_ok != 0);
  const uint32_t keep = MAP ? window_keep(a.h1) : 0u;

  int64_t tile = blockIdx.x;
  int b = 0;
  auto tile_count = [&](int64_t t) { return (int)min((int64_t)kTileIds, a.n - t * kTileIds); };
  if (tile < num_tiles) st.issue(0, a.ids + tile * kTileIds, tile_count(tile));

  for (; tile < num_tiles; tile += gridDim.x, b ^= 1) {
    const int64_t next = tile + gridDim.x;
    if (next < num_tiles) st.issue(b ^ 1, a.ids + next * kTileIds, tile_count(next));
    const int cnt = tile_count(tile);
    st.acquire(b, a.ids + tile * kTileIds, cnt);

    // hash in place: s_ids[b][i] <- row (-1 for a zero-filled pad position, -2 for a position outside
    // the sequence window: not read, not written)
    int64_t* rows = s_ids[b];
    for (int i = threadIdx.x; i < cnt; i += kThreads) {
      const int64_t id = rows[i];
      const bool pad = a.zero_pad && id == a.pad_id;
      if (TWO) s_rows2[i] = pad ? -1 : row_of(id, a.h2);
      const int64_t r1 = pad ? -1 : row_of(id, a.h1);  // -1: pad position or out-of-range identity id
      int64_t r = r1 < 0 ? -1 : r1 + table_offset(tile * kTileIds + i, a.h1);
      if (MAP && a.h1.win_len && out_row(tile * kTileIds + i, a.h1, keep) < 0) {
        r = -2;
        if (TWO) s_rows2[i] = -1;
      }
      rows[i] = r;
    }
    __syncthreads();

    uint4* out_tile = a.out + tile * kTileIds * (int64_t)a.row_vecs;
    for (int base = warp * RPW; base < cnt; base += kWarps * RPW * UNROLL) {
      // pull the rows of a later iteration of this warp into L2 (row indices of the whole tile
      // are already in shared memory): the loads below then mostly hit L2
      if (a.prefetch_iters > 0) {
        const int pbase = base + a.prefetch_iters * kWarps * RPW * UNROLL;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const int pl = pbase + u * kWarps * RPW + gi;
          if (pl < cnt) {
            const int64_t r = rows[pl];
#pragma unroll
            for (int j = 0; j < V; ++j) {
              const int vec = j * G + lig;
              if (r >= 0 && vec < a.row_vecs)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a.table + r * a.row_vecs + vec));
            }
          }
        }
      }
      uint4 v[UNROLL][V];
      uint4 w[TWO ? UNROLL : 1][TWO ? V : 1];
      int l[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        l[u] = base + u * kWarps * RPW + gi;
        const int64_t r = (l[u] < cnt) ? rows[l[u]] : -1;
        const int64_t r2 = (TWO && l[u] < cnt) ? s_rows2[l[u]] : -1;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int vec = j * G + lig;
          v[u][j] = make_uint4(0, 0, 0, 0);
          if (r >= 0 && vec < a.row_vecs)
            v[u][j] = a.l1_rows ? ldg_nc_l1_v4(a.table + r * a.row_vecs + vec) : ldg_nc_v4(a.table + r * a.row_vecs + vec);
          if (TWO) {
            w[u][j] = make_uint4(0, 0, 0, 0);
            if (r2 >= 0 && vec < a.row_vecs) w[u][j] = ldg_nc_v4(a.table2 + r2 * a.row_vecs + vec);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (TWO || EPI == RECEMB_EPI_L2NORM) {
          float f[V][E];
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < V; ++j) {
            Vec16<T>::unpack(v[u][j], f[j]);
            if (TWO) {
              float g[E];
              Vec16<T>::unpack(w[u][j], g);
#pragma unroll
              for (int e = 0; e < E; ++e) f[j][e] += g[e];
              if (sizeof(T) == 2) {  // the reference add rounds to the table dtype
                uint4 t = Vec16<T>::pack(f[j]);
                Vec16<T>::unpack(t, f[j]);
              }
            }
#pragma unroll
            for (int e = 0; e < E; ++e) ss += f[j][e] * f[j][e];
          }
          if (EPI == RECEMB_EPI_L2NORM) {
            ss = group_sum<G>(ss);
            const float denom = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
            for (int j = 0; j < V; ++j)
#pragma unroll
              for (int e = 0; e < E; ++e) f[j][e] = f[j][e] / denom;
            if (a.inv_norm && l[u] < cnt && lig == 0) {
              const int64_t o = MAP ? out_row(tile * kTileIds + l[u], a.h1, keep) : tile * kTileIds + l[u];
              if (o >= 0) a.inv_norm[o] = 1.f / denom;
            }
          }
#pragma unroll
          for (int j = 0; j < V; ++j) v[u][j] = Vec16<T>::pack(f[j]);
        }
        if (l[u] < cnt) {
          uint4* dst = out_tile + (int64_t)l[u] * a.row_vecs;
          if constexpr (MAP) {
            const int64_t o = out_row(tile * kTileIds + l[u], a.h1, keep);
            dst = o < 0 ? nullptr : a.out + o * a.row_vecs;
          }
          if (dst) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
              const int vec = j * G + lig;
              if (vec < a.row_vecs) stg_cs_v4(dst + vec, v[u][j]);
            }
          }
        }
      }
    }
    __syncthreads();  // tile fully consumed before its buffer is refilled
  }
}

// ------------------------------------------------------------ k-shift bag ----
struct KShiftArgs {
  const uint4* table;
  const int64_t* ids;
  uint4* out;
  float* inv_norm;
  int64_t n;
  int32_t row_vecs;
  int32_t k;
  ModN mod_rows;
  int epilogue;
  float sqrt_k;
  int bulk_ok;
  HashSpec h;  // flip_len / sequence window only (out_row)
  int l1_hot;  // rows of shifts >= 1 may stay in L1
};

template <int G, int V, typename T>
__global__ void __launch_bounds__(kThreads) kshift_kernel(const KShiftArgs a) {
  constexpr int RPW = 32 / G;
  constexpr int E = Vec16<T>::kElems;
  constexpr int BATCH = (V == 1) ? 8 : (V == 2 ? 4 : 2);  // row loads in flight per lane
  __shared__ alignas(16) int64_t s_ids[2][kTileIds];
  __shared__ alignas(8) uint64_t s_bar[2];

  const int64_t num_tiles = (a.n + kTileIds - 1) / kTileIds;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lig = lane % G, gi = lane / G;

  IdStager st;
  st.init(s_ids[0], s_ids[1], s_bar, a.bulk_ok != 0);
  const uint32_t keep = window_keep(a.h);
  int64_t tile = blockIdx.x;
  int b = 0;
  auto tile_count = [&](int64_t t) { return (int)min((int64_t)kTileIds, a.n - t * kTileIds); };
  if (tile < num_tiles) st.issue(0, a.ids + tile * kTileIds, tile_count(tile));

  for (; tile < num_tiles; tile += gridDim.x, b ^= 1) {
    const int64_t next = tile + gridDim.x;
    if (next < num_tiles) st.issue(b ^ 1, a.ids + next * kTileIds, tile_count(next));
    const int cnt = tile_count(tile);
    st.acquire(b, a.ids + tile * kTileIds, cnt);
    const int64_t* ids = s_ids[b];

    for (int base = warp * RPW; base < cnt; base += kWarps * RPW) {
      const int l = base + gi;
      const int64_t orow = l < cnt ? out_row(tile * kTileIds + l, a.h, keep) : -1;
      const bool live = orow >= 0;  // inside the tile and inside the sequence window
      const int64_t id = live ? ids[l] : 0;
      float acc[V][E];
#pragma unroll
      for (int j = 0; j < V; ++j)
#pragma unroll
        for (int e = 0; e < E; ++e) acc[j][e] = 0.f;

      // shifts are hashed G at a time (lane `lig` hashes shift c0+lig) and
      // broadcast inside the group; rows are then loaded BATCH at a time and
      // added in the order c = 0, 1, ... (the reference's order of adds).
      for (int c0 = 0; c0 < a.k; c0 += G) {
        const int my_c = c0 + lig;
        const int64_t my_row = (my_c < a.k) ? kshift_row(id, my_c, a.mod_rows) : 0;
        const int span = min(G, a.k - c0);
        for (int cb = 0; cb < span; cb += BATCH) {
          uint4 v[BATCH][V];
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            const int cc = cb + u;
            const int64_t r = __shfl_sync(0xffffffffu, my_row, gi * G + (cc < G ? cc : 0));
#pragma unroll
            for (int j = 0; j < V; ++j) {
              const int vec = j * G + lig;
              v[u][j] = make_uint4(0, 0, 0, 0);
              if (live && cc < span && vec < a.row_vecs) {
                // shift c >= 1 sends every negative id to one of 2^(c-1) rows (commons/layers.py:182,
                // arithmetic >>): those rows are read by every SM all the time -> keep them in L1
                if (a.l1_hot && c0 + cc > 0) v[u][j] = ldg_nc_l1_v4(a.table + r * a.row_vecs + vec);
                else v[u][j] = ldg_nc_v4(a.table + r * a.row_vecs + vec);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            if (cb + u < span) {
#pragma unroll
              for (int j = 0; j < V; ++j) {
                float f[E];
                Vec16<T>::unpack(v[u][j], f);
#pragma unroll
                for (int e = 0; e < E; ++e) acc[j][e] += f[e];
              }
            }
          }
        }
      }

      if (a.epilogue == RECEMB_EPI_L2NORM) {
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
          for (int e = 0; e < E; ++e) ss += acc[j][e] * acc[j][e];
        ss = group_sum<G>(ss);
        const float denom = fmaxf(sqrtf(ss), 1e-12f);
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
          for (int e = 0; e < E; ++e) acc[j][e] = acc[j][e] / denom;
        if (a.inv_norm && live && lig == 0) a.inv_norm[orow] = 1.f / denom;
      } else if (a.epilogue == RECEMB_EPI_RSQRT_K) {
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
          for (int e = 0; e < E; ++e) acc[j][e] = acc[j][e] / a.sqrt_k;
      }
      if (live) {
        uint4* dst = a.out + orow * (int64_t)a.row_vecs;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int vec = j * G + lig;
          if (vec < a.row_vecs) stg_cs_v4(dst + vec, Vec16<T>::pack(acc[j]));
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------- pooled bag ----
struct PoolArgs {
  const uint4* table;
  const int64_t* ids;
  const int32_t* lengths;
  const float* slot_weight;
  uint4* out;
  int64_t num_bags;
  int32_t bag_size;
  int32_t last_n;
  int32_t row_vecs;
  int32_t bags_per_tile;
  HashSpec h;
  int pool_mode;
  int zero_pad;
  int64_t pad_id;
  int bulk_ok;
  // PEER: global row r of a table lives on rank r % world (h.mod_world) at local row r / world of
  // peer_table[r % world] (every rank's stacked shard, peer-mapped over NVLink); table == nullptr
  const uint4* peer_table[RECEMB_MAX_PEERS];
  int64_t rows_div;  // num_rows / world: local rows of owner o = rows_div + (o < rows_rem)
  uint32_t rows_rem;
  int prefetch;      // local tables: rows of the bag this group handles `prefetch` iterations later go to L2 now
};

constexpr int kPoolTileIds = 2048;

// The walk over a bag is a chain (ids -> hash -> row loads -> accumulate) that a warp repeats bag after
// bag: with ragged bags only a few row loads are in flight per group and the kernel is bound by DRAM
// latency, not bandwidth (ncu, cfg 3: issue-active 49 %, 8 long-scoreboard stalls per issue, DRAM 35 %).
// The ids of the whole tile are already in shared memory, so every group hashes the slots of the bag it
// will process in its NEXT iteration and requests their rows with prefetch.global.L2 (no registers held):
// the real loads one iteration later find them in L2.
template <int G, int V, typename T, bool PEER, int BATCH_ = 0>
__global__ void __launch_bounds__(kThreads, (V <= 2) ? 4 : 1) pool_kernel(const PoolArgs a) {
  constexpr int RPW = 32 / G;
  constexpr int E = Vec16<T>::kElems;
  constexpr int BATCH = BATCH_ > 0 ? BATCH_ : ((V == 1) ? 8 : (V == 2 ? 4 : 2));
  __shared__ alignas(16) int64_t s_ids[2][kPoolTileIds];
  __shared__ alignas(8) uint64_t s_bar[2];
  __shared__ const uint4* s_peer[PEER ? RECEMB_MAX_PEERS : 1];
  if constexpr (PEER) {
    if (threadIdx.x < RECEMB_MAX_PEERS) s_peer[threadIdx.x] = a.peer_table[threadIdx.x];
  }  // published by the __syncthreads() of IdStager::init

  const int64_t num_tiles = (a.num_bags + a.bags_per_tile - 1) / a.bags_per_tile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lig = lane % G, gi = lane / G;
  const int P = a.bag_size;

  IdStager st;
  st.init(s_ids[0], s_ids[1], s_bar, a.bulk_ok != 0);
  int64_t tile = blockIdx.x;
  int b = 0;
  auto tile_bags = [&](int64_t t) {
    return (int)min((int64_t)a.bags_per_tile, a.num_bags - t * a.bags_per_tile);
  };
  if (tile < num_tiles)
    st.issue(0, a.ids + tile * a.bags_per_tile * (int64_t)P, tile_bags(tile) * P);

  for (; tile < num_tiles; tile += gridDim.x, b ^= 1) {
    const int64_t next = tile + gridDim.x;
    if (next < num_tiles)
      st.issue(b ^ 1, a.ids + next * a.bags_per_tile * (int64_t)P, tile_bags(next) * P);
    const int nb = tile_bags(tile);
    st.acquire(b, a.ids + tile * a.bags_per_tile * (int64_t)P, nb * P);
    const int64_t* ids = s_ids[b];

    for (int base = warp * RPW; base < nb; base += kWarps * RPW) {
      const int lb = base + gi;  // bag inside the tile
      const bool live = lb < nb;
      const int64_t bag = tile * a.bags_per_tile + lb;
      int hi = P, lo = 0;
      if (live && a.lengths) hi = min(max(a.lengths[bag], 0), P);
      if (a.last_n > 0) lo = max(0, hi - a.last_n);
      if (!live) hi = lo = 0;
      if constexpr (!PEER) {
        const int nlb = lb + a.prefetch * kWarps * RPW;  // this group's bag `prefetch` iterations ahead
        if (a.prefetch && nlb < nb) {
          const int64_t nbag = tile * a.bags_per_tile + nlb;
          int nhi = P, nlo = 0;
          if (a.lengths) nhi = min(max(a.lengths[nbag], 0), P);
          if (a.last_n > 0) nlo = max(0, nhi - a.last_n);
          for (int p = nlo + lig; p < nhi; p += G) {
            const int64_t id = ids[nlb * P + p];
            int64_t row = (a.zero_pad && id == a.pad_id) ? -1 : row_of(id, a.h);
            if (row >= 0) row = shard_local_row(row, a.h);
            if (row >= 0) {
              const uint4* src = a.table + (row + table_offset(nbag * P + p, a.h)) * a.row_vecs;
              for (int v8 = 0; v8 < a.row_vecs; v8 += 8)  // one request per 128-byte line of the row
                asm volatile("prefetch.global.L2 [%0];" ::"l"(src + v8));
            }
          }
        }
      }
      // the two groups of a warp may have different windows: iterate to the warp max
      int span = hi - lo;
#pragma unroll
      for (int o = 16; o >= G && o > 0; o >>= 1) span = max(span, __shfl_xor_sync(0xffffffffu, span, o));

      float acc[V][E];
#pragma unroll
      for (int j = 0; j < V; ++j)
#pragma unroll
        for (int e = 0; e < E; ++e) acc[j][e] = 0.f;
      int pooled = 0;

      for (int p0 = 0; p0 < span; p0 += G) {
        // lane `lig` hashes slot lo+p0+lig
        const int my_p = lo + p0 + lig;
        const bool my_ok = live && my_p < hi;
        const int64_t my_id = my_ok ? ids[lb * P + my_p] : 0;
        const bool my_use = my_ok && !(a.zero_pad && my_id == a.pad_id);
        // address of the slot's row (0 = nothing to pool), shuffled to the lanes that load it
        uint64_t my_src = 0;
        if constexpr (PEER) {
          const int64_t my_grow = my_use ? row_of(my_id, a.h) : -1;
          if (my_grow >= 0) {
            uint64_t q;
            const uint32_t o = (uint32_t)udivmod((uint64_t)my_grow, a.h.mod_world, &q);
            if (a.h.ids_per_table) {
              uint32_t t = (uint32_t)(bag * P + my_p) / a.h.ids_per_table;
              if (a.h.num_tables) t %= a.h.num_tables;
              q += (uint64_t)t * (uint64_t)(a.rows_div + (o < a.rows_rem ? 1 : 0));
            }
            my_src = (uint64_t)(s_peer[o] + q * (uint64_t)a.row_vecs);
          }
        } else {
          int64_t my_row = my_use ? row_of(my_id, a.h) : -1;
          if (my_row >= 0) my_row = shard_local_row(my_row, a.h);  // -1: another rank owns it
          if (my_row >= 0) my_row += table_offset(bag * P + my_p, a.h);
          if (my_row >= 0) my_src = (uint64_t)(a.table + my_row * a.row_vecs);
        }
        float my_w = 1.f;
        if (my_use && a.slot_weight) my_w = a.slot_weight[bag * P + my_p];
        const int sub = min(G, span - p0);
        for (int cb = 0; cb < sub; cb += BATCH) {
          uint4 v[BATCH][V];
          float wt[BATCH];
          bool use[BATCH];
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            const int cc = cb + u;
            const int src = gi * G + (cc < G ? cc : 0);
            const uint4* r = reinterpret_cast<const uint4*>(__shfl_sync(0xffffffffu, my_src, src));
            wt[u] = __shfl_sync(0xffffffffu, my_w, src);
            use[u] = (cc < sub) && r != nullptr;
#pragma unroll
            for (int j = 0; j < V; ++j) {
              const int vec = j * G + lig;
              v[u][j] = make_uint4(0, 0, 0, 0);
              if (use[u] && vec < a.row_vecs) v[u][j] = ldg_nc_v4(r + vec);
            }
          }
#pragma unroll
          for (int u = 0; u < BATCH; ++u) {
            if (use[u]) {
              ++pooled;
#pragma unroll
              for (int j = 0; j < V; ++j) {
                float f[E];
                Vec16<T>::unpack(v[u][j], f);
#pragma unroll
                for (int e = 0; e < E; ++e) acc[j][e] += wt[u] * f[e];
              }
            }
          }
        }
      }
      if (a.pool_mode == RECEMB_POOL_MEAN && pooled > 0) {
        const float cntf = (float)pooled;
#pragma unroll
        for (int j = 0; j < V; ++j)
#pragma unroll
          for (int e = 0; e < E; ++e) acc[j][e] = acc[j][e] / cntf;
      }
      if (live) {
        uint4* dst = a.out + bag_out_row(bag, a.h) * (int64_t)a.row_vecs;
#pragma unroll
        for (int j = 0; j < V; ++j) {
          const int vec = j * G + lig;
          if (vec < a.row_vecs) stg_cs_v4(dst + vec, Vec16<T>::pack(acc[j]));
        }
      }
    }
    __syncthreads();
  }
}

// --------------------------------------------------------------- dispatch ----
// rows of row_vecs 16-byte vectors are moved by G lanes x V vectors
struct RowShape {
  int G, V;
};
static bool pick_shape(int row_vecs, RowShape* s) {
  int g = 1;
  while (g < 32 && g < row_vecs) g <<= 1;
  int v = (row_vecs + g - 1) / g;
  int vp = 1;
  while (vp < v) vp <<= 1;
  if (vp > 8) return false;
  s->G = g;
  s->V = vp;
  return true;
}

static int grid_for(int device, int64_t tiles, int ctas_per_sm) {
  int64_t g = (int64_t)sm_count(device) * ctas_per_sm;
  if (tiles < g) g = tiles;
  if (g < 1) g = 1;
  return (int)g;
}

// Persistent kernels stride over tiles with a fixed grid: the grid must be exactly the number
// of CTAs that are resident at once (SMs x occupancy).  A larger grid leaves a second, partial
// wave of CTAs that start late and still own as many tiles as the first wave's.
template <auto kernel, typename Args>
static void launch_persistent(const Args& a, int64_t tiles, int device, cudaStream_t s) {
  static int occ[64];  // per kernel instantiation (the kernel is a template argument), per device
  const int d = (device >= 0 && device < 64) ? device : 0;
  if (occ[d] == 0) {
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kThreads, 0) != cudaSuccess || v < 1) v = 1;
    occ[d] = v;
  }
  kernel<<<grid_for(device, tiles, occ[d]), kThreads, 0, s>>>(a);
}

#define DISPATCH_GV(G_, V_, ...)                                         \
  if (shape.G == G_ && shape.V == V_) {                                  \
    constexpr int G = G_;                                                \
    constexpr int V = V_;                                                \
    __VA_ARGS__;                                                         \
    launched = true;                                                     \
  }
#define DISPATCH_SHAPES(...)       \
  DISPATCH_GV(1, 1, __VA_ARGS__)   \
  DISPATCH_GV(2, 1, __VA_ARGS__)   \
  DISPATCH_GV(4, 1, __VA_ARGS__)   \
  DISPATCH_GV(8, 1, __VA_ARGS__)   \
  DISPATCH_GV(16, 1, __VA_ARGS__)  \
  DISPATCH_GV(32, 1, __VA_ARGS__)  \
  DISPATCH_GV(32, 2, __VA_ARGS__)  \
  DISPATCH_GV(32, 4, __VA_ARGS__)  \
  DISPATCH_GV(32, 8, __VA_ARGS__)

template <typename T>
static int launch_gather(const GatherArgs& a, RowShape shape, int epilogue, bool two, int64_t tiles,
                         int device, cudaStream_t s) {
  bool launched = false;
  const bool map = (a.h1.flip_len | a.h1.win_len) != 0;
  if (!two && epilogue == RECEMB_EPI_NONE && !map) {
    DISPATCH_SHAPES((launch_persistent<gather_kernel<G, V, T, RECEMB_EPI_NONE, false, false>>(a, tiles, device, s)))
  } else if (!two && epilogue == RECEMB_EPI_NONE) {
    DISPATCH_SHAPES((launch_persistent<gather_kernel<G, V, T, RECEMB_EPI_NONE, false>>(a, tiles, device, s)))
  } else if (!two) {
    DISPATCH_SHAPES((launch_persistent<gather_kernel<G, V, T, RECEMB_EPI_L2NORM, false>>(a, tiles, device, s)))
  } else if (epilogue == RECEMB_EPI_NONE) {
    DISPATCH_SHAPES((launch_persistent<gather_kernel<G, V, T, RECEMB_EPI_NONE, true>>(a, tiles, device, s)))
  } else {
    DISPATCH_SHAPES((launch_persistent<gather_kernel<G, V, T, RECEMB_EPI_L2NORM, true>>(a, tiles, device, s)))
  }
  if (!launched) {
    set_error("gather: no kernel for G=%d V=%d", shape.G, shape.V);
    return RECEMB_ERR_UNSUPPORTED;
  }
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

template <typename T>
static int launch_kshift(const KShiftArgs& a, RowShape shape, int64_t tiles, int device, cudaStream_t s) {
  bool launched = false;
  DISPATCH_SHAPES((launch_persistent<kshift_kernel<G, V, T>>(a, tiles, device, s)))
  if (!launched) {
    set_error("kshift: no kernel for G=%d V=%d", shape.G, shape.V);
    return RECEMB_ERR_UNSUPPORTED;
  }
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

template <typename T>
static int launch_pool(const PoolArgs& a, RowShape shape, int64_t tiles, int device, cudaStream_t s, bool peer) {
  bool launched = false;
  // rows of one 16-byte vector per lane: 4 row loads per lane and batch (no register spills at the 64-register
  // cap of 4 CTAs per SM; with the L2 prefetch one iteration ahead cfg 3 runs 0.356 instead of 0.446 ms, and
  // 0.385 without the prefetch).  RECEMB_POOL_BATCH=8 restores the 8-deep batches.
  static const int batch4 = [] {
    const char* e = getenv("RECEMB_POOL_BATCH");
    return !(e && atoi(e) == 8);
  }();
  if (peer) {
    DISPATCH_SHAPES((launch_persistent<pool_kernel<G, V, T, true>>(a, tiles, device, s)))
  } else if (batch4 && shape.V == 1) {
    DISPATCH_GV(1, 1, (launch_persistent<pool_kernel<G, V, T, false, 4>>(a, tiles, device, s)))
    DISPATCH_GV(2, 1, (launch_persistent<pool_kernel<G, V, T, false, 4>>(a, tiles, device, s)))
    DISPATCH_GV(4, 1, (launch_persistent<pool_kernel<G, V, T, false, 4>>(a, tiles, device, s)))
    DISPATCH_GV(8, 1, (launch_persistent<pool_kernel<G, V, T, false, 4>>(a, tiles, device, s)))
    DISPATCH_GV(16, 1, (launch_persistent<pool_kernel<G, V, T, false, 4>>(a, tiles, device, s)))
    DISPATCH_GV(32, 1, (launch_persistent<pool_kernel<G, V, T, false, 4>>(a, tiles, device, s)))
  } else {
    DISPATCH_SHAPES((launch_persistent<pool_kernel<G, V, T, false>>(a, tiles, device, s)))
  }
  if (!launched) {
    set_error("pool: no kernel for G=%d V=%d", shape.G, shape.V);
    return RECEMB_ERR_UNSUPPORTED;
  }
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

static int row_vecs_of(int32_t dim, int dtype, int* row_vecs) {
  RECEMB_CHECK_ARG(dtype == RECEMB_F32 || dtype == RECEMB_BF16, "dtype %d unknown", dtype);
  const int64_t bytes = (int64_t)dim * (dtype == RECEMB_F32 ? 4 : 2);
  RECEMB_UNSUPPORTED(dim > 0 && bytes % 16 == 0,
                     "row of %d elements (%lld bytes) is not a multiple of 16 bytes", dim,
                     (long long)bytes);
  *row_vecs = (int)(bytes / 16);
  return RECEMB_OK;
}

}  // namespace recemb

using namespace recemb;

extern "C" int recemb_row_index(const int64_t* ids, int64_t n, int hash_mode, int64_t num_rows,
                                int64_t hash_arg, int64_t* rows_out, int device,
                                recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n >= 0, "n < 0");
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(ids && rows_out, "null pointer");
  HashSpec h;
  int rc = make_hash_spec(hash_mode, num_rows, hash_arg, &h);
  if (rc) return rc;
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  const int grid = grid_for(device, (n + kThreads - 1) / kThreads, 8);
  row_index_kernel<<<grid, kThreads, 0, (cudaStream_t)stream>>>(ids, n, h, rows_out);
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

extern "C" int recemb_gather_fwd(const void* table, int64_t num_rows, const void* table2,
                                 int64_t num_rows2, int32_t dim, int dtype, const int64_t* ids,
                                 int64_t n, const recemb_layout* layout, int hash_mode, int hash_mode2,
                                 int64_t hash_arg, int epilogue, int zero_pad, int64_t pad_id,
                                 void* out, float* inv_norm_out, int device,
                                 recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n >= 0, "n < 0");
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(table && ids && out, "null pointer");
  RECEMB_CHECK_ARG(epilogue == RECEMB_EPI_NONE || epilogue == RECEMB_EPI_L2NORM,
                   "gather epilogue must be NONE or L2NORM");
  RECEMB_CHECK_ARG(((uintptr_t)table | (uintptr_t)out | (uintptr_t)table2) % 16 == 0,
                   "table/out must be 16-byte aligned");
  GatherArgs a;
  int rc = row_vecs_of(dim, dtype, &a.row_vecs);
  if (rc) return rc;
  RowShape shape;
  RECEMB_UNSUPPORTED(pick_shape(a.row_vecs, &shape), "dim %d too large", dim);
  const bool batched = layout && layout->ids_per_table > 0;
  RECEMB_CHECK_ARG(!batched || table2 == nullptr, "table-batched lookups take one table");
  RECEMB_UNSUPPORTED(!batched || n < 0xffffffffll, "too many lookups for table-batched mode");
  RECEMB_UNSUPPORTED(!layout || layout->shard_world <= 1, "sharded sequence gather is not implemented");
  rc = make_hash_spec(hash_mode, num_rows, hash_arg, &a.h1, layout);
  RECEMB_CHECK_ARG(!layout || layout->flip_len == 0 || n % layout->flip_len == 0,
                   "n is not a multiple of flip_len");
  if (rc) return rc;
  RECEMB_CHECK_ARG(a.h1.win_len == 0 || n % a.h1.win_len == 0, "n is not a multiple of seq_len");
  RECEMB_CHECK_ARG(a.h1.win_len == 0 || !batched || layout->ids_per_table % a.h1.win_len == 0,
                   "ids_per_table is not a multiple of seq_len");
  a.h2 = a.h1;
  if (table2) {
    rc = make_hash_spec(hash_mode2, num_rows2, hash_arg, &a.h2);
    if (rc) return rc;
  }
  a.table = (const uint4*)table;
  a.table2 = (const uint4*)table2;
  a.ids = ids;
  a.out = (uint4*)out;
  a.inv_norm = inv_norm_out;
  a.n = n;
  a.zero_pad = zero_pad;
  a.pad_id = pad_id;
  a.bulk_ok = ((uintptr_t)ids % 16 == 0);
  {
    static int pf = -1;  // RECEMB_GATHER_PREFETCH=<iters> (tuning aid); default below
    if (pf < 0) {
      const char* e = getenv("RECEMB_GATHER_PREFETCH");
      pf = e ? atoi(e) : 0;
    }
    a.prefetch_iters = pf;
  }
  {
    // rows may stay in L1 (evict_last): Zipf heads are then served by every SM's L1 instead of
    // the one L2 slice that holds them (cfg 4 fwd 2.53 -> 2.47 ms); neutral-to-better for uniform
    // ids (cfg 2 step 3.82 -> 3.77 ms).  RECEMB_GATHER_L1=0 restores L1::no_allocate.
    static int l1 = -1;
    if (l1 < 0) {
      const char* e = getenv("RECEMB_GATHER_L1");
      l1 = e ? atoi(e) : 1;
    }
    a.l1_rows = l1;
  }
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  const int64_t tiles = (n + kTileIds - 1) / kTileIds;
  if (dtype == RECEMB_F32)
    return launch_gather<float>(a, shape, epilogue, table2 != nullptr, tiles, device, (cudaStream_t)stream);
  return launch_gather<__nv_bfloat16>(a, shape, epilogue, table2 != nullptr, tiles, device,
                                      (cudaStream_t)stream);
}

extern "C" int recemb_kshift_fwd(const void* table, int64_t num_rows, int32_t dim, int dtype,
                                 const int64_t* ids, int64_t n, int32_t num_shifts, int epilogue,
                                 int32_t flip_len, void* out, float* inv_norm_out, int device,
                                 recemb_stream_t stream) {
  RECEMB_CHECK_ARG(flip_len >= 0, "flip_len < 0");
  recemb_layout layout = {0, 0, 1, 0, flip_len, 0, 0, nullptr, 0, 0, 0};
  return recemb_kshift_fwd_layout(table, num_rows, dim, dtype, ids, n, num_shifts, epilogue, &layout, out,
                                  inv_norm_out, device, stream);
}

extern "C" int recemb_kshift_fwd_layout(const void* table, int64_t num_rows, int32_t dim, int dtype,
                                        const int64_t* ids, int64_t n, int32_t num_shifts, int epilogue,
                                        const recemb_layout* layout, void* out, float* inv_norm_out, int device,
                                        recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n >= 0, "n < 0");
  RECEMB_UNSUPPORTED(!layout || (layout->ids_per_table == 0 && layout->shard_world <= 1),
                     "the k-shift bag takes one unsharded table (layout: flip_len / sequence window only)");
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(table && ids && out, "null pointer");
  RECEMB_CHECK_ARG(num_shifts >= 1 && num_shifts <= 63, "num_shifts %d out of [1, 63]", num_shifts);
  RECEMB_CHECK_ARG(num_rows >= 1, "num_rows < 1");
  RECEMB_CHECK_ARG(((uintptr_t)table | (uintptr_t)out) % 16 == 0, "table/out must be 16-byte aligned");
  KShiftArgs a;
  int rc = row_vecs_of(dim, dtype, &a.row_vecs);
  if (rc) return rc;
  RowShape shape;
  RECEMB_UNSUPPORTED(pick_shape(a.row_vecs, &shape), "dim %d too large", dim);
  a.table = (const uint4*)table;
  a.ids = ids;
  a.out = (uint4*)out;
  a.inv_norm = inv_norm_out;
  a.n = n;
  a.k = num_shifts;
  a.mod_rows = make_modn((uint64_t)num_rows);
  a.epilogue = epilogue;
  a.sqrt_k = (float)sqrt((double)num_shifts);
  rc = make_hash_spec(RECEMB_HASH_FLOORMOD, num_rows, 0, &a.h, layout);
  if (rc) return rc;
  RECEMB_CHECK_ARG(a.h.flip_len == 0 || n % a.h.flip_len == 0, "n is not a multiple of flip_len");
  RECEMB_CHECK_ARG(a.h.win_len == 0 || n % a.h.win_len == 0, "n is not a multiple of seq_len");
  a.bulk_ok = ((uintptr_t)ids % 16 == 0);
  {
    static int l1 = -1;  // RECEMB_KSHIFT_L1=0 turns the L1 residency of the collapse rows off
    if (l1 < 0) {
      const char* e = getenv("RECEMB_KSHIFT_L1");
      l1 = e ? atoi(e) : 1;
    }
    a.l1_hot = l1;
  }
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  const int64_t tiles = (n + kTileIds - 1) / kTileIds;
  if (dtype == RECEMB_F32) return launch_kshift<float>(a, shape, tiles, device, (cudaStream_t)stream);
  return launch_kshift<__nv_bfloat16>(a, shape, tiles, device, (cudaStream_t)stream);
}

// shared by recemb_pool_fwd (one local table / shard) and recemb_peer_pool_fwd (group != nullptr:
// rows are read from the owning rank's shard over peer-mapped memory)
static int pool_fwd_common(const void* table, const recemb_peer_group* group, int64_t num_rows, int32_t dim,
                           int dtype, const int64_t* ids, int64_t num_bags, int32_t bag_size,
                           const int32_t* lengths, int32_t last_n, const float* per_slot_weight,
                           int hash_mode, int64_t hash_arg, int pool_mode, int zero_pad,
                           int64_t pad_id, const recemb_layout* layout, void* out, int device,
                           recemb_stream_t stream) {
  RECEMB_CHECK_ARG(num_bags >= 0, "num_bags < 0");
  if (num_bags == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG((table || group) && ids && out, "null pointer");
  RECEMB_CHECK_ARG(bag_size >= 1, "bag_size < 1");
  RECEMB_UNSUPPORTED(bag_size <= kPoolTileIds, "bag_size %d > %d", bag_size, kPoolTileIds);
  RECEMB_CHECK_ARG(pool_mode == RECEMB_POOL_SUM || pool_mode == RECEMB_POOL_MEAN, "bad pool_mode");
  RECEMB_CHECK_ARG(((uintptr_t)table | (uintptr_t)out) % 16 == 0, "table/out must be 16-byte aligned");
  PoolArgs a;
  int rc = row_vecs_of(dim, dtype, &a.row_vecs);
  if (rc) return rc;
  RowShape shape;
  RECEMB_UNSUPPORTED(pick_shape(a.row_vecs, &shape), "dim %d too large", dim);
  rc = make_hash_spec(hash_mode, num_rows, hash_arg, &a.h, layout);
  if (rc) return rc;
  RECEMB_UNSUPPORTED(!(layout && layout->ids_per_table > 0) || num_bags * bag_size < 0xffffffffll,
                     "too many slots for table-batched mode");
  if (a.h.out_feats) {
    RECEMB_CHECK_ARG(layout->ids_per_table % bag_size == 0, "ids_per_table is not a multiple of bag_size");
    a.h.out_bpt = (uint32_t)(layout->ids_per_table / bag_size);
    const int64_t tabs = (num_bags + a.h.out_bpt - 1) / a.h.out_bpt;
    RECEMB_CHECK_ARG(tabs + a.h.out_feat_off <= a.h.out_feats, "out_features %u < %lld tables + offset %u",
                     a.h.out_feats, (long long)tabs, a.h.out_feat_off);
  }
  a.table = (const uint4*)table;
  a.rows_div = num_rows;
  a.rows_rem = 0;
  for (int i = 0; i < RECEMB_MAX_PEERS; ++i) a.peer_table[i] = nullptr;
  if (group) {
    RECEMB_CHECK_ARG(group->world >= 1 && group->world <= RECEMB_MAX_PEERS && group->rank >= 0 &&
                         group->rank < group->world,
                     "peer group world %d / rank %d out of range", group->world, group->rank);
    RECEMB_CHECK_ARG(!layout || layout->shard_world <= 1, "peer pooling takes an unsharded layout");
    for (int i = 0; i < group->world; ++i) {
      RECEMB_CHECK_ARG(group->table[i] && (uintptr_t)group->table[i] % 16 == 0, "peer table %d null / misaligned", i);
      a.peer_table[i] = (const uint4*)group->table[i];
    }
    a.h.mod_world = make_modn((uint64_t)group->world);
    a.rows_div = num_rows / group->world;
    a.rows_rem = (uint32_t)(num_rows % group->world);
  }
  a.ids = ids;
  a.lengths = lengths;
  a.slot_weight = per_slot_weight;
  a.out = (uint4*)out;
  a.num_bags = num_bags;
  a.bag_size = bag_size;
  a.last_n = last_n;
  a.bags_per_tile = kPoolTileIds / bag_size;
  {
    // small problems: shrink the tile until there are >= 4 tiles per resident CTA, otherwise a
    // handful of CTAs get a second tile and the launch takes twice as long (tail effect)
    const int64_t want_tiles = (int64_t)sm_count(device) * 4 * 4;
    int64_t bpt = (num_bags + want_tiles - 1) / want_tiles;
    if (bpt < 8) bpt = 8;
    if ((bpt * bag_size) % 2) ++bpt;  // even id count per tile keeps every tile 16-byte aligned
    if (bpt < a.bags_per_tile) a.bags_per_tile = (int32_t)bpt;
  }
  a.pool_mode = pool_mode;
  a.zero_pad = zero_pad;
  a.pad_id = pad_id;
  {
    // RECEMB_POOL_PREFETCH = iterations ahead (0 = off; cfg 3: 0.385 / 0.358 / 0.373 / 0.383 ms for 0 / 1 / 2 / 3)
    static const int pf = [] {
      const char* e = getenv("RECEMB_POOL_PREFETCH");
      return e ? atoi(e) : 1;
    }();
    a.prefetch = pf;
  }
  // every tile start must be 16-byte aligned and every tile an even id count
  a.bulk_ok = ((uintptr_t)ids % 16 == 0) && (((int64_t)a.bags_per_tile * bag_size) % 2 == 0);
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  const int64_t tiles = (num_bags + a.bags_per_tile - 1) / a.bags_per_tile;
  if (dtype == RECEMB_F32) return launch_pool<float>(a, shape, tiles, device, (cudaStream_t)stream, group != nullptr);
  return launch_pool<__nv_bfloat16>(a, shape, tiles, device, (cudaStream_t)stream, group != nullptr);
}

extern "C" int recemb_pool_fwd(const void* table, int64_t num_rows, int32_t dim, int dtype,
                               const int64_t* ids, int64_t num_bags, int32_t bag_size,
                               const int32_t* lengths, int32_t last_n, const float* per_slot_weight,
                               int hash_mode, int64_t hash_arg, int pool_mode, int zero_pad,
                               int64_t pad_id, const recemb_layout* layout, void* out, int device,
                               recemb_stream_t stream) {
  RECEMB_CHECK_ARG(table || num_bags == 0, "null table");
  return pool_fwd_common(table, nullptr, num_rows, dim, dtype, ids, num_bags, bag_size, lengths, last_n,
                         per_slot_weight, hash_mode, hash_arg, pool_mode, zero_pad, pad_id, layout, out, device,
                         stream);
}

extern "C" int recemb_peer_pool_fwd(const recemb_peer_group* group, int64_t num_rows, int32_t dim, int dtype,
                                    const int64_t* ids, int64_t num_bags, int32_t bag_size,
                                    const int32_t* lengths, int32_t last_n, const float* per_slot_weight,
                                    int hash_mode, int64_t hash_arg, int pool_mode, int zero_pad, int64_t pad_id,
                                    const recemb_layout* layout, void* out, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(group, "null peer group");
  return pool_fwd_common(nullptr, group, num_rows, dim, dtype, ids, num_bags, bag_size, lengths, last_n,
                         per_slot_weight, hash_mode, hash_arg, pool_mode, zero_pad, pad_id, layout, out, device,
                         stream);
}
