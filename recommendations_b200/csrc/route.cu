// Routed exchange for row-wise sharded pooled lookups (north_star item 4, cfg 5).
//
// Every rank buckets ITS lookup slots by owner (stable counting sort, W bins), the buckets go to
// their owners with one variable-split all-to-all, and an owner then only ever touches the
// lookups it owns: its pooling, its sort and its segmented reduction are proportional to the
// slots it serves (~ n per rank whatever W is), not to W x n as with an id all-gather.
//
//   recemb_shard_bucket      sender:  ids -> entries (local row, global bag) grouped by owner + counts
//   recemb_pool_entries      owner:   runs of equal bag in the received entries -> partial pools
//   recemb_bwd_plan_entries  owner:   received entries -> sorted (row, gradient row) plan
//
// An entry is one int64: (local row in the owner's stacked shard) << 32 | (sender_rank * bags + bag).
// Inside a bucket the entries keep slot order, i.e. they are sorted by bag: the owner sees each
// (sender, bag) as one contiguous run and pools it in slot order (same fp32 order as the
// unsharded kernel).

#include "common.cuh"
#include "sort.cuh"

namespace recemb {

constexpr int kRtThreads = 256;
constexpr int kRtItems = 4;                         // slots per thread
constexpr int kRtBlock = kRtThreads * kRtItems;     // slots per CTA
constexpr int kMaxWorld = 32;

struct BucketArgs {
  const int64_t* ids;
  int64_t n;
  HashSpec h;           // mod_world / shard_world set; ids_per_table / num_tables for the table index
  int64_t num_rows;     // global rows of one table
  int zero_pad;
  int64_t pad_id;
  int32_t bag_size;
  const int32_t* lengths;
  int32_t last_n;
  uint32_t bag_base;    // sender_rank * bags_total
  uint8_t* owner8;      // [n]  owner of the slot, 0xff = dropped
  uint32_t* rowbuf;     // [n]  local row (incl. table offset) at the owner
  uint32_t* block_hist; // [num_ctas, world]
  int64_t* entries;     // [n]
  int64_t* counts;      // [world]
  uint32_t* block_base; // [num_ctas, world]  exclusive offsets, filled by the scan
  // peer exchange (PEER kernels): entries go straight into the owners' inboxes over NVLink
  int64_t* peer_inbox[RECEMB_MAX_PEERS];  // owner o's inbox region reserved for THIS rank, [cap]
  int64_t* peer_count[RECEMB_MAX_PEERS];  // owner o's count slot for THIS rank
  uint32_t cap;                           // inbox capacity per sender
  uint32_t* status;                       // this rank's sticky status word (bit 0: inbox overflow)
  uint32_t* peer_status[RECEMB_MAX_PEERS];  // owner o's status word: its inbox is incomplete -> it must skip its update
  // sequence mode (every lookup is its own gradient row): the entry names the slot of the owner's gradient
  // buffer the sender will store the row into, rank * cap + position in the bucket; dest[s] remembers
  // (owner << 32 | position) for that store, -1 for a dropped lookup
  int64_t* dest;
  uint32_t seq_base;  // rank * cap
};

__global__ void __launch_bounds__(kRtThreads) bucket_count_kernel(const BucketArgs a) {
  __shared__ uint32_t s_hist[kMaxWorld];
  if (threadIdx.x < kMaxWorld) s_hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t world = a.h.shard_world;
  const int64_t base = (int64_t)blockIdx.x * kRtBlock;
  for (int r = 0; r < kRtItems; ++r) {
    const int64_t s = base + r * kRtThreads + threadIdx.x;
    if (s >= a.n) break;
    const int64_t id = a.ids[s];
    bool ok = !(a.zero_pad && id == a.pad_id);
    const int64_t bag = s / a.bag_size;
    if (ok) {
      const int p = (int)(s - bag * a.bag_size);
      int hi = a.bag_size;
      if (a.lengths) hi = min(max(a.lengths[bag], 0), a.bag_size);
      const int lo = a.last_n > 0 ? max(0, hi - a.last_n) : 0;
      ok = p >= lo && p < hi;
    }
    uint8_t owner = 0xff;
    uint32_t lrow = 0;
    const int64_t srow = ok ? row_of(id, a.h) : -1;  // -1 also for an out-of-range identity id
    if (srow >= 0) {
      const uint64_t row = (uint64_t)srow;
      uint32_t t = 0;
      if (a.h.ids_per_table) {
        t = (uint32_t)s / a.h.ids_per_table;
        if (a.h.num_tables) t %= a.h.num_tables;
      }
      uint32_t o;
      if (a.h.partition) {
        // table-wise: table t lives whole on rank t % world as its local table t / world
        o = t % world;
        lrow = (uint32_t)(row + (uint64_t)(t / world) * (uint64_t)a.num_rows);
      } else {
        uint64_t q;
        o = (uint32_t)udivmod(row, a.h.mod_world, &q);
        // stacked shard of owner o: table t starts at t * local_rows(o)
        const uint64_t local_rows = ((uint64_t)a.num_rows - o + world - 1) / world;
        lrow = (uint32_t)(q + (uint64_t)t * local_rows);
      }
      owner = (uint8_t)o;
      atomicAdd(&s_hist[o], 1u);
    }
    a.owner8[s] = owner;
    a.rowbuf[s] = lrow;
  }
  __syncthreads();
  if (threadIdx.x < world) a.block_hist[(int64_t)blockIdx.x * world + threadIdx.x] = s_hist[threadIdx.x];
}

// one warp per owner: exclusive scan of that owner's per-CTA counts; then owner bases
template <bool PEER>
__global__ void __launch_bounds__(1024) bucket_scan_kernel(const BucketArgs a, int num_ctas) {
  __shared__ uint32_t s_total[kMaxWorld];
  const int world = (int)a.h.shard_world;
  const int o = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (o < world) {
    // each lane owns kScanItems consecutive CTAs per round: that many loads in flight per lane
    // instead of one dependent global load per 32 CTAs
    constexpr int kScanItems = 8;
    uint32_t carry = 0;
    for (int c0 = 0; c0 < num_ctas; c0 += 32 * kScanItems) {
      uint32_t v[kScanItems];
      uint32_t mine = 0;
#pragma unroll
      for (int j = 0; j < kScanItems; ++j) {
        const int c = c0 + lane * kScanItems + j;
        v[j] = c < num_ctas ? a.block_hist[(int64_t)c * world + o] : 0u;
      }
#pragma unroll
      for (int j = 0; j < kScanItems; ++j) mine += v[j];
      uint32_t inc = mine;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      uint32_t run = carry + inc - mine;
#pragma unroll
      for (int j = 0; j < kScanItems; ++j) {
        const int c = c0 + lane * kScanItems + j;
        if (c < num_ctas) a.block_base[(int64_t)c * world + o] = run;
        run += v[j];
      }
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) s_total[o] = carry;
  }
  __syncthreads();
  if constexpr (PEER) {
    // every bucket starts at 0 of its own inbox region; the owner learns the size from its count slot
    if (threadIdx.x < world) {
      uint32_t t = s_total[threadIdx.x];
      if (t > a.cap) {
        t = a.cap;
        atomicOr(a.status, 1u);
        // the owner would apply an incomplete gradient: flag it as well (system scope, peer memory) so
        // that its guarded update leaves the shard untouched in THIS step
        atomicOr_system(a.peer_status[threadIdx.x], 1u);
      }
      *a.peer_count[threadIdx.x] = (int64_t)t;
      if (a.counts) a.counts[threadIdx.x] = (int64_t)t;
    }
    return;
  }
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int i = 0; i < world; ++i) {
      a.counts[i] = (int64_t)s_total[i];
      const uint32_t t = s_total[i];
      s_total[i] = run;  // base offset of owner i's bucket
      run += t;
    }
  }
  __syncthreads();
  // fold the owner bases into the per-CTA offsets
  for (int i = threadIdx.x; i < num_ctas * world; i += blockDim.x) a.block_base[i] += s_total[i % world];
}

// stable scatter: position = base[cta][owner] + (valid slots of that owner earlier in the CTA)
template <bool PEER>
__global__ void __launch_bounds__(kRtThreads) bucket_scatter_kernel(const BucketArgs a) {
  __shared__ uint32_t s_run[kMaxWorld];                       // running offset per owner
  __shared__ uint32_t s_wc[kRtThreads / 32][kMaxWorld];       // per-warp counts of this round
  __shared__ int64_t* s_dst[PEER ? RECEMB_MAX_PEERS : 1];     // PEER: my region of every owner's inbox
  // PEER: the CTA's entries are first grouped by owner in shared memory and then copied out by consecutive
  // threads -- contiguous runs of ~kRtBlock / world entries per owner leave as full-width NVLink writes instead
  // of the 8-byte stores of the lanes that happen to share an owner (a handful of bytes per packet)
  __shared__ int64_t s_stage[PEER ? kRtBlock : 1];
  __shared__ uint32_t s_base0[PEER ? kMaxWorld : 1];          // the CTA's first position in owner o's bucket
  __shared__ uint32_t s_lbase[PEER ? kMaxWorld + 1 : 1];      // first staging slot of owner o
  const uint32_t world = a.h.shard_world;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < world) s_run[threadIdx.x] = a.block_base[(int64_t)blockIdx.x * world + threadIdx.x];
  if constexpr (PEER) {
    if (threadIdx.x < RECEMB_MAX_PEERS) s_dst[threadIdx.x] = a.peer_inbox[threadIdx.x];
    if (threadIdx.x < world) s_base0[threadIdx.x] = a.block_base[(int64_t)blockIdx.x * world + threadIdx.x];
    if (threadIdx.x == 0) {
      uint32_t run = 0;
      for (uint32_t o = 0; o < world; ++o) {
        s_lbase[o] = run;
        run += a.block_hist[(int64_t)blockIdx.x * world + o];
      }
      s_lbase[world] = run;
    }
  }  // published by the first __syncthreads() of the loop
  const int64_t base = (int64_t)blockIdx.x * kRtBlock;
  for (int r = 0; r < kRtItems; ++r) {
    for (int i = threadIdx.x; i < (kRtThreads / 32) * kMaxWorld; i += kRtThreads) (&s_wc[0][0])[i] = 0;
    __syncthreads();
    const int64_t s = base + r * kRtThreads + threadIdx.x;
    const uint32_t owner = s < a.n ? a.owner8[s] : 0xffu;
    const bool valid = owner != 0xffu;
    // rank among the lanes of this warp with the same owner
    const uint32_t peers = __match_any_sync(0xffffffffu, owner);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) s_wc[warp][owner] = __popc(peers);
    __syncthreads();
    int64_t dest = -1;
    if (valid) {
      uint32_t pos = s_run[owner] + rank;
      for (int w = 0; w < warp; ++w) pos += s_wc[w][owner];
      const uint32_t bag = (uint32_t)(s / a.bag_size);
      const uint32_t grow = a.dest ? a.seq_base + pos : a.bag_base + bag;  // gradient row at the owner
      const int64_t entry = (int64_t)(((uint64_t)a.rowbuf[s] << 32) | (uint64_t)grow);
      if constexpr (PEER) {
        s_stage[s_lbase[owner] + (pos - s_base0[owner])] = entry;
        if (pos < a.cap) dest = (int64_t)(((uint64_t)owner << 32) | pos);
      } else {
        a.entries[pos] = entry;
      }
    }
    if (PEER && a.dest && s < a.n) a.dest[s] = dest;
    __syncthreads();
    if (threadIdx.x < world) {
      uint32_t add = 0;
      for (int w = 0; w < kRtThreads / 32; ++w) add += s_wc[w][threadIdx.x];
      s_run[threadIdx.x] += add;
    }
    __syncthreads();
  }
  if constexpr (PEER) {
    const uint32_t total = s_lbase[world];
    for (uint32_t i = threadIdx.x; i < total; i += kRtThreads) {
      uint32_t o = 0;
      while (i >= s_lbase[o + 1]) ++o;
      const uint32_t pos = s_base0[o] + (i - s_lbase[o]);
      if (pos < a.cap) s_dst[o][pos] = s_stage[i];  // into the owner's HBM over NVLink, consecutive threads ->
    }                                                // consecutive inbox positions
  }
}

// ------------------------------------------------------------ owner-side pooling ----
// Group of G lanes per chunk of kPeChunk entries; a run (equal bag key) belongs to the chunk in
// which it starts: the group skips a leading run that started earlier and walks past its chunk
// end to finish its last run.  fp32 accumulation in entry (= slot) order; each output row is
// written exactly once (rows of bags without owned slots stay at the caller's zero fill).
constexpr int kPeChunk = 32;

struct PoolEntriesArgs {
  const int64_t* entries;
  int32_t n;
  const void* table;
  uint32_t dim;
  int32_t vecs;  // 16-byte vectors per row
  void* out;
};

template <int G, typename T>
__global__ void __launch_bounds__(kRtThreads, 4) pool_entries_kernel(const PoolEntriesArgs a) {
  constexpr int E = Vec16<T>::kElems;
  constexpr int B = 4;
  const int lane = threadIdx.x & 31, lig = lane % G;
  const int chunk = blockIdx.x * (kRtThreads / G) + threadIdx.x / G;
  const int start = chunk * kPeChunk;
  if (start >= a.n) return;
  const int end = min(start + kPeChunk, a.n);
  const uint4* table = reinterpret_cast<const uint4*>(a.table);
  uint4* out = reinterpret_cast<uint4*>(a.out);
  auto key_of = [&](int i) { return (uint32_t)((uint64_t)a.entries[i] & 0xffffffffull); };

  int i = start;
  if (start > 0) {  // skip the tail of a run that started in an earlier chunk
    const uint32_t prev = key_of(start - 1);
    while (i < a.n && i < end && key_of(i) == prev) ++i;
    if (i == end && i < a.n && key_of(i) == prev) return;  // whole chunk inside an older run
  }
  float acc[E];
  while (i < end) {  // runs starting in [start, end)
    const uint32_t key = key_of(i);
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    bool more = true;
    while (more) {
      uint4 v[B];
      bool use[B];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int j = i + u;
        use[u] = false;
        v[u] = make_uint4(0, 0, 0, 0);
        if (j < a.n) {
          const uint64_t ent = (uint64_t)a.entries[j];
          if ((uint32_t)(ent & 0xffffffffull) == key) {
            use[u] = true;
            if (lig < a.vecs) v[u] = ldg_nc_v4(table + (size_t)(ent >> 32) * a.vecs + lig);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < B; ++u) {
        if (use[u] && more) {
          float f[E];
          Vec16<T>::unpack(v[u], f);
#pragma unroll
          for (int e = 0; e < E; ++e) acc[e] += f[e];
          ++i;
        } else {
          more = false;
        }
      }
    }
    if (lig < a.vecs) stg_cs_v4(out + (size_t)key * a.vecs + lig, Vec16<T>::pack(acc));
  }
}

// Owner side of the push forward: the same run pooling over MY inbox ([world][cap] regions with
// per-sender counts), each partial row stored into the requesting rank's parts[rank][bag] over
// NVLink (256-byte contiguous stores: posted writes, no round trip).
struct PoolInboxArgs {
  const int64_t* inbox;
  const int64_t* counts;
  int64_t cap;
  int64_t bags_total;
  int32_t world;
  int32_t rank;
  int32_t chunks_per_sender;
  const void* table;
  int32_t vecs;
  uint4* parts[RECEMB_MAX_PEERS];  // requester s: its parts region, slice of THIS owner
  int64_t bags_per_table;          // > 0: table-wise partitioning, only bags of tables t % world == rank are mine
};

// zero rows for the bags [k0, k1) of a sender's region that have no entry on this owner (table-wise
// partitioning: only the bags of the tables this rank owns, skipping foreign tables a whole table at a time)
__device__ __noinline__ void zero_gap_rows(uint4* out, uint32_t key_base, uint32_t k0, uint32_t k1, int vecs, int lig,
                                           int64_t bags_per_table, int world, int rank) {
  if (lig >= vecs) return;
  if (bags_per_table > 0) {
    uint32_t z = k0;
    while (z != k1) {
      const uint32_t t = (uint32_t)((z - key_base) / (uint32_t)bags_per_table);
      const uint32_t t_end = key_base + (t + 1u) * (uint32_t)bags_per_table;
      const uint32_t stop = (t_end - z) < (k1 - z) ? t_end : k1;
      if ((int)(t % (uint32_t)world) == rank)
        for (uint32_t y = z; y != stop; ++y) stg_v4(out + (size_t)(y - key_base) * vecs + lig, make_uint4(0, 0, 0, 0));
      z = stop;
    }
    return;
  }
  for (uint32_t z = k0; z != k1; ++z) stg_v4(out + (size_t)(z - key_base) * vecs + lig, make_uint4(0, 0, 0, 0));
}

template <int G, typename T>
__global__ void __launch_bounds__(kRtThreads, 4) pool_inbox_push_kernel(const PoolInboxArgs a) {
  constexpr int E = Vec16<T>::kElems;
  constexpr int B = 4;
  __shared__ uint4* s_parts[RECEMB_MAX_PEERS];
  if (threadIdx.x < RECEMB_MAX_PEERS) s_parts[threadIdx.x] = a.parts[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31, lig = lane % G;
  const int64_t chunk_global = (int64_t)blockIdx.x * (kRtThreads / G) + threadIdx.x / G;
  const int sender_slot = (int)(chunk_global / a.chunks_per_sender);
  if (sender_slot >= a.world) return;
  // rotated: the first CTAs of rank r serve sender r + 1, so the ranks do not all store their partial
  // rows into peer 0 first (NVSwitch gives every pair full bandwidth, one ingress port does not)
  int sender = sender_slot + a.rank + 1;
  while (sender >= a.world) sender -= a.world;
  const int n = (int)a.counts[sender];
  const int chunk_local = (int)(chunk_global - (int64_t)sender_slot * a.chunks_per_sender);
  const int start = chunk_local * kPeChunk;
  uint4* out = s_parts[sender];
  const uint32_t key_base = (uint32_t)((int64_t)sender * a.bags_total);
  // Bags of this sender without an entry on this owner get a zero row from here as well (every row of
  // parts[rank] is written exactly once per step, the requester never zero-fills): the run that follows
  // a gap of bag numbers fills it, the last run of the region fills the tail.
  // gaps are rare (a bag without an entry on this owner): the filler stays out of line so that the walk below
  // keeps its registers and its instruction footprint
  auto zero_rows = [&](uint32_t k0, uint32_t k1) {  // keys [k0, k1)
    if (k0 != k1) zero_gap_rows(out, key_base, k0, k1, a.vecs, lig, a.bags_per_table, a.world, a.rank);
  };
  if (n == 0) {
    if (chunk_local == 0) zero_rows(key_base, key_base + (uint32_t)a.bags_total);
    return;
  }
  if (start >= n) return;
  const int end = min(start + kPeChunk, n);
  const int64_t* entries = a.inbox + (int64_t)sender * a.cap;
  const uint4* table = reinterpret_cast<const uint4*>(a.table);
  auto key_of = [&](int i) { return (uint32_t)((uint64_t)entries[i] & 0xffffffffull); };

  int i = start;
  uint32_t prev_key = key_base - 1u;  // key of the run before the next one that starts in this chunk
  if (start > 0) {  // skip the tail of a run that started in an earlier chunk
    prev_key = key_of(start - 1);
    while (i < n && i < end && key_of(i) == prev_key) ++i;
    if (i == end) return;  // whole chunk inside an older run: its owner walks on (and fills the tail)
  }
  float acc[E];
  while (i < end) {  // runs starting in [start, end)
    const uint32_t key = key_of(i);
    zero_rows(prev_key + 1u, key);
    prev_key = key;
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e] = 0.f;
    bool more = true;
    while (more) {
      uint4 v[B];
      bool use[B];
#pragma unroll
      for (int u = 0; u < B; ++u) {
        const int j = i + u;
        use[u] = false;
        v[u] = make_uint4(0, 0, 0, 0);
        if (j < n) {
          const uint64_t ent = (uint64_t)entries[j];
          if ((uint32_t)(ent & 0xffffffffull) == key) {
            use[u] = true;
            if (lig < a.vecs) v[u] = ldg_nc_v4(table + (size_t)(ent >> 32) * a.vecs + lig);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < B; ++u) {
        if (use[u] && more) {
          float f[E];
          Vec16<T>::unpack(v[u], f);
#pragma unroll
          for (int e = 0; e < E; ++e) acc[e] += f[e];
          ++i;
        } else {
          more = false;
        }
      }
    }
    if (lig < a.vecs) stg_v4(out + (size_t)(key - key_base) * a.vecs + lig, Vec16<T>::pack(acc));
  }
  if (i >= n) zero_rows(prev_key + 1u, key_base + (uint32_t)a.bags_total);  // I closed the region's last run
}

// Sequence mode, backward: every gradient row travels ONCE, to the rank that owns its table row -- stored
// at (my rank * cap + position in my bucket for that owner) of the owner's gradient buffer, the slot the
// entry pushed during the forward names.  G lanes per row, 16 bytes each: contiguous row_bytes stores.
struct RowsPushArgs {
  const uint4* rows;
  const int64_t* dest;
  int64_t n;
  int32_t vecs;
  uint32_t seq_base;
  uint4* grads[RECEMB_MAX_PEERS];
};

template <int G>
__global__ void __launch_bounds__(kRtThreads) rows_scatter_push_kernel(const RowsPushArgs a) {
  __shared__ uint4* s_grads[RECEMB_MAX_PEERS];
  if (threadIdx.x < RECEMB_MAX_PEERS) s_grads[threadIdx.x] = a.grads[threadIdx.x];
  __syncthreads();
  const int lig = threadIdx.x % G;
  int64_t s = (int64_t)blockIdx.x * (kRtThreads / G) + threadIdx.x / G;
  const int64_t stride = (int64_t)gridDim.x * (kRtThreads / G);
  for (; s < a.n; s += stride) {
    const int64_t d = a.dest[s];
    if (d < 0 || lig >= a.vecs) continue;
    const uint4 v = ldg_nc_v4(a.rows + s * a.vecs + lig);
    stg_v4(s_grads[(uint32_t)(d >> 32)] + (size_t)(a.seq_base + (uint32_t)d) * a.vecs + lig, v);
  }
}

// ------------------------------------------------------------------ plan from entries ----
__global__ void __launch_bounds__(kRtThreads) unpack_entries_kernel(const int64_t* __restrict__ entries, int64_t n,
                                                                   uint32_t* __restrict__ keys,
                                                                   uint32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * kRtThreads + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * kRtThreads;
  for (; i < n; i += stride) {
    const uint64_t e = (uint64_t)entries[i];
    keys[i] = (uint32_t)(e >> 32);
    vals[i] = (uint32_t)(e & 0xffffffffull);
  }
}

// inbox [world][cap] + counts [world] -> (key, value) pairs; unused positions get the sentinel key
__global__ void __launch_bounds__(kRtThreads) unpack_inbox_kernel(const int64_t* __restrict__ inbox,
                                                                 const int64_t* __restrict__ counts, int64_t cap,
                                                                 int64_t n, uint32_t sentinel,
                                                                 uint32_t* __restrict__ keys,
                                                                 uint32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * kRtThreads + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * kRtThreads;
  for (; i < n; i += stride) {
    const int64_t sender = i / cap;
    const bool valid = (i - sender * cap) < counts[sender];
    const uint64_t e = valid ? (uint64_t)inbox[i] : 0ull;
    keys[i] = valid ? (uint32_t)(e >> 32) : sentinel;
    vals[i] = (uint32_t)(e & 0xffffffffull);
  }
}

}  // namespace recemb

using namespace recemb;

extern "C" size_t recemb_shard_bucket_workspace_bytes(int64_t n_slots, int32_t world) {
  if (n_slots < 0 || world < 1) return 0;
  const int64_t ctas = (n_slots + kRtBlock - 1) / kRtBlock;
  return align_up((size_t)n_slots, 256) + align_up((size_t)n_slots * 4, 256) +
         2 * align_up((size_t)(ctas > 0 ? ctas : 1) * world * 4, 256) + 256;
}

// group == nullptr: buckets into entries_out / counts_out (routed NCCL exchange);
// group != nullptr: buckets straight into the owners' inboxes (peer exchange)
static int bucket_common(const recemb_peer_group* group, const recemb_peer_arena* arena, const int64_t* ids,
                         int64_t n_ids, const recemb_layout* layout, int hash_mode, int64_t num_rows,
                         int64_t hash_arg, int zero_pad, int64_t pad_id, int32_t bag_size, const int32_t* lengths,
                         int32_t last_n, int64_t bags_total, int64_t* entries_out, int64_t* counts_out,
                         void* workspace, size_t workspace_bytes, int device, recemb_stream_t stream,
                         int64_t* seq_dest = nullptr) {
  RECEMB_CHECK_ARG(layout && layout->shard_world >= 1 && layout->shard_world <= kMaxWorld,
                   "shard_world outside [1, %d]", kMaxWorld);
  RECEMB_CHECK_ARG(bag_size >= 1 && n_ids >= 0 && n_ids % bag_size == 0, "n_ids not a multiple of bag_size");
  RECEMB_CHECK_ARG((group || counts_out) && workspace, "null pointer");
  RECEMB_UNSUPPORTED(n_ids < 0x7fffffffll, "too many slots");
  RECEMB_UNSUPPORTED((int64_t)layout->shard_world * bags_total < 0xffffffffll, "bag keys overflow 32 bits");
  RECEMB_UNSUPPORTED(recemb_layout_total_rows(num_rows, layout, n_ids) < 0xfffffff0ll, "local rows overflow 32 bits");
  const size_t need = recemb_shard_bucket_workspace_bytes(n_ids, layout->shard_world);
  if (workspace_bytes < need) {
    set_error("workspace %zu < required %zu", workspace_bytes, need);
    return RECEMB_ERR_WORKSPACE;
  }
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  cudaStream_t s = (cudaStream_t)stream;
  BucketArgs a;
  int rc = make_hash_spec(hash_mode, num_rows, hash_arg, &a.h, layout);
  if (rc) return rc;
  a.h.shard_world = (uint32_t)layout->shard_world;  // also for world == 1 (one bucket)
  a.h.mod_world = make_modn((uint64_t)layout->shard_world);
  a.ids = ids;
  a.n = n_ids;
  a.num_rows = num_rows;
  a.zero_pad = zero_pad;
  a.pad_id = pad_id;
  a.bag_size = bag_size;
  a.lengths = lengths;
  a.last_n = last_n;
  a.bag_base = (uint32_t)(layout->shard_rank * bags_total);
  const int64_t ctas = (n_ids + kRtBlock - 1) / kRtBlock;
  char* w = (char*)workspace;
  a.owner8 = (uint8_t*)w;
  w += align_up((size_t)n_ids, 256);
  a.rowbuf = (uint32_t*)w;
  w += align_up((size_t)n_ids * 4, 256);
  a.block_hist = (uint32_t*)w;
  w += align_up((size_t)(ctas > 0 ? ctas : 1) * layout->shard_world * 4, 256);
  a.block_base = (uint32_t*)w;
  a.entries = entries_out;
  a.counts = counts_out;
  a.cap = 0;
  a.status = nullptr;
  a.dest = seq_dest;
  a.seq_base = 0;
  for (int i = 0; i < RECEMB_MAX_PEERS; ++i) {
    a.peer_inbox[i] = a.peer_count[i] = nullptr;
    a.peer_status[i] = nullptr;
  }
  if (group) {
    RECEMB_CHECK_ARG(arena && group->world == layout->shard_world && group->rank == layout->shard_rank &&
                         group->world <= RECEMB_MAX_PEERS,
                     "peer group does not match the layout (world %d rank %d)", group->world, group->rank);
    RECEMB_CHECK_ARG(arena->cap >= 1 && arena->cap < 0xffffffffll && (seq_dest || arena->bags_total == bags_total),
                     "arena capacity / bags_total mismatch");
    for (int o = 0; o < group->world; ++o) {
      RECEMB_CHECK_ARG(group->arena[o], "peer arena %d not mapped", o);
      char* base = (char*)group->arena[o];
      a.peer_inbox[o] = (int64_t*)(base + arena->off_inbox) + (int64_t)group->rank * arena->cap;
      a.peer_count[o] = (int64_t*)(base + arena->off_counts) + group->rank;
      a.peer_status[o] = (uint32_t*)(base + arena->off_status);
    }
    a.cap = (uint32_t)arena->cap;
    a.status = (uint32_t*)((char*)group->arena[group->rank] + arena->off_status);
    RECEMB_UNSUPPORTED(!seq_dest || (int64_t)group->world * arena->cap < 0xffffffffll, "gradient buffer too large");
    a.seq_base = (uint32_t)((int64_t)group->rank * arena->cap);
  }
  if (n_ids == 0) {
    if (group) {
      bucket_scan_kernel<true><<<1, 32 * layout->shard_world, 0, s>>>(a, 0);  // zero counts at the owners
      RECEMB_LAUNCHED();
    } else {
      RECEMB_CUDA(cudaMemsetAsync(counts_out, 0, sizeof(int64_t) * layout->shard_world, s));
    }
    return RECEMB_OK;
  }
  RECEMB_CHECK_ARG(ids && (group || entries_out), "null ids / entries");
  bucket_count_kernel<<<(unsigned)ctas, kRtThreads, 0, s>>>(a);
  RECEMB_LAUNCHED();
  if (group) {
    bucket_scan_kernel<true><<<1, 32 * layout->shard_world, 0, s>>>(a, (int)ctas);
    RECEMB_LAUNCHED();
    bucket_scatter_kernel<true><<<(unsigned)ctas, kRtThreads, 0, s>>>(a);
    RECEMB_LAUNCHED();
  } else {
    bucket_scan_kernel<false><<<1, 32 * layout->shard_world, 0, s>>>(a, (int)ctas);
    RECEMB_LAUNCHED();
    bucket_scatter_kernel<false><<<(unsigned)ctas, kRtThreads, 0, s>>>(a);
    RECEMB_LAUNCHED();
  }
  return RECEMB_OK;
}

extern "C" int recemb_shard_bucket(const int64_t* ids, int64_t n_ids, const recemb_layout* layout, int hash_mode,
                                   int64_t num_rows, int64_t hash_arg, int zero_pad, int64_t pad_id,
                                   int32_t bag_size, const int32_t* lengths, int32_t last_n, int64_t bags_total,
                                   int64_t* entries_out, int64_t* counts_out, void* workspace,
                                   size_t workspace_bytes, int device, recemb_stream_t stream) {
  return bucket_common(nullptr, nullptr, ids, n_ids, layout, hash_mode, num_rows, hash_arg, zero_pad, pad_id,
                       bag_size, lengths, last_n, bags_total, entries_out, counts_out, workspace, workspace_bytes,
                       device, stream);
}

extern "C" int recemb_peer_bucket_push(const recemb_peer_group* group, const recemb_peer_arena* arena,
                                       const int64_t* ids, int64_t n_ids, const recemb_layout* layout,
                                       int hash_mode, int64_t num_rows, int64_t hash_arg, int zero_pad,
                                       int64_t pad_id, int32_t bag_size, const int32_t* lengths, int32_t last_n,
                                       void* workspace, size_t workspace_bytes, int device,
                                       recemb_stream_t stream) {
  RECEMB_CHECK_ARG(group && arena, "null peer group / arena");
  RECEMB_CHECK_ARG(bag_size >= 1 && n_ids == arena->bags_total * bag_size,
                   "n_ids %lld != arena bags_total %lld x bag_size %d", (long long)n_ids,
                   (long long)arena->bags_total, bag_size);
  return bucket_common(group, arena, ids, n_ids, layout, hash_mode, num_rows, hash_arg, zero_pad, pad_id, bag_size,
                       lengths, last_n, arena->bags_total, nullptr, nullptr, workspace, workspace_bytes, device,
                       stream);
}

extern "C" int recemb_peer_bucket_push_rows(const recemb_peer_group* group, const recemb_peer_arena* arena,
                                            const int64_t* ids, int64_t n_ids, const recemb_layout* layout,
                                            int hash_mode, int64_t num_rows, int64_t hash_arg, int zero_pad,
                                            int64_t pad_id, int64_t* dest_out, void* workspace,
                                            size_t workspace_bytes, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(group && arena && dest_out, "null peer group / arena / dest");
  RECEMB_CHECK_ARG(arena->bags_total == arena->cap,
                   "sequence mode needs an arena laid out with bags_total == cap (gradient buffer [world][cap][dim])");
  return bucket_common(group, arena, ids, n_ids, layout, hash_mode, num_rows, hash_arg, zero_pad, pad_id, 1, nullptr,
                       0, n_ids, nullptr, nullptr, workspace, workspace_bytes, device, stream, dest_out);
}

extern "C" int recemb_peer_rows_scatter_push(const recemb_peer_group* group, const recemb_peer_arena* arena,
                                             const void* rows, int64_t n, int32_t dim, int dtype,
                                             const int64_t* dest, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(group && arena, "null peer group / arena");
  RECEMB_CHECK_ARG(group->world >= 1 && group->world <= RECEMB_MAX_PEERS && group->rank >= 0 &&
                       group->rank < group->world, "peer group world / rank out of range");
  RECEMB_CHECK_ARG(n >= 0 && dim > 0, "bad shape");
  RECEMB_CHECK_ARG(dtype == RECEMB_F32 || dtype == RECEMB_BF16, "bad dtype");
  RECEMB_CHECK_ARG(arena->bags_total == arena->cap, "sequence mode needs an arena laid out with bags_total == cap");
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(rows && dest && (uintptr_t)rows % 16 == 0, "rows / dest null or misaligned");
  const int64_t row_bytes = (int64_t)dim * (dtype == RECEMB_F32 ? 4 : 2);
  RECEMB_UNSUPPORTED(row_bytes % 16 == 0 && row_bytes <= 512, "row of %lld bytes unsupported (16-byte multiple, <= 512)",
                     (long long)row_bytes);
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  RowsPushArgs a;
  a.rows = (const uint4*)rows;
  a.dest = dest;
  a.n = n;
  a.vecs = (int32_t)(row_bytes / 16);
  a.seq_base = (uint32_t)((int64_t)group->rank * arena->cap);
  for (int i = 0; i < RECEMB_MAX_PEERS; ++i) a.grads[i] = nullptr;
  for (int o = 0; o < group->world; ++o) {
    RECEMB_CHECK_ARG(group->arena[o], "peer arena %d not mapped", o);
    a.grads[o] = (uint4*)((char*)group->arena[o] + arena->off_grads);
  }
  int G = 1;
  while (G < a.vecs) G <<= 1;
  cudaStream_t s = (cudaStream_t)stream;
#define RSP(G_)                                                                               \
  if (G == G_) {                                                                              \
    const int64_t per = kRtThreads / G_;                                                      \
    int64_t grid = (n + per - 1) / per;                                                       \
    const int64_t cap_grid = (int64_t)sm_count(device) * 8;                                   \
    if (grid > cap_grid) grid = cap_grid;                                                     \
    rows_scatter_push_kernel<G_><<<(unsigned)grid, kRtThreads, 0, s>>>(a);                    \
  }
  RSP(1) RSP(2) RSP(4) RSP(8) RSP(16) RSP(32)
#undef RSP
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

extern "C" int recemb_pool_entries(const void* table, int32_t dim, int dtype, const int64_t* entries, int64_t n,
                                   void* out, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n >= 0 && dim > 0, "bad shape");
  RECEMB_CHECK_ARG(dtype == RECEMB_F32 || dtype == RECEMB_BF16, "bad dtype");
  RECEMB_UNSUPPORTED(n < 0x7fffffffll, "too many entries");
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(table && entries && out, "null pointer");
  const int64_t row_bytes = (int64_t)dim * (dtype == RECEMB_F32 ? 4 : 2);
  RECEMB_UNSUPPORTED(row_bytes % 16 == 0 && row_bytes <= 512, "row of %lld bytes unsupported (16-byte multiple, <= 512)",
                     (long long)row_bytes);
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  PoolEntriesArgs a;
  a.entries = entries;
  a.n = (int32_t)n;
  a.table = table;
  a.dim = (uint32_t)dim;
  a.vecs = (int32_t)(row_bytes / 16);
  a.out = out;
  int G = 1;
  while (G < a.vecs) G <<= 1;
  const int chunks = (a.n + kPeChunk - 1) / kPeChunk;
  cudaStream_t s = (cudaStream_t)stream;
#define PE(G_)                                                                                        \
  if (G == G_) {                                                                                      \
    const int groups = kRtThreads / G_;                                                               \
    const unsigned grid = (unsigned)((chunks + groups - 1) / groups);                                 \
    if (dtype == RECEMB_F32) pool_entries_kernel<G_, float><<<grid, kRtThreads, 0, s>>>(a);            \
    else pool_entries_kernel<G_, __nv_bfloat16><<<grid, kRtThreads, 0, s>>>(a);                        \
  }
  PE(1) PE(2) PE(4) PE(8) PE(16) PE(32)
#undef PE
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

// plan layout (see bwd.cu): [256 B counters][keys_in][vals_in][keys_out][vals_out][sort temp]
extern "C" int recemb_bwd_plan_entries(const int64_t* entries, int64_t n, int64_t total_rows, void* plan,
                                       size_t plan_bytes, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n >= 0 && total_rows >= 1, "bad n / total_rows");
  RECEMB_CHECK_ARG(plan && (uintptr_t)plan % 256 == 0, "plan buffer missing / misaligned");
  RECEMB_UNSUPPORTED(n < 0x7fffffffll && total_rows < 0xfffffff0ll, "sizes overflow 32-bit keys");
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(entries, "null entries");
  const size_t need = recemb_bwd_plan_bytes(n, total_rows);
  if (need == 0) return RECEMB_ERR_CUDA;
  if (plan_bytes < need) {
    set_error("plan buffer %zu < required %zu", plan_bytes, need);
    return RECEMB_ERR_WORKSPACE;
  }
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t arr = align_up((size_t)n * 4, 256);
  char* base = (char*)plan;
  uint32_t* keys_in = (uint32_t*)(base + 256);
  uint32_t* vals_in = (uint32_t*)(base + 256 + arr);
  uint32_t* keys_out = (uint32_t*)(base + 256 + 2 * arr);
  uint32_t* vals_out = (uint32_t*)(base + 256 + 3 * arr);
  void* temp = base + 256 + 4 * arr;
  size_t temp_bytes = plan_bytes - (256 + 4 * arr);
  int64_t grid = (n + kRtThreads - 1) / kRtThreads;
  const int64_t cap = (int64_t)sm_count(device) * 16;
  if (grid > cap) grid = cap;
  int bits = 0;
  for (uint64_t x = (uint64_t)total_rows; x; x >>= 1) ++bits;
  const bool in_b = sort_input_in_b(sort_shape(n, bits, device));  // the sort ends in the "out" pair buffers
  unpack_entries_kernel<<<(unsigned)grid, kRtThreads, 0, s>>>(entries, n, in_b ? keys_out : keys_in,
                                                             in_b ? vals_out : vals_in);
  RECEMB_LAUNCHED();
  return sort_pairs(keys_in, vals_in, keys_out, vals_out, n, bits, temp, temp_bytes, device, s);
}


extern "C" int recemb_peer_plan(const recemb_peer_group* group, const recemb_peer_arena* arena, int64_t total_rows,
                                void* plan, size_t plan_bytes, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(group && arena && group->world >= 1 && group->world <= RECEMB_MAX_PEERS, "bad peer group");
  RECEMB_CHECK_ARG(total_rows >= 1 && arena->cap >= 1, "bad total_rows / capacity");
  RECEMB_CHECK_ARG(plan && (uintptr_t)plan % 256 == 0, "plan buffer missing / misaligned");
  const int64_t n = (int64_t)group->world * arena->cap;
  RECEMB_UNSUPPORTED(n < 0x7fffffffll && total_rows < 0xfffffff0ll, "sizes overflow 32-bit keys");
  const size_t need = recemb_bwd_plan_bytes(n, total_rows);
  if (need == 0) return RECEMB_ERR_CUDA;
  if (plan_bytes < need) {
    set_error("plan buffer %zu < required %zu", plan_bytes, need);
    return RECEMB_ERR_WORKSPACE;
  }
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  cudaStream_t s = (cudaStream_t)stream;
  const char* mine = (const char*)group->arena[group->rank];
  RECEMB_CHECK_ARG(mine, "local arena missing");
  const size_t arr = align_up((size_t)n * 4, 256);
  char* base = (char*)plan;
  uint32_t* keys_in = (uint32_t*)(base + 256);
  uint32_t* vals_in = (uint32_t*)(base + 256 + arr);
  uint32_t* keys_out = (uint32_t*)(base + 256 + 2 * arr);
  uint32_t* vals_out = (uint32_t*)(base + 256 + 3 * arr);
  void* temp = base + 256 + 4 * arr;
  size_t temp_bytes = plan_bytes - (256 + 4 * arr);
  int64_t grid = (n + kRtThreads - 1) / kRtThreads;
  const int64_t cap_grid = (int64_t)sm_count(device) * 16;
  if (grid > cap_grid) grid = cap_grid;
  int bits = 0;  // the sentinel key == total_rows must sort last
  for (uint64_t x = (uint64_t)total_rows; x; x >>= 1) ++bits;
  const bool in_b = sort_input_in_b(sort_shape(n, bits, device));
  unpack_inbox_kernel<<<(unsigned)grid, kRtThreads, 0, s>>>((const int64_t*)(mine + arena->off_inbox),
                                                           (const int64_t*)(mine + arena->off_counts), arena->cap,
                                                           n, (uint32_t)total_rows, in_b ? keys_out : keys_in,
                                                           in_b ? vals_out : vals_in);
  RECEMB_LAUNCHED();
  return sort_pairs(keys_in, vals_in, keys_out, vals_out, n, bits, temp, temp_bytes, device, s);
}


static int peer_pool_push_impl(const recemb_peer_group* group, const recemb_peer_arena* arena, int32_t dim,
                               int dtype, int64_t bags_per_table, int device, recemb_stream_t stream);

extern "C" int recemb_peer_pool_push(const recemb_peer_group* group, const recemb_peer_arena* arena, int32_t dim,
                                     int dtype, int device, recemb_stream_t stream) {
  return peer_pool_push_impl(group, arena, dim, dtype, 0, device, stream);
}

extern "C" int recemb_peer_pool_push_tablewise(const recemb_peer_group* group, const recemb_peer_arena* arena,
                                               int32_t dim, int dtype, int64_t bags_per_table, int device,
                                               recemb_stream_t stream) {
  RECEMB_CHECK_ARG(arena && bags_per_table >= 1 && arena->bags_total % bags_per_table == 0,
                   "bags_total is not a multiple of bags_per_table");
  return peer_pool_push_impl(group, arena, dim, dtype, bags_per_table, device, stream);
}

static int peer_pool_push_impl(const recemb_peer_group* group, const recemb_peer_arena* arena, int32_t dim,
                               int dtype, int64_t bags_per_table, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(group && arena && group->world >= 1 && group->world <= RECEMB_MAX_PEERS && group->rank >= 0 &&
                       group->rank < group->world,
                   "bad peer group");
  RECEMB_CHECK_ARG(dim > 0 && (dtype == RECEMB_F32 || dtype == RECEMB_BF16), "bad dim / dtype");
  const int64_t row_bytes = (int64_t)dim * (dtype == RECEMB_F32 ? 4 : 2);
  RECEMB_UNSUPPORTED(row_bytes % 16 == 0 && row_bytes <= 512, "row of %lld bytes unsupported (16-byte multiple, <= 512)",
                     (long long)row_bytes);
  RECEMB_CHECK_ARG(arena->cap >= 1 && arena->bags_total >= 0, "bad arena");
  if (arena->bags_total == 0) return RECEMB_OK;
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  PoolInboxArgs a;
  const char* mine = (const char*)group->arena[group->rank];
  RECEMB_CHECK_ARG(mine && group->table[group->rank], "local arena / table missing");
  a.inbox = (const int64_t*)(mine + arena->off_inbox);
  a.counts = (const int64_t*)(mine + arena->off_counts);
  a.cap = arena->cap;
  a.bags_total = arena->bags_total;
  a.world = group->world;
  a.rank = group->rank;
  a.chunks_per_sender = (int32_t)((arena->cap + kPeChunk - 1) / kPeChunk);
  a.table = group->table[group->rank];
  a.vecs = (int32_t)(row_bytes / 16);
  a.bags_per_table = bags_per_table;
  for (int i = 0; i < RECEMB_MAX_PEERS; ++i) a.parts[i] = nullptr;
  for (int sd = 0; sd < group->world; ++sd) {
    RECEMB_CHECK_ARG(group->arena[sd], "peer arena %d not mapped", sd);
    a.parts[sd] = (uint4*)((char*)group->arena[sd] + arena->off_parts +
                           (int64_t)group->rank * arena->bags_total * row_bytes);
  }
  int G = 1;
  while (G < a.vecs) G <<= 1;
  const int64_t chunks = (int64_t)a.chunks_per_sender * group->world;
  cudaStream_t s = (cudaStream_t)stream;
#define PIP(G_)                                                                                       \
  if (G == G_) {                                                                                      \
    const int groups = kRtThreads / G_;                                                               \
    const unsigned grid = (unsigned)((chunks + groups - 1) / groups);                                 \
    if (dtype == RECEMB_F32) pool_inbox_push_kernel<G_, float><<<grid, kRtThreads, 0, s>>>(a);         \
    else pool_inbox_push_kernel<G_, __nv_bfloat16><<<grid, kRtThreads, 0, s>>>(a);                     \
  }
  PIP(1) PIP(2) PIP(4) PIP(8) PIP(16) PIP(32)
#undef PIP
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}
