// Peer-memory plumbing of the row-wise sharded exchange (north_star item 4, cfg 5): CUDA IPC
// export / mapping of the table shards and exchange arenas, the device-side barrier over flags in
// peer memory and the push all-gather of the pooled gradients.  The lookup kernels that read and
// write the mapped pointers live next to their single-GPU versions: pool_kernel<PEER> (fwd.cu) and
// bucket_scatter_kernel<PEER> / recemb_peer_plan (route.cu).
//
// NVLink 5 / NVSwitch make every peer's HBM addressable at ~770 GB/s per direction with ~2 us
// load latency; a pooled lookup keeps 8 x 256-byte row loads in flight per half-warp, far more
// than the bandwidth-delay product needs, so the forward pull needs no staging or collective.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace recemb {

struct PeerPtrs {
  char* arena[RECEMB_MAX_PEERS];
  int32_t world;
  int32_t rank;
};

// One warp.  Lane p publishes this rank's new epoch in rank p's flag slot and waits until rank p
// has published at least the same epoch here.  The release store orders every earlier store of
// this stream (kernel boundaries are system-scope ordered, the fence makes it cumulative); the
// acquire load orders the kernels that follow.
// Timeout: a rank may legitimately stall for a long time (data loader, rank-0 checkpoint or eval), so the
// default is collective-library-like (RECEMB_PEER_BARRIER_TIMEOUT_S, default 600 s).  A rank that still
// gives up sets status bit 2 in ITS arena; the guarded update (recemb_bwd_apply_guarded) then leaves the
// table untouched for that step and the host raises at its next status check -- a stalled peer can
// delay a step or lose it, never corrupt the shard.
__global__ void __launch_bounds__(32) peer_barrier_kernel(const PeerPtrs g, int64_t off_flags, int64_t off_epoch,
                                                          int64_t off_status, long long kBarrierTimeoutCycles) {
  char* mine = g.arena[g.rank];
  uint64_t* epoch_ptr = (uint64_t*)(mine + off_epoch);
  uint64_t epoch = 0;
  if (threadIdx.x == 0) {
    epoch = *epoch_ptr + 1;
    *epoch_ptr = epoch;
  }
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  if ((int)threadIdx.x < g.world) {
    __threadfence_system();
    st_release_sys_u64((uint64_t*)(g.arena[threadIdx.x] + off_flags) + g.rank, epoch);
    const uint64_t* flag = (const uint64_t*)(mine + off_flags) + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(flag) < epoch) {
      if (clock64() - t0 > kBarrierTimeoutCycles) {
        atomicOr((uint32_t*)(mine + off_status), 2u);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncwarp();
  __threadfence_system();
}

// The two halves of the barrier as separate launches, for pipelined steps where producer and consumer
// live on different streams: `signal` (enqueued right after the producer, never blocks) publishes this
// rank's next epoch of the channel to every rank; `wait` (enqueued right before the consumer, after the
// local signal in stream / event order) spins until every rank has published the epoch this rank is at.
// A stream is then only ever blocked in front of a kernel that really needs the peers' data.
__global__ void __launch_bounds__(32) peer_signal_kernel(const PeerPtrs g, int64_t off_flags, int64_t off_epoch) {
  char* mine = g.arena[g.rank];
  uint64_t* epoch_ptr = (uint64_t*)(mine + off_epoch);
  uint64_t epoch = 0;
  if (threadIdx.x == 0) {
    epoch = *epoch_ptr + 1;
    *epoch_ptr = epoch;
  }
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  if ((int)threadIdx.x < g.world) {
    __threadfence_system();
    st_release_sys_u64((uint64_t*)(g.arena[threadIdx.x] + off_flags) + g.rank, epoch);
  }
}

__global__ void __launch_bounds__(32) peer_wait_kernel(const PeerPtrs g, int64_t off_flags, int64_t off_epoch,
                                                       int64_t off_status, long long kBarrierTimeoutCycles) {
  char* mine = g.arena[g.rank];
  const uint64_t epoch = *(const volatile uint64_t*)(mine + off_epoch);
  if ((int)threadIdx.x < g.world) {
    const uint64_t* flag = (const uint64_t*)(mine + off_flags) + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(flag) < epoch) {
      if (clock64() - t0 > kBarrierTimeoutCycles) {
        atomicOr((uint32_t*)(mine + off_status), 2u);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncwarp();
  __threadfence_system();
}

// src -> slice `rank` of the same buffer on every rank: one read, world stores per 16-byte vector
__global__ void __launch_bounds__(256) allgather_push_kernel(const PeerPtrs g, const uint4* __restrict__ src,
                                                            int64_t vecs, int64_t dst_offset) {
  __shared__ uint4* s_dst[RECEMB_MAX_PEERS];
  if ((int)threadIdx.x < g.world)
    s_dst[threadIdx.x] = (uint4*)(g.arena[threadIdx.x] + dst_offset) + (int64_t)g.rank * vecs;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * 256 * 4;
  for (int64_t i0 = (int64_t)blockIdx.x * 256 * 4 + threadIdx.x; i0 < vecs; i0 += stride) {
    uint4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * 256;
      if (i < vecs) v[u] = ldg_nc_v4(src + i);
    }
    for (int q = 1; q <= g.world; ++q) {  // rotated: all ranks never store into the same peer at once
      int p = g.rank + q;
      if (p >= g.world) p -= g.world;
      uint4* dst = s_dst[p];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + u * 256;
        if (i < vecs) stg_v4(dst + i, v[u]);
      }
    }
  }
}

static int make_ptrs(const recemb_peer_group* group, PeerPtrs* out) {
  RECEMB_CHECK_ARG(group && group->world >= 1 && group->world <= RECEMB_MAX_PEERS && group->rank >= 0 &&
                       group->rank < group->world,
                   "peer group world / rank out of range");
  for (int i = 0; i < RECEMB_MAX_PEERS; ++i) out->arena[i] = nullptr;
  for (int i = 0; i < group->world; ++i) {
    RECEMB_CHECK_ARG(group->arena[i] && (uintptr_t)group->arena[i] % 256 == 0, "peer arena %d null / misaligned", i);
    out->arena[i] = (char*)group->arena[i];
  }
  out->world = group->world;
  out->rank = group->rank;
  return RECEMB_OK;
}

typedef int (*cuMemGetAddressRange_fn)(unsigned long long* pbase, size_t* psize, unsigned long long dptr);

}  // namespace recemb

using namespace recemb;

extern "C" int recemb_peer_export(const void* ptr, uint8_t handle_out[RECEMB_PEER_HANDLE_BYTES],
                                  int64_t* offset_out, int64_t* alloc_bytes_out, int device) {
  RECEMB_CHECK_ARG(ptr && handle_out && offset_out, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == RECEMB_PEER_HANDLE_BYTES, "IPC handle size");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  // base of the allocation that contains ptr (the caching allocator of the host framework hands
  // out interior pointers): driver entry point through the runtime, no link against libcuda
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  RECEMB_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess) {
    set_error("cuMemGetAddressRange entry point unavailable");
    return RECEMB_ERR_CUDA;
  }
  unsigned long long base = 0;
  size_t size = 0;
  const int drc = ((cuMemGetAddressRange_fn)fn)(&base, &size, (unsigned long long)(uintptr_t)ptr);
  if (drc != 0) {
    set_error("cuMemGetAddressRange failed (CUresult %d)", drc);
    return RECEMB_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  RECEMB_CUDA(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
  memcpy(handle_out, &h, sizeof(h));
  *offset_out = (int64_t)((uintptr_t)ptr - (uintptr_t)base);
  if (alloc_bytes_out) *alloc_bytes_out = (int64_t)size;
  return RECEMB_OK;
}

extern "C" int recemb_peer_open(const uint8_t handle[RECEMB_PEER_HANDLE_BYTES], void** base_out, int device) {
  RECEMB_CHECK_ARG(handle && base_out, "null pointer");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  RECEMB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *base_out = p;
  return RECEMB_OK;
}

extern "C" int recemb_peer_close(void* base, int device) {
  if (!base) return RECEMB_OK;
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  RECEMB_CUDA(cudaIpcCloseMemHandle(base));
  return RECEMB_OK;
}

extern "C" int recemb_peer_arena_layout(int32_t world, int64_t cap, int64_t bags_total, int32_t dim, int dtype,
                                        recemb_peer_arena* out) {
  RECEMB_CHECK_ARG(out, "null out");
  RECEMB_CHECK_ARG(world >= 1 && world <= RECEMB_MAX_PEERS, "world %d outside [1, %d]", world, RECEMB_MAX_PEERS);
  RECEMB_CHECK_ARG(cap >= 1 && bags_total >= 0 && dim > 0, "bad cap / bags_total / dim");
  RECEMB_CHECK_ARG(dtype == RECEMB_F32 || dtype == RECEMB_BF16, "bad dtype");
  const int64_t row_bytes = (int64_t)dim * (dtype == RECEMB_F32 ? 4 : 2);
  RECEMB_UNSUPPORTED(row_bytes % 16 == 0, "row bytes not a multiple of 16");
  RECEMB_UNSUPPORTED((int64_t)world * cap < 0x7fffffffll, "inbox too large for 32-bit slots");
  int64_t off = 0;
  out->off_flags = off;
  off += align_up((size_t)RECEMB_PEER_CHANNELS * RECEMB_MAX_PEERS * 8, 256);
  out->off_epoch = off;
  off += 128;
  out->off_status = off;
  off += 128;
  out->off_counts = off;
  off += 256;
  out->off_gate = off;
  off += (int64_t)align_up((size_t)kGateBytes, 256);
  out->off_inbox = off;
  off += (int64_t)align_up((size_t)((int64_t)world * cap * 8), 256);
  out->off_grads = off;
  off += (int64_t)align_up((size_t)((int64_t)world * bags_total * row_bytes), 256);
  out->off_parts = off;
  off += (int64_t)align_up((size_t)((int64_t)world * bags_total * row_bytes), 256);
  out->bytes = off;
  out->cap = cap;
  out->bags_total = bags_total;
  return RECEMB_OK;
}

static long long barrier_timeout_cycles() {
  const char* e = getenv("RECEMB_PEER_BARRIER_TIMEOUT_S");
  double sec = e ? atof(e) : 600.0;
  if (!(sec > 0.0)) sec = 600.0;
  return (long long)(sec * 2.0e9);  // clock64 ticks at <= 2 GHz: at least `sec` seconds
}

// what: 0 = barrier (signal + wait in one launch), 1 = signal only, 2 = wait only
static int barrier_launch(const recemb_peer_group* group, const recemb_peer_arena* arena, int channel, int what,
                          int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(arena, "null arena");
  RECEMB_CHECK_ARG(channel >= 0 && channel < RECEMB_PEER_CHANNELS, "barrier channel %d out of range", channel);
  PeerPtrs p;
  int rc = make_ptrs(group, &p);
  if (rc) return rc;
  if (p.world == 1) return RECEMB_OK;
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  const long long timeout_cycles = barrier_timeout_cycles();
  const int64_t off_flags = arena->off_flags + (int64_t)channel * RECEMB_MAX_PEERS * 8;
  const int64_t off_epoch = arena->off_epoch + (int64_t)channel * 8;
  cudaStream_t s = (cudaStream_t)stream;
  if (what == 0) peer_barrier_kernel<<<1, 32, 0, s>>>(p, off_flags, off_epoch, arena->off_status, timeout_cycles);
  else if (what == 1) peer_signal_kernel<<<1, 32, 0, s>>>(p, off_flags, off_epoch);
  else peer_wait_kernel<<<1, 32, 0, s>>>(p, off_flags, off_epoch, arena->off_status, timeout_cycles);
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

extern "C" int recemb_peer_barrier(const recemb_peer_group* group, const recemb_peer_arena* arena, int channel,
                                   int device, recemb_stream_t stream) {
  return barrier_launch(group, arena, channel, 0, device, stream);
}

extern "C" int recemb_peer_signal(const recemb_peer_group* group, const recemb_peer_arena* arena, int channel,
                                  int device, recemb_stream_t stream) {
  return barrier_launch(group, arena, channel, 1, device, stream);
}

extern "C" int recemb_peer_wait(const recemb_peer_group* group, const recemb_peer_arena* arena, int channel,
                                int device, recemb_stream_t stream) {
  return barrier_launch(group, arena, channel, 2, device, stream);
}

extern "C" int recemb_peer_allgather_push(const recemb_peer_group* group, const void* src, int64_t bytes,
                                          int64_t dst_offset, int device, recemb_stream_t stream) {
  PeerPtrs p;
  int rc = make_ptrs(group, &p);
  if (rc) return rc;
  RECEMB_CHECK_ARG(bytes >= 0 && bytes % 16 == 0 && dst_offset >= 0 && dst_offset % 16 == 0,
                   "bytes / dst_offset must be 16-byte multiples");
  if (bytes == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(src && (uintptr_t)src % 16 == 0, "src null / misaligned");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  const int64_t vecs = bytes / 16;
  int64_t grid = (vecs + 1023) / 1024;
  const int64_t cap = (int64_t)sm_count(device) * 4;
  if (grid > cap) grid = cap;
  allgather_push_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(p, (const uint4*)src, vecs, dst_offset);
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}
