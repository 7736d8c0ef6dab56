// Library-level plumbing of recemb_b200: error text, launch counter, hash-spec
// construction, and the host-buffer end-to-end entry point.
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace recemb {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launch_count{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

int sm_count(int device) {
  static std::mutex mu;
  static int cached[64];
  if (device < 0 || device >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (cached[device] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0)
      v = 148;
    cached[device] = v;
  }
  return cached[device];
}

int64_t layout_local_rows(int64_t num_rows, const recemb_layout* layout) {
  if (!layout || layout->shard_world <= 1) return num_rows;
  return (num_rows - layout->shard_rank + layout->shard_world - 1) / layout->shard_world;
}
int64_t layout_tables(const recemb_layout* layout, int64_t n_ids) {
  if (!layout || layout->ids_per_table <= 0) return 1;
  if (layout->num_tables > 0) return layout->num_tables;
  const int64_t t = (n_ids + layout->ids_per_table - 1) / layout->ids_per_table;
  return t > 0 ? t : 1;
}

int make_hash_spec(int hash_mode, int64_t num_rows, int64_t hash_arg, HashSpec* out,
                   const recemb_layout* layout) {
  HashSpec h;
  h.mode = hash_mode;
  h.shift = 0;
  h.ids_per_table = 0;
  h.num_tables = 0;
  h.rows_per_table = num_rows;
  h.shard_world = 1;
  h.shard_rank = 0;
  h.mod_world = make_modn(1);
  h.flip_len = 0;
  h.win_len = 0;
  h.win_side = 0;
  h.win_keep = nullptr;
  h.partition = 0;
  h.out_feats = h.out_feat_off = h.out_bpt = 0;  // pool_fwd_common / recemb_bwd_plan fill out_bpt from the bag size
  if (layout) {
    RECEMB_CHECK_ARG(layout->out_features >= 0 && layout->out_feature_offset >= 0, "out_features / offset < 0");
    RECEMB_CHECK_ARG(layout->partition == 0 || layout->partition == 1, "partition must be 0 (row-wise) or 1 (table-wise)");
    h.partition = (uint32_t)layout->partition;
    if (layout->out_features > 0) {
      RECEMB_CHECK_ARG(layout->ids_per_table > 0, "out_features needs table-batched lookups (ids_per_table)");
      h.out_feats = (uint32_t)layout->out_features;
      h.out_feat_off = (uint32_t)layout->out_feature_offset;
    }
    RECEMB_CHECK_ARG(layout->flip_len >= 0, "flip_len < 0");
    h.flip_len = (uint32_t)layout->flip_len;
    if (layout->window_keep) {
      RECEMB_CHECK_ARG(layout->seq_len > 0, "a sequence window needs seq_len > 0");
      RECEMB_CHECK_ARG(layout->flip_len == 0 || layout->flip_len == layout->seq_len,
                       "flip_len must equal seq_len when a window is set");
      RECEMB_CHECK_ARG(layout->window_side == 0 || layout->window_side == 1, "window_side must be 0 or 1");
      h.win_len = (uint32_t)layout->seq_len;
      h.win_side = (uint32_t)layout->window_side;
      h.win_keep = layout->window_keep;
    }
    RECEMB_CHECK_ARG(layout->ids_per_table >= 0 && layout->ids_per_table < 0xffffffffll,
                     "ids_per_table out of range");
    RECEMB_CHECK_ARG(layout->num_tables >= 0, "num_tables < 0");
    h.ids_per_table = (uint32_t)layout->ids_per_table;
    h.num_tables = (uint32_t)layout->num_tables;
    if (layout->shard_world > 1) {
      RECEMB_CHECK_ARG(layout->shard_rank >= 0 && layout->shard_rank < layout->shard_world,
                       "shard_rank %d outside [0, %d)", layout->shard_rank, layout->shard_world);
      h.shard_world = (uint32_t)layout->shard_world;
      h.shard_rank = (uint32_t)layout->shard_rank;
      h.mod_world = make_modn((uint64_t)layout->shard_world);
    }
    h.rows_per_table = layout_local_rows(num_rows, layout);
  }
  h.mod_rows = make_modn(1);
  h.mod_sq = make_modn(1);
  switch (hash_mode) {
    case RECEMB_HASH_IDENTITY:
      RECEMB_CHECK_ARG(num_rows >= 1, "IDENTITY needs num_rows >= 1 (the range ids are checked against)");
      h.mod_rows.n = (uint64_t)num_rows;  // range bound only; no division in this mode
      break;
    case RECEMB_HASH_FLOORMOD:
      RECEMB_CHECK_ARG(num_rows >= 1, "FLOORMOD needs num_rows >= 1");
      h.mod_rows = make_modn((uint64_t)num_rows);
      break;
    case RECEMB_HASH_ROTL_FLOORMOD:
      RECEMB_CHECK_ARG(num_rows >= 1, "ROTL_FLOORMOD needs num_rows >= 1");
      RECEMB_CHECK_ARG(hash_arg >= 0 && hash_arg <= 63, "rotation %lld out of [0, 63]",
                       (long long)hash_arg);
      h.shift = (int)hash_arg;
      h.mod_rows = make_modn((uint64_t)num_rows);
      break;
    case RECEMB_HASH_QR_QUOTIENT:
    case RECEMB_HASH_QR_REMAINDER:
      RECEMB_CHECK_ARG(hash_arg >= 1 && hash_arg < (1ll << 31), "QR divisor %lld out of range",
                       (long long)hash_arg);
      h.mod_rows = make_modn((uint64_t)hash_arg);
      h.mod_sq = make_modn((uint64_t)hash_arg * (uint64_t)hash_arg);
      break;
    case RECEMB_HASH_DIV_FLOORMOD:
      RECEMB_CHECK_ARG(num_rows >= 1, "DIV_FLOORMOD needs num_rows >= 1");
      RECEMB_CHECK_ARG(hash_arg >= 1, "DIV_FLOORMOD divisor %lld < 1", (long long)hash_arg);
      h.mod_rows = make_modn((uint64_t)num_rows);
      h.mod_sq = make_modn((uint64_t)hash_arg);
      break;
    default:
      set_error("unknown hash mode %d", hash_mode);
      return RECEMB_ERR_INVALID;
  }
  *out = h;
  return RECEMB_OK;
}

}  // namespace recemb

using namespace recemb;

extern "C" int64_t recemb_layout_total_rows(int64_t num_rows, const recemb_layout* layout, int64_t n_ids) {
  return layout_local_rows(num_rows, layout) * layout_tables(layout, n_ids);
}
extern "C" int recemb_abi_version(void) { return RECEMB_ABI_VERSION; }
extern "C" const char* recemb_last_error(void) { return t_error; }
extern "C" uint64_t recemb_launch_count(void) { return g_launch_count.load(); }

extern "C" int recemb_flat_step_host(const int64_t* ids_host, int64_t n, int64_t ids_per_table,
                                     int64_t* ids_dev_scratch,
                                     void* table, int64_t num_rows, int32_t dim, int dtype,
                                     void* out, const void* grad, int update, void* state1,
                                     void* state2, const recemb_optim_params* hp_host, void* plan,
                                     size_t plan_bytes, void* workspace, size_t workspace_bytes,
                                     int64_t* counters_host, void* wait_event_after_copy,
                                     recemb_stream_t plan_stream, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(ids_host && ids_dev_scratch, "null ids");
  RECEMB_CHECK_ARG(n >= 0, "n < 0");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  cudaStream_t s = (cudaStream_t)stream;
  cudaStream_t ps = (cudaStream_t)plan_stream;
  const bool fork = plan_stream != nullptr && plan_stream != stream;
  // fork / join events of this thread (one pair per device, created on first use)
  static thread_local cudaEvent_t ev_fork[64] = {}, ev_join[64] = {};
  if (fork) {
    RECEMB_CHECK_ARG(device >= 0 && device < 64, "device index out of range");
    if (!ev_fork[device]) {
      RECEMB_CUDA(cudaEventCreateWithFlags(&ev_fork[device], cudaEventDisableTiming));
      RECEMB_CUDA(cudaEventCreateWithFlags(&ev_join[device], cudaEventDisableTiming));
    }
  }
  if (n > 0)
    RECEMB_CUDA(cudaMemcpyAsync(ids_dev_scratch, ids_host, (size_t)n * 8, cudaMemcpyHostToDevice, s));
  // software pipelining across steps: the copy above may run while the previous step (on
  // another stream) is still computing; the kernels below must not.
  if (wait_event_after_copy)
    RECEMB_CUDA(cudaStreamWaitEvent(s, (cudaEvent_t)wait_event_after_copy, 0));
  recemb_layout layout = {ids_per_table, 0, 1, 0, 0, 0, 0, nullptr, 0, 0, 0};  // no sharding, no flip, no window
  int rc;
  if (fork) {
    // the plan (hash + radix sort) only needs the ids: it runs on plan_stream while the gather
    // streams rows on `stream`; both wait for the copy, the update waits for both
    RECEMB_CUDA(cudaEventRecord(ev_fork[device], s));
    RECEMB_CUDA(cudaStreamWaitEvent(ps, ev_fork[device], 0));
    rc = recemb_bwd_plan(ids_dev_scratch, n, &layout, 1, RECEMB_HASH_FLOORMOD, num_rows, 0, 0,
                         0, -1, 0, nullptr, 0, plan, plan_bytes, device, plan_stream);
    if (rc) return rc;
    RECEMB_CUDA(cudaEventRecord(ev_join[device], ps));
  }
  rc = recemb_gather_fwd(table, num_rows, nullptr, 0, dim, dtype, ids_dev_scratch, n,
                         &layout, RECEMB_HASH_FLOORMOD, 0, 0, RECEMB_EPI_NONE, 0, 0, out, nullptr, device,
                         stream);
  if (rc) return rc;
  if (fork) {
    RECEMB_CUDA(cudaStreamWaitEvent(s, ev_join[device], 0));
  } else {
    rc = recemb_bwd_plan(ids_dev_scratch, n, &layout, 1, RECEMB_HASH_FLOORMOD, num_rows, 0, 0,
                         0, -1, 0, nullptr, 0, plan, plan_bytes, device, stream);
    if (rc) return rc;
  }
  const int64_t total_rows = recemb_layout_total_rows(num_rows, &layout, n);
  rc = recemb_bwd_apply(plan, plan_bytes, n, grad, dtype, n, dim, 1, nullptr, nullptr, update, table,
                        dtype, total_rows, state1, state2, hp_host, workspace, workspace_bytes, device,
                        stream);
  if (rc) return rc;
  if (counters_host) {
    rc = recemb_plan_count(plan, plan_bytes, n, total_rows, device, stream);
    if (rc) return rc;
    RECEMB_CUDA(cudaMemcpyAsync(counters_host, plan, 16, cudaMemcpyDeviceToHost, s));
  }
  return RECEMB_OK;
}
