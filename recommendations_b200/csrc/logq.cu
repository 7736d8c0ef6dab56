// Streaming logQ correction (sm_100a): the D = 1 cousin of the embedding gather / scatter.
//   forward  out[i] = min_m( -log(b_m[h_m(id_i)]) ),  h_m(id) = floor_mod(id + offset_m, buckets)
//            (StreamingLogQCorrectionModule.forward / hash_fn, commons/layers.py:202-208;
//             CascadedStreamingLogQCorrectionModule.forward, :224-232: torch.minimum over modules)
//   update   b_m[h] = (1 - alpha) * b_m[h] + alpha * (batch_idx - a_m[h]);  a_m[h] = batch_idx
//            (train_step, :210-213, with the reference's `self.alpha[hash] = batch_idx` read as
//             `self.a[hash] = batch_idx` -- `alpha` is a Python float there)
//
// Duplicate indices.  torch evaluates the right-hand side from the OLD b / a (a gather), then
// scatters.  The new value depends on the bucket only, so every duplicate writes the same number:
// there is no last-writer ambiguity to pin.  The update therefore runs as two kernels -- gather
// new values into a scratch array, then scatter -- which keeps exactly that read-all-then-write-all
// order without atomics or a dedup pass; duplicate writers race benignly on identical values.
// All M cascaded tables (lthm.yaml: 7 offsets x 2^24 buckets = 64 MB each) are served by one launch.
#include "common.cuh"

namespace recemb {

constexpr int kLogqThreads = 256;
constexpr int kLogqMaxTables = 16;

struct LogqArgs {
  float* b[kLogqMaxTables];
  float* a[kLogqMaxTables];
  int64_t offset[kLogqMaxTables];
  int32_t num_tables;
  ModN mod_buckets;
  const int64_t* ids;
  const uint8_t* skip;  // optional: 1 = this id takes no part in the update (masked position)
  int64_t n;
  float* out;      // fwd: [n]
  float* scratch;  // update: [n, num_tables] new bucket values
  float alpha;            // (float)alpha
  float one_minus_alpha;  // (float)(1.0 - alpha): the reference forms 1 - alpha in double
  float batch_idx;
};

__device__ __forceinline__ int64_t logq_bucket(int64_t id, int64_t offset, const ModN& m) {
  // int64 addition wraps in torch as well
  return floor_mod((int64_t)((uint64_t)id + (uint64_t)offset), m);
}

__global__ void __launch_bounds__(kLogqThreads) logq_fwd_kernel(const LogqArgs a) {
  int64_t i = (int64_t)blockIdx.x * kLogqThreads + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * kLogqThreads;
  for (; i < a.n; i += stride) {
    const int64_t id = a.ids[i];
    float v[kLogqMaxTables];
#pragma unroll
    for (int m = 0; m < kLogqMaxTables; ++m)  // all bucket reads in flight before the first log
      if (m < a.num_tables) v[m] = __ldg(a.b[m] + logq_bucket(id, a.offset[m], a.mod_buckets));
    float r = 0.f;
#pragma unroll
    for (int m = 0; m < kLogqMaxTables; ++m) {
      if (m < a.num_tables) {
        const float lq = -logf(v[m]);
        // torch.minimum propagates NaN (b <= 0 never happens from a positive p_init, kept anyway)
        r = (m == 0) ? lq : ((lq != lq || r != r) ? (lq + r) : fminf(r, lq));
      }
    }
    a.out[i] = r;
  }
}

// phase 1: new bucket values from the OLD tables (nothing is written to b / a)
__global__ void __launch_bounds__(kLogqThreads) logq_update_gather_kernel(const LogqArgs a) {
  int64_t i = (int64_t)blockIdx.x * kLogqThreads + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * kLogqThreads;
  for (; i < a.n; i += stride) {
    if (a.skip && a.skip[i]) continue;
    const int64_t id = a.ids[i];
    float bv[kLogqMaxTables], av[kLogqMaxTables];
#pragma unroll
    for (int m = 0; m < kLogqMaxTables; ++m) {
      if (m < a.num_tables) {
        const int64_t h = logq_bucket(id, a.offset[m], a.mod_buckets);
        bv[m] = a.b[m][h];
        av[m] = a.a[m][h];
      }
    }
#pragma unroll
    for (int m = 0; m < kLogqMaxTables; ++m) {
      if (m < a.num_tables) {
        // ((1 - alpha) * b) + (alpha * (batch_idx - a)): the reference's two products and one sum,
        // each rounded to fp32 (no FMA contraction)
        const float t1 = __fmul_rn(a.one_minus_alpha, bv[m]);
        const float t2 = __fmul_rn(a.alpha, __fsub_rn(a.batch_idx, av[m]));
        a.scratch[i * a.num_tables + m] = __fadd_rn(t1, t2);
      }
    }
  }
}

// phase 2: scatter; duplicates of a bucket carry identical values
__global__ void __launch_bounds__(kLogqThreads) logq_update_scatter_kernel(const LogqArgs a) {
  int64_t i = (int64_t)blockIdx.x * kLogqThreads + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * kLogqThreads;
  for (; i < a.n; i += stride) {
    if (a.skip && a.skip[i]) continue;
    const int64_t id = a.ids[i];
#pragma unroll
    for (int m = 0; m < kLogqMaxTables; ++m) {
      if (m < a.num_tables) {
        const int64_t h = logq_bucket(id, a.offset[m], a.mod_buckets);
        a.b[m][h] = a.scratch[i * a.num_tables + m];
        a.a[m][h] = a.batch_idx;
      }
    }
  }
}

static int fill_args(LogqArgs* a, float* const* b, float* const* av, int32_t num_tables,
                     const int64_t* offsets, int64_t num_buckets) {
  RECEMB_CHECK_ARG(num_tables >= 1 && num_tables <= kLogqMaxTables, "num_tables %d outside [1, %d]", num_tables,
                   kLogqMaxTables);
  RECEMB_CHECK_ARG(num_buckets >= 1, "num_buckets < 1");
  RECEMB_CHECK_ARG(b && offsets, "null table / offset array");
  for (int m = 0; m < kLogqMaxTables; ++m) {
    a->b[m] = m < num_tables ? b[m] : nullptr;
    a->a[m] = (m < num_tables && av) ? av[m] : nullptr;
    a->offset[m] = m < num_tables ? offsets[m] : 0;
    RECEMB_CHECK_ARG(m >= num_tables || a->b[m], "bucket table %d is null", m);
  }
  a->num_tables = num_tables;
  a->mod_buckets = make_modn((uint64_t)num_buckets);
  return RECEMB_OK;
}

static unsigned logq_grid(int device, int64_t n) {
  int64_t grid = (n + kLogqThreads - 1) / kLogqThreads;
  const int64_t cap = (int64_t)sm_count(device) * 8;
  return (unsigned)(grid > cap ? cap : grid);
}

}  // namespace recemb

using namespace recemb;

extern "C" int recemb_logq_fwd(const float* const* b_tables_host, int32_t num_tables,
                               const int64_t* hash_offsets_host, int64_t num_buckets, const int64_t* ids,
                               int64_t n, float* out, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n >= 0, "n < 0");
  LogqArgs a{};
  int rc = fill_args(&a, (float* const*)b_tables_host, nullptr, num_tables, hash_offsets_host, num_buckets);
  if (rc) return rc;
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(ids && out, "null pointer");
  a.ids = ids;
  a.n = n;
  a.out = out;
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  logq_fwd_kernel<<<logq_grid(device, n), kLogqThreads, 0, (cudaStream_t)stream>>>(a);
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

extern "C" int recemb_logq_update(float* const* b_tables_host, float* const* a_tables_host, int32_t num_tables,
                                  const int64_t* hash_offsets_host, int64_t num_buckets, const int64_t* ids,
                                  int64_t n, const uint8_t* skip_mask, double alpha, int64_t batch_idx,
                                  float* scratch, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n >= 0, "n < 0");
  RECEMB_CHECK_ARG(a_tables_host, "null `a` table array");
  LogqArgs a{};
  int rc = fill_args(&a, b_tables_host, a_tables_host, num_tables, hash_offsets_host, num_buckets);
  if (rc) return rc;
  for (int m = 0; m < num_tables; ++m) RECEMB_CHECK_ARG(a.a[m], "`a` table %d is null", m);
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(ids && scratch, "null ids / scratch");
  a.ids = ids;
  a.skip = skip_mask;
  a.n = n;
  a.scratch = scratch;
  a.alpha = (float)alpha;
  a.one_minus_alpha = (float)(1.0 - alpha);
  a.batch_idx = (float)batch_idx;
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  const unsigned grid = logq_grid(device, n);
  logq_update_gather_kernel<<<grid, kLogqThreads, 0, (cudaStream_t)stream>>>(a);
  RECEMB_LAUNCHED();
  logq_update_scatter_kernel<<<grid, kLogqThreads, 0, (cudaStream_t)stream>>>(a);
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}
