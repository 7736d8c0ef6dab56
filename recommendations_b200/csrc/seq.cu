// Sequence front end of the LTHM query tower (sm_100a):
//   seq_window_kernel - the batch-wide trim of all-pad columns (models/lthm/sequence/query_tower.py:73-79)
//                       computed on the device, result left in device memory for the windowed gather /
//                       k-shift / plan (recemb_layout.window_keep) -- the rows outside the window are then
//                       never moved at all
#include "common.cuh"

namespace recemb {

constexpr int kSeqThreads = 256;
constexpr int kSeqMaxLen = 8192;

struct WinArgs {
  const void* data;     // kind 0: int64 ids [batch, L] (padded = pad_id); kind 1: uint8 mask [batch, L] (non-zero = padded)
  int kind;
  int64_t batch;
  int32_t L;
  int64_t pad_id;
  int32_t min_keep;     // export_span: at least this many columns stay
  int side;             // 0: all-pad columns are dropped at the END of the sequences, 1: at the START
  uint32_t* col_live;   // [L] zero-filled: column holds at least one real token
  uint32_t* done;       // zero-filled ticket counter
  int32_t* out;         // [0] = keep, [1] = trim
  int64_t per_cta;      // elements per CTA (multiple of 16)
};

__global__ void __launch_bounds__(kSeqThreads) seq_window_kernel(const WinArgs a) {
  __shared__ uint32_t s_live[kSeqMaxLen];
  __shared__ int s_cnt, s_first, s_last, s_is_last;
  for (int c = threadIdx.x; c < a.L; c += kSeqThreads) s_live[c] = 0;
  if (threadIdx.x == 0) {
    s_cnt = 0;
    s_first = a.L;
    s_last = -1;
  }
  __syncthreads();
  const int64_t total = a.batch * (int64_t)a.L;
  const int64_t i0 = (int64_t)blockIdx.x * a.per_cta;
  const int64_t i1 = min(total, i0 + a.per_cta);
  if (a.kind == 0) {
    const int64_t* ids = (const int64_t*)a.data;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += kSeqThreads)
      if (ids[i] != a.pad_id) s_live[(uint32_t)(i % a.L)] = 1u;  // racing stores of the same value
  } else {
    const uint8_t* m = (const uint8_t*)a.data;
    for (int64_t i = i0 + threadIdx.x; i < i1; i += kSeqThreads)
      if (m[i] == 0) s_live[(uint32_t)(i % a.L)] = 1u;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < a.L; c += kSeqThreads)
    if (s_live[c]) a.col_live[c] = 1u;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_is_last = atomicAdd(a.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  int cnt = 0, first = a.L, last = -1;  // all-pad columns; first / last column with a token
  for (int c = threadIdx.x; c < a.L; c += kSeqThreads) {
    if (__ldcg(a.col_live + c)) {
      first = min(first, c);
      last = max(last, c);
    } else {
      ++cnt;
    }
  }
  atomicAdd(&s_cnt, cnt);
  atomicMin(&s_first, first);
  atomicMax(&s_last, last);
  __syncthreads();
  if (threadIdx.x == 0) {
    // query_tower.py:75-79: more all-pad columns than L - export_span -> exactly export_span columns stay;
    // otherwise the run of all-pad columns at the padded end goes
    const int floor_trim = max(a.L - a.min_keep, 0);
    const int lead = a.side ? s_first : (a.L - 1 - s_last);
    const int trim = s_cnt > floor_trim ? floor_trim : lead;
    a.out[0] = a.L - trim;
    a.out[1] = trim;
  }
}

}  // namespace recemb

using namespace recemb;

extern "C" size_t recemb_sequence_window_workspace_bytes(int32_t seq_len) {
  return seq_len > 0 ? align_up(((size_t)seq_len + 1) * 4, 256) : 256;
}

extern "C" int recemb_sequence_window(const void* data, int kind, int64_t batch, int32_t seq_len, int64_t pad_id,
                                      int32_t min_keep, int window_side, void* workspace, size_t workspace_bytes,
                                      int32_t* out, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(data && workspace && out, "null pointer");
  RECEMB_CHECK_ARG(kind == 0 || kind == 1, "kind must be 0 (int64 ids) or 1 (uint8 mask)");
  RECEMB_CHECK_ARG(batch >= 1 && seq_len >= 1 && min_keep >= 0, "bad batch / seq_len / min_keep");
  RECEMB_CHECK_ARG(window_side == 0 || window_side == 1, "window_side must be 0 or 1");
  RECEMB_UNSUPPORTED(seq_len <= kSeqMaxLen, "seq_len %d > %d", seq_len, kSeqMaxLen);
  RECEMB_CHECK_ARG(workspace_bytes >= recemb_sequence_window_workspace_bytes(seq_len), "workspace too small");
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  cudaStream_t s = (cudaStream_t)stream;
  RECEMB_CUDA(cudaMemsetAsync(workspace, 0, ((size_t)seq_len + 1) * 4, s));
  WinArgs a;
  a.data = data;
  a.kind = kind;
  a.batch = batch;
  a.L = seq_len;
  a.pad_id = pad_id;
  a.min_keep = min_keep;
  a.side = window_side;
  a.col_live = (uint32_t*)workspace;
  a.done = (uint32_t*)workspace + seq_len;
  a.out = out;
  const int64_t total = batch * (int64_t)seq_len;
  int64_t ctas = (int64_t)sm_count(device) * 4;
  int64_t per = (total + ctas - 1) / ctas;
  per = (per + 4095) / 4096 * 4096;  // at least a few loads per thread
  a.per_cta = per;
  ctas = (total + per - 1) / per;
  seq_window_kernel<<<(unsigned)ctas, kSeqThreads, 0, s>>>(a);
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}

// ---- QueryTower's input sum as one kernel (models/lthm/sequence/query_tower.py:89-104) -------------------
//   x = inp_proj(input) + action_embedding(labels) + hod(ts) + how(ts) + dow(ts);  x = where(mask, pad, x)
// The reference runs four gathers (each writes [B, L, D]), four adds (two reads + one write each) and a
// where: ~19 passes over [B, L, D].  Here: one read of the dense term, one write of the result; the tables
// have 4 / 24 / 168 / 7 rows and live in L1.  fp32 adds in the reference's left-to-right order (bit-exact).
namespace recemb {

constexpr int kMaxTerms = 8;

struct MultiGatherArgs {
  const uint4* base;  // [n, dim] or nullptr (zeros)
  const uint4* table[kMaxTerms];
  const int64_t* ids[kMaxTerms];
  HashSpec h[kMaxTerms];
  int32_t num_terms;
  int64_t n;
  int32_t vecs;
  const uint8_t* mask;      // [n] non-zero = take masked_row
  const uint4* masked_row;  // [dim]
  uint4* out;
};

// Two phases per round of 256 positions: (1) thread t hashes ALL terms of position t once (a 64-bit
// divide-by-constant each; done by every lane of a row group it cost 16x the instructions and the kernel sat at
// 64 % issue-active with DRAM at 31 %) and leaves the rows in shared memory; (2) groups of G lanes move the rows.
template <int G, typename T>
__global__ void __launch_bounds__(kSeqThreads) multi_gather_add_kernel(const MultiGatherArgs a) {
  constexpr int E = Vec16<T>::kElems;
  constexpr int GROUPS = kSeqThreads / G;
  __shared__ int32_t s_row[kMaxTerms][kSeqThreads];
  __shared__ uint8_t s_masked[kSeqThreads];
  const int lig = threadIdx.x % G, grp = threadIdx.x / G;
  for (int64_t base = (int64_t)blockIdx.x * kSeqThreads; base < a.n; base += (int64_t)gridDim.x * kSeqThreads) {
    const int64_t i = base + threadIdx.x;
    if (i < a.n) {
      const bool m = a.mask && a.mask[i];
      s_masked[threadIdx.x] = m;
#pragma unroll
      for (int k = 0; k < kMaxTerms; ++k)
        if (k < a.num_terms) s_row[k][threadIdx.x] = m ? -1 : (int32_t)row_of(a.ids[k][i], a.h[k]);
    }
    __syncthreads();
    const int cnt = (int)min((int64_t)kSeqThreads, a.n - base);
    if (lig < a.vecs) {
      for (int p = grp; p < cnt; p += GROUPS) {
        uint4* dst = a.out + (base + p) * a.vecs + lig;
        if (s_masked[p]) {
          stg_cs_v4(dst, ldg_nc_l1_v4(a.masked_row + lig));
          continue;
        }
        float acc[E];
        if (a.base) {
          Vec16<T>::unpack(ldg_nc_v4(a.base + (base + p) * a.vecs + lig), acc);
        } else {
#pragma unroll
          for (int e = 0; e < E; ++e) acc[e] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < kMaxTerms; ++k) {
          if (k >= a.num_terms) break;
          const int32_t row = s_row[k][p];
          if (row < 0) continue;  // out-of-range identity id: contributes nothing
          float f[E];
          Vec16<T>::unpack(ldg_nc_l1_v4(a.table[k] + (int64_t)row * a.vecs + lig), f);
#pragma unroll
          for (int e = 0; e < E; ++e) acc[e] += f[e];
          if (sizeof(T) == 2) {  // every add of the reference rounds to the tensor dtype
            uint4 t = Vec16<T>::pack(acc);
            Vec16<T>::unpack(t, acc);
          }
        }
        stg_cs_v4(dst, Vec16<T>::pack(acc));
      }
    }
    __syncthreads();  // the rows of this round are consumed before the next round overwrites them
  }
}

}  // namespace recemb

extern "C" int recemb_multi_gather_add_fwd(const void* base, const recemb_gather_term* terms_host, int32_t num_terms,
                                           int64_t n, int32_t dim, int dtype, const uint8_t* mask,
                                           const void* masked_row, void* out, int device, recemb_stream_t stream) {
  RECEMB_CHECK_ARG(n >= 0 && dim > 0, "bad shape");
  RECEMB_CHECK_ARG(num_terms >= 0 && num_terms <= kMaxTerms, "num_terms %d outside [0, %d]", num_terms, kMaxTerms);
  RECEMB_CHECK_ARG(dtype == RECEMB_F32 || dtype == RECEMB_BF16, "bad dtype");
  if (n == 0) return RECEMB_OK;
  RECEMB_CHECK_ARG(out && (num_terms == 0 || terms_host), "null pointer");
  RECEMB_CHECK_ARG(!mask || masked_row, "a mask needs the row that masked positions receive");
  const int64_t row_bytes = (int64_t)dim * (dtype == RECEMB_F32 ? 4 : 2);
  RECEMB_UNSUPPORTED(row_bytes % 16 == 0 && row_bytes <= 512, "row of %lld bytes unsupported (16-byte multiple, <= 512)",
                     (long long)row_bytes);
  RECEMB_CHECK_ARG(((uintptr_t)base | (uintptr_t)out | (uintptr_t)masked_row) % 16 == 0, "base / out / masked_row misaligned");
  MultiGatherArgs a;
  a.base = (const uint4*)base;
  a.num_terms = num_terms;
  a.n = n;
  a.vecs = (int32_t)(row_bytes / 16);
  a.mask = mask;
  a.masked_row = (const uint4*)masked_row;
  a.out = (uint4*)out;
  for (int k = 0; k < kMaxTerms; ++k) {
    a.table[k] = nullptr;
    a.ids[k] = nullptr;
  }
  for (int k = 0; k < num_terms; ++k) {
    RECEMB_CHECK_ARG(terms_host[k].table && terms_host[k].ids && (uintptr_t)terms_host[k].table % 16 == 0,
                     "term %d: null / misaligned table or ids", k);
    RECEMB_UNSUPPORTED(terms_host[k].num_rows >= 1 && terms_host[k].num_rows < 0x7fffffffll,
                       "term %d: num_rows outside [1, 2^31)", k);
    a.table[k] = (const uint4*)terms_host[k].table;
    a.ids[k] = terms_host[k].ids;
    int rc = make_hash_spec(terms_host[k].hash_mode, terms_host[k].num_rows, terms_host[k].hash_arg, &a.h[k]);
    if (rc) return rc;
  }
  DeviceGuard g(device);
  RECEMB_CUDA(g.err);
  int G = 1;
  while (G < a.vecs) G <<= 1;
  cudaStream_t s = (cudaStream_t)stream;
#define MGA(G_)                                                                                   \
  if (G == G_) {                                                                                  \
    int64_t grid = (n + kSeqThreads - 1) / kSeqThreads;                                           \
    const int64_t cap_grid = (int64_t)sm_count(device) * 16;                                      \
    if (grid > cap_grid) grid = cap_grid;                                                         \
    if (dtype == RECEMB_F32) multi_gather_add_kernel<G_, float><<<(unsigned)grid, kSeqThreads, 0, s>>>(a);       \
    else multi_gather_add_kernel<G_, __nv_bfloat16><<<(unsigned)grid, kSeqThreads, 0, s>>>(a);                    \
  }
  MGA(1) MGA(2) MGA(4) MGA(8) MGA(16) MGA(32)
#undef MGA
  RECEMB_LAUNCHED();
  return RECEMB_OK;
}
