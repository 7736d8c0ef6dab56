/*
 * recemb_b200.h -- C ABI of the B200 (sm_100a) embedding hot path.
 *
 * The reference (ranjanbalappa-nykaa/recommendations) is pure Python: it has no
 * FFI of its own.  The boundary this library sits behind is the nn.Module
 * surface of commons/layers.py and commons/transformers/layers.py; every entry
 * point below names the reference call site whose ATen library call it
 * replaces.  The Python host layer (the modules under recommendations_b200/) binds these
 * symbols with ctypes and keeps the reference constructors / forward /
 * state_dict keys (see INTEGRATION.md for the binding a maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless the name
 *     ends in _host; the library never allocates, frees or retains pointers.
 *   - `device` / `stream` are explicit (backward runs on the autograd engine
 *     thread, not the thread that ran forward); all work is asynchronous on
 *     `stream` (a cudaStream_t passed as void*).
 *   - return value: 0 = RECEMB_OK, negative = error; recemb_last_error() gives
 *     the thread-local message.  Nothing throws across the ABI.
 *   - ids are signed 64-bit (xxh64 - 2^63, commons/feature_utils.py:40-46).
 *   - tables are row-major [num_rows, dim], fp32 or bf16.
 */
#ifndef RECEMB_B200_H_
#define RECEMB_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RECEMB_ABI_VERSION 1

#if defined(__GNUC__)
#define RECEMB_API __attribute__((visibility("default")))
#else
#define RECEMB_API
#endif

typedef void* recemb_stream_t; /* cudaStream_t */

enum recemb_status {
  RECEMB_OK = 0,
  RECEMB_ERR_INVALID = -1,     /* bad argument */
  RECEMB_ERR_CUDA = -2,        /* CUDA runtime error (see recemb_last_error) */
  RECEMB_ERR_UNSUPPORTED = -3, /* shape / dtype not supported by the kernels */
  RECEMB_ERR_WORKSPACE = -4    /* caller-provided workspace too small */
};

enum recemb_dtype { RECEMB_F32 = 0, RECEMB_BF16 = 1 };

/* id -> row index transforms (all bit-exact restatements of the reference):
 *   IDENTITY      row = id                                   (ids already rows)
 *   FLOORMOD      row = floor_mod(id, num_rows)              commons/layers.py:57, :185 (col 0)
 *   ROTL_FLOORMOD row = floor_mod((id << c) | (id >> (64-c)), num_rows), c = hash_arg
 *                 wrapping <<, ARITHMETIC >> on signed int64  commons/layers.py:174-185
 *   QR_QUOTIENT   d = hash_arg; x = floor_mod(id, d*d); row = floor_mod(floor(x / d), d)
 *   QR_REMAINDER  d = hash_arg; x = floor_mod(id, d*d); row = floor_mod(x, d)
 *                                                             commons/layers.py:115-119
 *   DIV_FLOORMOD  row = floor_mod(floor_divide(id, hash_arg), num_rows)
 *                 PatternFromTimelocal (hour-of-day / hour-of-week / day-of-week tables of
 *                 QueryTower, models/lthm/sequence/query_tower.py:27-33)  commons/layers.py:39-41
 */
enum recemb_hash {
  RECEMB_HASH_IDENTITY = 0,
  RECEMB_HASH_FLOORMOD = 1,
  RECEMB_HASH_ROTL_FLOORMOD = 2,
  RECEMB_HASH_QR_QUOTIENT = 3,
  RECEMB_HASH_QR_REMAINDER = 4,
  RECEMB_HASH_DIV_FLOORMOD = 5
};

/* forward epilogues */
enum recemb_epilogue {
  RECEMB_EPI_NONE = 0,
  RECEMB_EPI_L2NORM = 1,   /* x / max(||x||_2, 1e-12)   F.normalize, commons/layers.py:60,:121,:168 */
  RECEMB_EPI_RSQRT_K = 2   /* x / sqrt(num_shifts)      commons/layers.py:170 */
};

/* pooled-bag modes (nn.EmbeddingBag(mode='sum') at commons/transformers/layers.py:457;
 * mean / last_n are the build-defined variants named by BASELINE.json north_star) */
enum recemb_pool { RECEMB_POOL_SUM = 0, RECEMB_POOL_MEAN = 1 };

/* what recemb_bwd_apply does with the de-duplicated per-row gradient sums */
enum recemb_update {
  RECEMB_UPD_DENSE_GRAD = 0,      /* grad_weight[row] = sum (torch-compatible .grad; caller zero-fills) */
  RECEMB_UPD_SGD = 1,             /* torch.optim.SGD, no momentum */
  RECEMB_UPD_ADAGRAD = 2,         /* torch.optim.Adagrad element-wise (embedding_module_gen.py:97,:137) */
  RECEMB_UPD_ROWWISE_ADAGRAD = 3, /* state[num_rows]: s += mean_d(g^2); w -= lr*g/(sqrt(s)+eps) */
  RECEMB_UPD_ADAM = 4,            /* lazy (touched-rows) Adam, L2 weight decay folded into g */
  RECEMB_UPD_ADAMW = 5            /* lazy AdamW (models/lthm/sequence/wrapper.py:265), decoupled decay */
};

/* Where a lookup's row lives.  NULL / all-zero = one unsharded table.
 *   table batching: T equal-shaped tables stacked as [T * rows, dim]; lookup i uses table
 *     t = i / ids_per_table (wrapped modulo num_tables when num_tables > 0) and reads row
 *     t * rows + transform(ids[i]) -- one launch for all tables.
 *   row-wise sharding: global row r lives on rank r % shard_world at local row r / shard_world;
 *     `table` is this rank's shard, num_rows stays the GLOBAL count used for hashing, lookups
 *     owned by other ranks are skipped (forward) / dropped (plan).  With both, the stacked local
 *     table is [T * local_rows, dim], local_rows = ceil((num_rows - shard_rank) / shard_world).
 *   flip_len = L > 0 (sequence gather / k-shift): the lookups form sequences of L; output row i is
 *     written at the mirrored position inside its sequence, (i / L) * L + (L - 1 - i % L), and the
 *     plan maps every slot to that mirrored gradient row -- Encoder.flip_all
 *     (models/lthm/sequence/encoder.py:52-54, :60-61: right-padded -> left-padded) folded into the
 *     gather's addressing instead of a torch.flip copy of [B, L, D].
 *   window_keep != NULL (sequence gather / k-shift / their plan): the lookups form sequences of seq_len = L;
 *     K = *window_keep (a DEVICE int32, 0 <= K <= L, typically written by recemb_sequence_window on the same
 *     stream) positions of every sequence are kept -- the first K (window_side 0) or the last K
 *     (window_side 1) -- the others are neither read nor written, and the kept ones are stored compactly:
 *     output row (i / L) * K + q, q = position inside the window (mirrored when flip_len == L).  `out`
 *     must hold n rows (K = L); the first (n / L) * K are written.  The plan drops the slots outside the
 *     window and maps the others to the compact gradient rows.  This is QueryTower's batch-wide trim
 *     (models/lthm/sequence/query_tower.py:73-86) applied BEFORE the rows are moved instead of after.
 *   partition = 1 (peer exchange, table-batched, shard_world > 1): TABLE-wise instead of row-wise -- table t
 *     lives whole on rank t % shard_world as local table t / shard_world (recemb_peer_bucket_push routes every
 *     lookup of table t to that rank, local row = (t / W) * num_rows + row).
 *   out_features = F > 0 (table-batched pooled lookups and their plan; b = ids_per_table / bag_size bags per
 *     table): the pooled row of bag g is written at row (g % b) * F + g / b + out_feature_offset of `out`,
 *     i.e. feature-interleaved into a [b, F, dim] tensor -- the input layout of the ranker's dot interaction
 *     (recemb_dot_interaction_fwd), so no concatenation copy sits between the lookup and the interaction;
 *     the plan maps every slot to the same row of the [b, F, dim] gradient the interaction's backward
 *     produces. */
typedef struct recemb_layout {
  int64_t ids_per_table;
  int32_t num_tables;
  int32_t shard_world;
  int32_t shard_rank;
  int32_t flip_len;
  int32_t seq_len;
  int32_t window_side;
  const int32_t* window_keep;
  int32_t out_features;
  int32_t out_feature_offset;
  int32_t partition;
} recemb_layout;

typedef struct recemb_optim_params {
  float lr;           /* already includes lr_decay / schedule: the host passes clr */
  float eps;
  float weight_decay;
  float beta1;
  float beta2;
  float bias_correction1; /* 1 - beta1^step (host-computed) */
  float bias_correction2; /* 1 - beta2^step */
  float grad_div;         /* > 0: every gradient element is scaled by 1 / grad_div before the
                             reduction -- the backward of KShiftEmbedding's x / sqrt(num_shifts)
                             (commons/layers.py:170) folded into the segmented reduction instead of
                             a separate pass that materialises dx (one multiply by the rounded
                             reciprocal: <= 1 ulp from the reference's division, exact when
                             sqrt(num_shifts) is a power of two, e.g. k = 4, 16); 0 = off */
} recemb_optim_params;

/* ---- library info -------------------------------------------------------- */
RECEMB_API int recemb_abi_version(void);
RECEMB_API const char* recemb_last_error(void);
/* number of kernels this library has launched since load (process-wide, monotonic) */
RECEMB_API uint64_t recemb_launch_count(void);

/* ---- index hashing (a2) -------------------------------------------------- */
/* rows_out[i] = transform(ids[i]).  Replaces KShiftEmbedding.get_row_idx
 * (commons/layers.py:174-185) and the torch.remainder at :57 / :116-118. */
RECEMB_API int recemb_row_index(const int64_t* ids, int64_t n, int hash_mode, int64_t num_rows,
                     int64_t hash_arg, int64_t* rows_out, int device, recemb_stream_t stream);

/* ---- forward: sequence gather (a1, a4, a6) -------------------------------- */
/* out[i, :] = epilogue(table[transform(ids[i]), :]).  Replaces F.embedding at
 * commons/layers.py:58 (FlatEmbedding.forward).  If zero_pad != 0, positions
 * with ids[i] == pad_id are written as zeros without reading the table (the
 * fused form of ProductTower's `ids == 0` mask, product_tower.py:47-59).
 * If table2 != NULL the row table2[transform2(ids[i])] is added before the
 * epilogue (QREmbedding: emb_q(q) + emb_r(r), commons/layers.py:115-123):
 * hash_mode applies to `table` and hash_mode2 to `table2`, both with hash_arg.
 * inv_norm_out (optional, fp32 [n]) receives 1/max(||x||,1e-12) for the L2NORM backward.
 * `layout` (optional) selects table batching / sharding, see recemb_layout. */
RECEMB_API int recemb_gather_fwd(const void* table, int64_t num_rows, const void* table2, int64_t num_rows2,
                      int32_t dim, int dtype, const int64_t* ids, int64_t n, const recemb_layout* layout,
                      int hash_mode, int hash_mode2, int64_t hash_arg, int epilogue, int zero_pad, int64_t pad_id,
                      void* out, float* inv_norm_out, int device, recemb_stream_t stream);

/* ---- forward: fused k-shift bag (a3) -------------------------------------- */
/* out[i, :] = epilogue( sum_{c=0}^{k-1} table[rotl_floormod(ids[i], c), :] ), fp32
 * accumulation in the order c = 0, 1, ... (bit-exact pre-epilogue vs. the
 * reference's chain of adds).  Replaces KShiftEmbedding.forward
 * (commons/layers.py:152-172): 2k-1 launches -> 1.  inv_norm_out (optional,
 * fp32 [n]) receives 1/max(||x||,eps) for the L2NORM backward. */
RECEMB_API int recemb_kshift_fwd(const void* table, int64_t num_rows, int32_t dim, int dtype,
                      const int64_t* ids, int64_t n, int32_t num_shifts, int epilogue, int32_t flip_len,
                      void* out, float* inv_norm_out, int device, recemb_stream_t stream);
/* Same with a recemb_layout (flip_len and the sequence window; one unsharded table). */
RECEMB_API int recemb_kshift_fwd_layout(const void* table, int64_t num_rows, int32_t dim, int dtype,
                             const int64_t* ids, int64_t n, int32_t num_shifts, int epilogue,
                             const recemb_layout* layout, void* out, float* inv_norm_out, int device,
                             recemb_stream_t stream);

/* ---- several small lookups summed into one row (QueryTower's input, f1) ----- */
/* out[i, :] = mask[i] ? masked_row : base[i, :] + sum_k table_k[transform_k(ids_k[i]), :], the terms added in
 * order k = 0, 1, ... in fp32 (bf16: rounded after every add, as a chain of torch adds does) -- bit-exact
 * models/lthm/sequence/query_tower.py:89-104:
 *   x = inp_proj(input) + action_embedding(labels) + hod(ts) + how(ts) + dow(ts);  x = where(mask, pad, x)
 * as ONE pass (one read of `base`, one write of `out`; the 4 / 24 / 168 / 7-row tables stay in L1) instead of
 * four gathers, four adds and a where over [B, L, D].  base may be NULL (zeros); mask (uint8 [n]) optional;
 * up to 8 terms; rows of at most 512 bytes.  terms_host is a HOST array. */
typedef struct recemb_gather_term {
  const void* table;   /* [num_rows, dim] in `dtype` */
  int64_t num_rows;
  const int64_t* ids;  /* [n] */
  int hash_mode;       /* recemb_hash, e.g. RECEMB_HASH_FLOORMOD / RECEMB_HASH_DIV_FLOORMOD */
  int64_t hash_arg;    /* DIV_FLOORMOD: the divisor (PatternFromTimelocal.div) */
} recemb_gather_term;
RECEMB_API int recemb_multi_gather_add_fwd(const void* base, const recemb_gather_term* terms_host, int32_t num_terms,
                                int64_t n, int32_t dim, int dtype, const uint8_t* mask, const void* masked_row,
                                void* out, int device, recemb_stream_t stream);

/* ---- sequence window: the batch-wide trim (a7) ---------------------------- */
/* QueryTower.forward's trim (models/lthm/sequence/query_tower.py:73-79) as one kernel, result left on the
 * device: a column is all-pad when every row of the batch is padded there -- data[b, l] == pad_id for
 * kind 0 (int64 ids [batch, seq_len]) or data[b, l] != 0 for kind 1 (uint8 / bool mask, QueryTower's
 * mask_inp).  trim = seq_len - min_keep when MORE than that many columns are all-pad, else the length of
 * the all-pad run at the padded end: the START of the sequences for window_side 1 (left-padded, what
 * QueryTower sees after Encoder.flip_all), the END for window_side 0 (right-padded, before the flip).
 * out[0] = keep = seq_len - trim, out[1] = trim (device int32[2]); pass out as recemb_layout.window_keep
 * with the same window_side.  seq_len <= 8192; workspace as recemb_sequence_window_workspace_bytes. */
RECEMB_API size_t recemb_sequence_window_workspace_bytes(int32_t seq_len);
RECEMB_API int recemb_sequence_window(const void* data, int kind, int64_t batch, int32_t seq_len, int64_t pad_id,
                           int32_t min_keep, int window_side, void* workspace, size_t workspace_bytes,
                           int32_t* out, int device, recemb_stream_t stream);

/* ---- forward: pooled multi-hot bag (a5, a11) ------------------------------ */
/* ids [num_bags, bag_size]; out[b, :] = pool_{p in window(b)} w[b,p] * table[transform(ids[b,p]), :]
 * window(b): all p with (lengths == NULL || p < lengths[b]) and, if last_n > 0,
 * p >= lengths[b] - last_n; slots with ids == pad_id are skipped when zero_pad != 0
 * (EmbeddingBag padding_idx semantics).  fp32 accumulation in slot order (the CPU
 * EmbeddingBag sum order).  MEAN divides by the number of pooled slots (>=1).
 * Replaces nn.EmbeddingBag(mode='sum') at commons/transformers/layers.py:457,:469.
 * per_slot_weight (optional fp32 [num_bags, bag_size]) = per_sample_weights.
 * `layout` (optional): table batching (slot index = bag * bag_size + p) and / or row-wise
 * sharding -- with sharding `out` is this owner's PARTIAL pool (sum the partials of all owners,
 * in owner order, for the full result: recemb_sum_partials). */
RECEMB_API int recemb_pool_fwd(const void* table, int64_t num_rows, int32_t dim, int dtype,
                    const int64_t* ids, int64_t num_bags, int32_t bag_size, const int32_t* lengths,
                    int32_t last_n, const float* per_slot_weight, int hash_mode, int64_t hash_arg,
                    int pool_mode, int zero_pad, int64_t pad_id, const recemb_layout* layout,
                    void* out, int device, recemb_stream_t stream);

/* ---- backward: plan = sort-based dedup (a8, K5/K6) ------------------------ */
/* A plan turns the lookup slots of one forward call into (row, slot) pairs
 * sorted by row (stable in slot): the input of the segmented reduction.
 * Slot s in [0, n_ids * slots_per_id):
 *   slots_per_id == 1 : row = transform(ids[s])                       (gather / pooled bag)
 *   slots_per_id == k : row = rotl_floormod(ids[s / k], s % k)        (k-shift; hash_mode must be ROTL_FLOORMOD)
 * A slot is dropped from the plan when: zero_pad && id == pad_id; row == pad_row
 * (nn.Embedding padding_idx: that row never receives gradient, commons/layers.py:51);
 * bag_size > 0 and the slot is outside its bag's window (lengths / last_n as in
 * recemb_pool_fwd).
 * With a `layout` the sorted keys are rows of the stacked / local table (t * local_rows + local
 * row); one sort covers all tables.  The plan lives in caller memory of
 * recemb_bwd_plan_bytes(n_slots, total_rows) with total_rows = recemb_layout_total_rows(...)
 * (also the num_rows to pass to recemb_bwd_apply). */
RECEMB_API int64_t recemb_layout_total_rows(int64_t num_rows, const recemb_layout* layout, int64_t n_ids);

RECEMB_API size_t recemb_bwd_plan_bytes(int64_t n_slots, int64_t num_rows);

RECEMB_API int recemb_bwd_plan(const int64_t* ids, int64_t n_ids, const recemb_layout* layout,
                    int32_t slots_per_id, int hash_mode, int64_t num_rows, int64_t hash_arg, int zero_pad, int64_t pad_id,
                    int64_t pad_row, int32_t bag_size, const int32_t* lengths, int32_t last_n,
                    void* plan, size_t plan_bytes, int device, recemb_stream_t stream);

/* Fills the plan's two device counters {valid slots, distinct rows} (int64[2] at the start of
 * the plan buffer).  Separate from recemb_bwd_plan because the update path does not need them. */
RECEMB_API int recemb_plan_count(void* plan, size_t plan_bytes, int64_t n_slots, int64_t num_rows,
                      int device, recemb_stream_t stream);

/* Device-side views into a built plan (valid until the plan memory is reused). */
RECEMB_API int recemb_plan_views(const void* plan, size_t plan_bytes, const uint32_t** sorted_rows,
                      const uint32_t** sorted_slots, const int64_t** counters /* [0]=n_valid, [1]=n_unique */,
                      int64_t* n_slots_host);

/* ---- backward: segmented reduction + update (a8, a9, K5-K7) ---------------- */
/* n_slots = number of (row, slot) entries in the plan.
 * For every distinct row r in the plan: g = sum over its slots s (ascending) of
 *     slot_weight[s] * grad_row_scale[s / slots_per_grad_row] * grad[s / slots_per_grad_row, :]
 * (both scale arrays optional, fp32; without them and with hp_host->grad_div > 0 the term is
 * grad[...] / grad_div) accumulated in fp32, then `update` is
 * applied to row r of `table` (and state1/state2).  Deterministic: fixed
 * chunking, no atomics.  Workspace: recemb_bwd_apply_workspace_bytes().
 *   DENSE_GRAD        table = grad_weight (dtype `dtype`), states unused
 *   SGD               states unused
 *   ADAGRAD           state1 = sum of squares, fp32 [num_rows, dim]
 *   ROWWISE_ADAGRAD   state1 = fp32 [num_rows]
 *   ADAM / ADAMW      state1 = exp_avg, state2 = exp_avg_sq, fp32 [num_rows, dim]
 * Replaces autograd's embedding_dense_backward + torch.optim.*.step
 * (embedding_module_gen.py:113-114, :152-153). */
RECEMB_API size_t recemb_bwd_apply_workspace_bytes(int64_t n_slots, int32_t dim);

RECEMB_API int recemb_bwd_apply(const void* plan, size_t plan_bytes, int64_t n_slots, const void* grad, int grad_dtype,
                     int64_t grad_rows, int32_t dim, int32_t slots_per_grad_row,
                     const float* slot_weight, const float* grad_row_scale, int update,
                     void* table, int dtype, int64_t num_rows, void* state1, void* state2,
                     const recemb_optim_params* hp_host, void* workspace, size_t workspace_bytes,
                     int device, recemb_stream_t stream);

/* recemb_bwd_apply with a device-side guard: when *skip_if_nonzero (device memory, may be NULL) is
 * non-zero at kernel start, every kernel of the call returns without reading or writing the table
 * or the optimizer state.  The peer exchange passes its arena's status word: a step whose inbox
 * overflowed or whose barrier timed out leaves the shard untouched instead of applying an
 * incomplete gradient; the host learns about it at its next (asynchronous) status check. */
RECEMB_API int recemb_bwd_apply_guarded(const void* plan, size_t plan_bytes, int64_t n_slots, const void* grad,
                             int grad_dtype, int64_t grad_rows, int32_t dim, int32_t slots_per_grad_row,
                             const float* slot_weight, const float* grad_row_scale, int update,
                             void* table, int dtype, int64_t num_rows, void* state1, void* state2,
                             const recemb_optim_params* hp_host, void* workspace, size_t workspace_bytes,
                             const uint32_t* skip_if_nonzero, int device, recemb_stream_t stream);

/* Measurement hook: the next recemb_bwd_apply on this thread records the two CUDA events
 * (cudaEvent_t) around its level-0 segmented-reduction launch, then disarms the hook. */
RECEMB_API int recemb_time_next_apply(void* start_event, void* stop_event);

/* ---- backward of the k-shift / pooled epilogues ---------------------------- */
/* dx[i,:] for y = epilogue(x): L2NORM: (g - y (y.g)) * inv_norm[i]; RSQRT_K: g / sqrt(k).
 * (autograd of commons/layers.py:167-170).  dx is fp32 [n, dim]. */
RECEMB_API int recemb_epilogue_bwd(const void* grad_out, const void* out, int dtype, const float* inv_norm,
                        int64_t n, int32_t dim, int epilogue, int32_t num_shifts, float* dx,
                        int device, recemb_stream_t stream);

/* ---- row-wise sharding: requester-side reduction (a12) ----------------------- */
/* out[r, :] = row_scale[r] * sum_{s=0}^{world-1} parts[s, r, :] (fp32 accumulation in owner
 * order; row_scale optional, e.g. 1/count for MEAN).  parts = the all-to-all receive buffer
 * [world, rows, dim] of per-owner partial pools produced by recemb_pool_fwd(shard_world>1). */
RECEMB_API int recemb_sum_partials(const void* parts, int32_t world, int64_t rows, int32_t dim, int dtype,
                        const float* row_scale, void* out, int device, recemb_stream_t stream);

/* ---- row-wise sharding: routed exchange (a12) -------------------------------- */
/* Sender side.  Buckets this rank's lookup slots by owning rank (stable counting sort with
 * shard_world bins): entries_out[n_ids] receives, bucket after bucket, one int64 per kept slot
 *     (local row in the owner's stacked shard) << 32 | (shard_rank * bags_total + bag)
 * and counts_out[shard_world] the bucket sizes (device int64).  Slots outside their bag window
 * (lengths / last_n) or equal to pad_id (zero_pad) are dropped here and never travel.  Inside
 * a bucket the entries keep slot order (sorted by bag).  layout: shard_world / shard_rank as in
 * recemb_layout (shard_rank = THIS rank), ids_per_table / num_tables for stacked tables. */
RECEMB_API size_t recemb_shard_bucket_workspace_bytes(int64_t n_slots, int32_t world);
RECEMB_API int recemb_shard_bucket(const int64_t* ids, int64_t n_ids, const recemb_layout* layout, int hash_mode,
                        int64_t num_rows, int64_t hash_arg, int zero_pad, int64_t pad_id, int32_t bag_size,
                        const int32_t* lengths, int32_t last_n, int64_t bags_total, int64_t* entries_out,
                        int64_t* counts_out, void* workspace, size_t workspace_bytes, int device,
                        recemb_stream_t stream);
/* Owner side.  entries = the buckets received from all ranks, concatenated in rank order.  Every
 * run of equal low word (sender, bag) is pooled (fp32 accumulation in entry order) from `table`
 * (this rank's stacked shard) into out[low word, :]; rows of bags with no entry are not written
 * (the caller zero-fills out [shard_world * bags_total, dim]). */
RECEMB_API int recemb_pool_entries(const void* table, int32_t dim, int dtype, const int64_t* entries, int64_t n,
                        void* out, int device, recemb_stream_t stream);
/* Owner side.  Builds a backward plan straight from the received entries: key = local row,
 * slot = low word = row of the all-gathered gradient [shard_world * bags_total, dim]
 * (use recemb_bwd_apply with slots_per_grad_row = 1).  plan as in recemb_bwd_plan_bytes(n, total_rows). */
RECEMB_API int recemb_bwd_plan_entries(const int64_t* entries, int64_t n, int64_t total_rows, void* plan,
                            size_t plan_bytes, int device, recemb_stream_t stream);

/* ---- row-wise sharding over peer memory (a12; NVLink 5 / NVSwitch P2P) -------- */
/* The "peer" exchange replaces the NCCL collectives of the routed exchange by loads and stores
 * on peer-mapped device memory inside the lookup kernels themselves (no host synchronisation,
 * fixed shapes, CUDA-graph capturable):
 *   forward   ONE kernel: every lookup reads its row straight from the owning rank's shard over
 *             NVLink (ld.global on the mapped peer pointer) and is pooled on the requester in
 *             slot order -- bit-identical to the unsharded pooled bag.
 *   backward  bucket-by-owner (as recemb_shard_bucket) whose scatter stores each entry directly
 *             into the owner's inbox, a push all-gather of the pooled gradients, one device-side
 *             barrier over flags in peer memory, then the owner's sort + segmented reduction +
 *             fused update of ITS rows; a second barrier closes the step (peers must not read
 *             rows that are still being updated).
 * Memory is shared with CUDA IPC: a rank exports the allocation behind a device pointer
 * (recemb_peer_export), the handles travel over the caller's host channel (torch.distributed),
 * every other rank maps them (recemb_peer_open).  The library keeps no global state: the mapped
 * pointers live in a caller-owned recemb_peer_group. */
#define RECEMB_PEER_HANDLE_BYTES 64
#define RECEMB_MAX_PEERS 16

typedef struct recemb_peer_group {
  int32_t world;
  int32_t rank;
  void* arena[RECEMB_MAX_PEERS];       /* exchange arena of every rank as mapped HERE (arena[rank] is local) */
  const void* table[RECEMB_MAX_PEERS]; /* stacked table shard of every rank as mapped HERE */
} recemb_peer_group;

/* Byte offsets inside one rank's exchange arena (identical on every rank). */
typedef struct recemb_peer_arena {
  int64_t bytes;       /* total size; the caller allocates it zero-filled, 256-byte aligned */
  int64_t off_flags;   /* uint64 [channels][RECEMB_MAX_PEERS]  barrier flags, slot s is written by rank s */
  int64_t off_epoch;   /* uint64 [channels] this rank's barrier counts */
  int64_t off_status;  /* uint32           sticky bits: 1 = inbox overflow (entries dropped), 2 = barrier timeout */
  int64_t off_counts;  /* int64 [world]    entries rank s pushed into my inbox this step */
  int64_t off_inbox;   /* int64 [world][cap] entries, region s written by rank s */
  int64_t off_grads;   /* [world][bags_total][dim] pooled gradients, slice s written by rank s */
  int64_t cap;         /* inbox capacity per sender (entries) */
  int64_t bags_total;
  int64_t off_parts;   /* [world][bags_total][dim] partial pools of MY bags, slice o written completely by
                          owner o (push forward) */
  int64_t off_gate;    /* gate region of the fused gradient push (recemb_peer_bwd_apply_fused): step counter,
                          per-table completion counts, per-table flags [64][RECEMB_MAX_PEERS] */
} recemb_peer_arena;

/* handle_out identifies the whole allocation that contains ptr; *offset_out = ptr - its base. */
RECEMB_API int recemb_peer_export(const void* ptr, uint8_t handle_out[RECEMB_PEER_HANDLE_BYTES],
                       int64_t* offset_out, int64_t* alloc_bytes_out, int device);
/* Maps an exported allocation into this process (peer access is enabled on demand); *base_out is
 * the mapped base, add the exporter's offset.  A handle must be opened once per process. */
RECEMB_API int recemb_peer_open(const uint8_t handle[RECEMB_PEER_HANDLE_BYTES], void** base_out, int device);
RECEMB_API int recemb_peer_close(void* base, int device);

RECEMB_API int recemb_peer_arena_layout(int32_t world, int64_t cap, int64_t bags_total, int32_t dim, int dtype,
                             recemb_peer_arena* out);

/* Device-side barrier over all ranks of the group (one tiny kernel on `stream`): everything the
 * ranks enqueued before it on that stream -- including their stores into peer memory -- is
 * visible to everything enqueued after it on any rank.  `channel` (0 .. RECEMB_PEER_CHANNELS-1)
 * selects an independent set of flags, so that two streams of a rank can each run their own
 * barrier sequence concurrently; every rank must call a channel the same number of times.
 * A rank that waits longer than RECEMB_PEER_BARRIER_TIMEOUT_S (environment, default 600 s) sets status
 * bit 2 in its own arena and continues (no GPU hang); see recemb_bwd_apply_guarded. */
#define RECEMB_PEER_CHANNELS 4
RECEMB_API int recemb_peer_barrier(const recemb_peer_group* group, const recemb_peer_arena* arena, int channel,
                        int device, recemb_stream_t stream);
/* The two halves of the barrier as separate launches (pipelined steps: producer and consumer on different
 * streams).  recemb_peer_signal publishes this rank's next epoch of `channel` to every rank and never
 * blocks: enqueue it right after the kernel whose peer stores it announces.  recemb_peer_wait blocks `stream`
 * until every rank has published the epoch this rank is at: enqueue it right before the consumer, AFTER
 * this rank's own signal of the same round in stream / event order.  On a channel, rounds of
 * (signal, wait) and full barriers may be mixed as long as every rank issues the same sequence. */
RECEMB_API int recemb_peer_signal(const recemb_peer_group* group, const recemb_peer_arena* arena, int channel,
                       int device, recemb_stream_t stream);
RECEMB_API int recemb_peer_wait(const recemb_peer_group* group, const recemb_peer_arena* arena, int channel,
                     int device, recemb_stream_t stream);

/* Forward.  As recemb_pool_fwd on the unsharded table of num_rows (GLOBAL) rows per table, except
 * that global row r is read from group->table[r % world] at local row r / world (+ the table
 * offset t * local_rows(owner) for stacked tables: layout->ids_per_table / num_tables).  `out` is
 * the complete pool of this rank's bags. */
RECEMB_API int recemb_peer_pool_fwd(const recemb_peer_group* group, int64_t num_rows, int32_t dim, int dtype,
                         const int64_t* ids, int64_t num_bags, int32_t bag_size, const int32_t* lengths,
                         int32_t last_n, const float* per_slot_weight, int hash_mode, int64_t hash_arg,
                         int pool_mode, int zero_pad, int64_t pad_id, const recemb_layout* layout,
                         void* out, int device, recemb_stream_t stream);

/* Forward, "push" variant (fewer NVLink bytes than the pull: one partial row per (bag, owner) pair
 * instead of one row per lookup).  Owner side, after recemb_peer_bucket_push + barrier: every run
 * of equal (sender, bag) in MY inbox is pooled from MY shard (fp32 accumulation in slot order) and
 * the partial row is stored straight into the sender's arena at parts[rank][bag] over NVLink.
 * Pairs without an entry get a zero row from the same kernel (the run that follows a gap of bag numbers in
 * a sender's region fills it; the last run fills the tail), so every row of parts[rank] is written exactly
 * once per step and the requester needs no zero fill; it sums the world slices after the barrier
 * (recemb_sum_partials, fixed owner order). */
RECEMB_API int recemb_peer_pool_push(const recemb_peer_group* group, const recemb_peer_arena* arena, int32_t dim,
                          int dtype, int device, recemb_stream_t stream);
/* Table-wise partitioning (recemb_layout.partition = 1): every bag has exactly one owner, so the partial row
 * IS the pooled row; zero rows are written only for empty bags of the tables this rank owns (bag g belongs to
 * table g / bags_per_table), and the requester picks parts[t % world][bags of table t] instead of summing. */
RECEMB_API int recemb_peer_pool_push_tablewise(const recemb_peer_group* group, const recemb_peer_arena* arena,
                                    int32_t dim, int dtype, int64_t bags_per_table, int device,
                                    recemb_stream_t stream);

/* Backward, sender side.  recemb_shard_bucket whose entries land in the owners' inboxes: entry
 * k of my bucket for owner o is stored at inbox(o)[rank][k], the bucket size at counts(o)[rank].
 * Entries beyond arena->cap are dropped and status bit 1 is set on this rank AND on the owner whose
 * inbox is now incomplete (so that its guarded update skips the step).  workspace as
 * recemb_shard_bucket_workspace_bytes. */
RECEMB_API int recemb_peer_bucket_push(const recemb_peer_group* group, const recemb_peer_arena* arena,
                            const int64_t* ids, int64_t n_ids, const recemb_layout* layout, int hash_mode,
                            int64_t num_rows, int64_t hash_arg, int zero_pad, int64_t pad_id, int32_t bag_size,
                            const int32_t* lengths, int32_t last_n, void* workspace, size_t workspace_bytes,
                            int device, recemb_stream_t stream);
/* Sequence mode (every lookup is its own output / gradient row; forward = recemb_peer_pool_fwd with
 * bag_size 1, i.e. each row pulled from its owner).  Backward, sender side, two halves:
 *   recemb_peer_bucket_push_rows   (needs only the ids: issue it in the forward's shadow) as
 *     recemb_peer_bucket_push with bag_size 1, except that entry k of my bucket for owner o names the slot
 *     rank * cap + k of the owner's gradient buffer, and dest_out[i] (int64 [n_ids]) = (o << 32 | k) for
 *     lookup i, -1 if it was dropped (pad id / out-of-range identity id / inbox overflow);
 *   recemb_peer_rows_scatter_push  stores gradient row i (rows [n, dim]) into that slot of its owner's
 *     buffer over NVLink: every row crosses the link at most once (an all-gather would send it W - 1 times).
 * The arena must be laid out with bags_total == cap (gradient buffer [world][cap][dim]); the owner then
 * continues as in the pooled case: barrier, recemb_peer_plan, recemb_bwd_apply(grad = arena + off_grads). */
RECEMB_API int recemb_peer_bucket_push_rows(const recemb_peer_group* group, const recemb_peer_arena* arena,
                                 const int64_t* ids, int64_t n_ids, const recemb_layout* layout, int hash_mode,
                                 int64_t num_rows, int64_t hash_arg, int zero_pad, int64_t pad_id,
                                 int64_t* dest_out, void* workspace, size_t workspace_bytes, int device,
                                 recemb_stream_t stream);
RECEMB_API int recemb_peer_rows_scatter_push(const recemb_peer_group* group, const recemb_peer_arena* arena,
                                  const void* rows, int64_t n, int32_t dim, int dtype, const int64_t* dest,
                                  int device, recemb_stream_t stream);
/* Backward, sender side.  Push all-gather: src (bytes, 16-byte multiple) is stored at
 * arena(p) + dst_offset + rank * bytes of EVERY rank p (its own included). */
RECEMB_API int recemb_peer_allgather_push(const recemb_peer_group* group, const void* src, int64_t bytes,
                               int64_t dst_offset, int device, recemb_stream_t stream);
/* Backward, owner side (after the barrier).  Plan over my inbox: world * cap (row, gradient row)
 * pairs, unused inbox positions carry the sentinel key total_rows and sort last.  Use with
 * recemb_bwd_apply(n_slots = world * cap, grad = my arena + off_grads, slots_per_grad_row = 1);
 * plan as in recemb_bwd_plan_bytes(world * cap, total_rows). */
RECEMB_API int recemb_peer_plan(const recemb_peer_group* group, const recemb_peer_arena* arena, int64_t total_rows,
                     void* plan, size_t plan_bytes, int device, recemb_stream_t stream);

/* Backward, sender AND owner side in ONE launch (the pooled push / pull exchange, update kinds whose
 * per-row state is a scalar: row-wise Adagrad, SGD; rows of 256 or 512 bytes, gradients in the table dtype).
 * Replaces recemb_peer_allgather_push + recemb_peer_barrier + recemb_bwd_apply_guarded: the first
 * `push_ctas` CTAs of the level-0 kernel store this rank's pooled gradients my_grad
 * [tables][bags_per_table][dim] into every rank's gradient buffer table by table and publish one flag per
 * table; the remaining CTAs are the segmented reduction + fused update over `plan` (recemb_peer_plan), and a
 * chunk of sorted entries waits only for the flags of the last table it touches (keys are table-major:
 * rows_per_table local rows per table) -- the NVLink transfer of table t + 1 overlaps the HBM-bound update
 * of table t inside one kernel.  The update is skipped when the arena's status word is non-zero; a gate that
 * times out (RECEMB_PEER_BARRIER_TIMEOUT_S) sets status bit 2 and skips its chunk.  Every rank must call it
 * once per step.  workspace as recemb_bwd_apply_workspace_bytes(world * cap, dim). */
RECEMB_API int recemb_peer_bwd_apply_fused(const recemb_peer_group* group, const recemb_peer_arena* arena,
                                const void* plan, size_t plan_bytes, const void* my_grad, int32_t tables,
                                int64_t bags_per_table, int32_t dim, int dtype, int update, void* table,
                                int64_t total_rows, int64_t rows_per_table, void* state1,
                                const recemb_optim_params* hp, void* workspace, size_t workspace_bytes,
                                int32_t push_ctas, int device, recemb_stream_t stream);
/* Same for table-wise partitioning: the gradients of table t are pushed to rank t % world ONLY (every gradient
 * row crosses NVLink at most once) and gate its local table t / world; `table` holds
 * ceil((tables - rank) / world) whole tables of rows_per_table rows. */
RECEMB_API int recemb_peer_bwd_apply_fused_tablewise(const recemb_peer_group* group, const recemb_peer_arena* arena,
                                          const void* plan, size_t plan_bytes, const void* my_grad, int32_t tables,
                                          int64_t bags_per_table, int32_t dim, int dtype, int update, void* table,
                                          int64_t total_rows, int64_t rows_per_table, void* state1,
                                          const recemb_optim_params* hp, void* workspace, size_t workspace_bytes,
                                          int32_t push_ctas, int device, recemb_stream_t stream);

/* ---- ranker pairwise dot interaction (a11) --------------------------------- */
/* feats bf16 [batch, num_feats, dim] -> out bf16 [batch, num_feats*(num_feats-1)/2]:
 * the strictly-lower triangle of feats[b] @ feats[b]^T, fp32 accumulation on the
 * tcgen05 tensor cores (num_feats <= 32, dim % 64 == 0).  No reference code
 * exists (models/ranker/fdlrm/ is empty): canonical DLRM interaction. */
RECEMB_API int recemb_dot_interaction_fwd(const void* feats, int64_t batch, int32_t num_feats, int32_t dim,
                               void* out, int device, recemb_stream_t stream);
/* grad_feats[b] = (G + G^T) @ feats[b] with G the lower-triangular unpack of grad_out[b]. */
RECEMB_API int recemb_dot_interaction_bwd(const void* feats, const void* grad_out, int64_t batch,
                               int32_t num_feats, int32_t dim, void* grad_feats, int device,
                               recemb_stream_t stream);

/* ---- id production (feeder of the path) ------------------------ */
/* ids_out[i] = XXH64(bytes[offsets[i] : offsets[i+1]], seed) - 2^63 as signed int64: bit-exact
 * hash_string_to_long (commons/feature_utils.py:40-46; xxhash==3.5.0).  to_lower lower-cases ASCII
 * A-Z on the fly (value_to_lower; non-ASCII text must be lower-cased by the caller).  seed is
 * hash_feature_name_to_int(feature) (:36-37, an xxh32 of the feature name, computed on the host). */
RECEMB_API int recemb_xxh64_ids(const uint8_t* bytes, const int64_t* offsets, int64_t n, uint64_t seed,
                     int to_lower, int64_t* ids_out, int device, recemb_stream_t stream);
/* out[r, :] = the first history_length elements of values[offsets[r] : offsets[r+1]] that differ
 * from remove_ids[r] (remove_ids optional), right-padded with pad_token: pad_array (:21-25) +
 * handle_categorical_history_feature (:149-179, remove_history_id_from_history). */
RECEMB_API int recemb_pad_histories(const int64_t* values, const int64_t* offsets, const int64_t* remove_ids,
                         int64_t rows, int32_t history_length, int64_t pad_token, int64_t* out, int device,
                         recemb_stream_t stream);

/* ---- streaming logQ correction ------------------------------- */
/* The D = 1 cousin of the gather / scatter: num_tables cascaded bucket tables b_m [num_buckets]
 * fp32 with hash h_m(id) = floor_mod(id + hash_offsets[m], num_buckets) (int64 wrap-around add).
 * b_tables_host / a_tables_host / hash_offsets_host are HOST arrays of num_tables device
 * pointers / offsets (num_tables <= 16; lthm.yaml: 7 offsets x 2^24 buckets).
 *   fwd     out[i] = min_m( -log(b_m[h_m(ids[i])]) )
 *           StreamingLogQCorrectionModule.forward (commons/layers.py:202-204) under
 *           CascadedStreamingLogQCorrectionModule.forward (:224-232, torch.minimum over modules) */
RECEMB_API int recemb_logq_fwd(const float* const* b_tables_host, int32_t num_tables,
                    const int64_t* hash_offsets_host, int64_t num_buckets, const int64_t* ids, int64_t n,
                    float* out, int device, recemb_stream_t stream);
/*   update  for every id (skip_mask[i] == 0): b_m[h] = (1 - alpha) * b_m[h] + alpha * (batch_idx - a_m[h]),
 *           a_m[h] = batch_idx -- StreamingLogQCorrectionModule.train_step (commons/layers.py:210-213;
 *           its last line assigns into the float `alpha`, read here as `self.a[hash] = batch_idx`).
 *           The right-hand side is evaluated from the OLD tables for all ids before anything is
 *           written (torch's gather-then-index_put order); duplicates of a bucket write the same
 *           value.  skip_mask (optional, uint8 [n]) fuses the reference's boolean compaction
 *           `product_ids.view(-1)[mask.view(-1) == 0]` (models/lthm/sequence/wrapper.py:133).
 *           scratch: fp32 [n * num_tables]. */
RECEMB_API int recemb_logq_update(float* const* b_tables_host, float* const* a_tables_host, int32_t num_tables,
                       const int64_t* hash_offsets_host, int64_t num_buckets, const int64_t* ids, int64_t n,
                       const uint8_t* skip_mask, double alpha, int64_t batch_idx, float* scratch,
                       int device, recemb_stream_t stream);

/* ---- host-buffer entry points (end-to-end path) ---------------------------- */
/* One fused training step of a FlatEmbedding-style table with HOST ids:
 * H2D copy of ids_host (pinned or pageable) into ids_dev_scratch, forward gather
 * into out, backward plan + fused update with grad (device, fp32/bf16 [n, dim]),
 * then counters_host[0..1] = {n_valid, n_unique} are copied back (async on
 * `stream`; the caller synchronises the stream before reading them).  ids_per_table > 0
 * selects the table-batched mode of recemb_gather_fwd (num_rows per table, stacked table).
 * wait_event_after_copy (optional cudaEvent_t): `stream` waits for it after the H2D copy and
 * before the first kernel, so a caller alternating two streams can copy step s+1's ids while
 * step s still computes without letting the two steps' kernels overlap.
 * plan_stream (optional, != stream): the backward plan (hash + radix sort, a function of the ids
 * only) is built there while the forward gather runs on `stream`; the update joins both. */
RECEMB_API int recemb_flat_step_host(const int64_t* ids_host, int64_t n, int64_t ids_per_table,
                          int64_t* ids_dev_scratch,
                          void* table, int64_t num_rows, int32_t dim, int dtype, void* out,
                          const void* grad, int update, void* state1, void* state2,
                          const recemb_optim_params* hp_host, void* plan, size_t plan_bytes,
                          void* workspace, size_t workspace_bytes, int64_t* counters_host,
                          void* wait_event_after_copy, recemb_stream_t plan_stream, int device,
                          recemb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RECEMB_B200_H_ */
