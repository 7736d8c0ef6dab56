"""CPU oracle of the embedding hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (recommendations_b200/) never does.

It restates, with plain torch-CPU / numpy operations, what the reference computes
on this path.  The reference is pure Python over torch (pinned torch==2.8.0,
requirements.txt:56; torch 2.11 here): its arithmetic lives in torch's
nn.Embedding / nn.EmbeddingBag / torch.remainder / F.normalize / torch.optim, so
the restatement calls the same torch primitives on CPU (that IS the reference's
CPU path) and adds independent numpy versions of the integer hashing.

Pinning: the reference ships no tests and no golden vectors.  The
oracle is pinned instead against (i) the reference classes themselves, imported
from /root/reference in the build container by oracle/make_golden.py and
tests/test_oracle_vs_reference.py, and (ii) the fixtures that script wrote to
tests/golden/ (committed; they travel to the GPU box where /root/reference does
not exist).  The ranker dot-interaction and row-wise Adagrad have no reference
implementation at all: for those two "parity unpinned" applies (canonical DLRM /
FBGEMM definitions restated here).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

PAD_TOKEN = 0  # commons/feature_utils.py:8 CATEGORICAL_VAR_HASH_PAD_TOKEN
TWO63 = 2 ** 63  # commons/feature_utils.py:6


# ------------------------------------------------------------ index hashing ----
def row_index(ids: torch.Tensor, num_rows: int, col: int = 0) -> torch.Tensor:
    """KShiftEmbedding.get_row_idx, commons/layers.py:174-185 (col 0 == FlatEmbedding's
    torch.remainder at :57).  Signed int64: `<<` wraps, `>>` is arithmetic."""
    x = ids
    if col != 0:
        x = torch.bitwise_or(torch.bitwise_left_shift(x, col), torch.bitwise_right_shift(x, 64 - col))
    return torch.remainder(x, num_rows)


def row_index_np(ids: np.ndarray, num_rows: int, col: int = 0) -> np.ndarray:
    """Independent numpy restatement of the same hashing through Python integers."""
    out = np.empty(ids.shape, dtype=np.int64)
    flat_in, flat_out = ids.reshape(-1), out.reshape(-1)
    for i, v in enumerate(flat_in.tolist()):
        if col != 0:
            u = v & 0xFFFFFFFFFFFFFFFF
            left = (u << col) & 0xFFFFFFFFFFFFFFFF
            right = (v >> (64 - col)) & 0xFFFFFFFFFFFFFFFF  # Python >> on a signed int is arithmetic
            w = left | right
            v = w - (1 << 64) if w >= TWO63 else w
        flat_out[i] = v % num_rows  # Python % is floor-mod
    return out


def qr_indices(ids: torch.Tensor, num_embeddings: int) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """QREmbedding index math, commons/layers.py:106-108, :116-118."""
    d = int(math.sqrt(num_embeddings))
    x = torch.remainder(ids, d * d)
    q = torch.remainder(torch.div(x, d, rounding_mode="floor"), d)
    r = torch.remainder(x, d)
    return q, r, d


# ------------------------------------------------------------------ forward ----
def flat_embedding(weight: torch.Tensor, ids: torch.Tensor, normalize: bool = False,
                   padding_idx: Optional[int] = None) -> torch.Tensor:
    """FlatEmbedding.forward, commons/layers.py:56-61."""
    rows = torch.remainder(ids, weight.shape[0]).long()
    out = F.embedding(rows, weight, padding_idx=padding_idx)
    return F.normalize(out, p=2.0, dim=-1) if normalize else out


def kshift_embedding(weight: torch.Tensor, ids: torch.Tensor, num_shifts: int,
                     normalize: bool = False, sparse: bool = False) -> torch.Tensor:
    """KShiftEmbedding.forward, commons/layers.py:152-172: sum over c = 0..k-1 in that
    order, then L2-normalise or divide by sqrt(k)."""
    n_rows = weight.shape[0]
    acc = F.embedding(row_index(ids, n_rows, 0), weight, sparse=sparse)
    for c in range(1, num_shifts):
        acc = acc + F.embedding(row_index(ids, n_rows, c), weight, sparse=sparse)
    if normalize:
        return F.normalize(acc, p=2.0, dim=-1)
    return acc / math.sqrt(num_shifts)


def qr_embedding(weight_q: torch.Tensor, weight_r: torch.Tensor, ids: torch.Tensor,
                 num_embeddings: int, normalize: bool) -> torch.Tensor:
    """QREmbedding.forward, commons/layers.py:115-123 (ctor repaired with super().__init__())."""
    q, r, _ = qr_indices(ids, num_embeddings)
    out = F.embedding(q, weight_q) + F.embedding(r, weight_r)
    return F.normalize(out, p=2.0, dim=-1) if normalize else out


def cosine_bucket_indices(x: torch.Tensor, projection_mat: torch.Tensor, grid: torch.Tensor,
                          pos_offset: torch.Tensor) -> torch.Tensor:
    """CosineVectorEmbedding index half, commons/transformers/layers.py:462-468."""
    z = F.normalize(x, p=2.0, dim=-1) @ projection_mat
    buckets = torch.bucketize(z, grid)
    return buckets.contiguous().view(-1, projection_mat.shape[1]) + pos_offset.unsqueeze(0)


def embedding_bag_sum(weight: torch.Tensor, idxs: torch.Tensor,
                      per_sample_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.EmbeddingBag(mode='sum') on 2-D input, commons/transformers/layers.py:457, :469."""
    return F.embedding_bag(idxs, weight, mode="sum", per_sample_weights=per_sample_weights)


def pooled_bag(weight: torch.Tensor, ids: torch.Tensor, *, lengths: Optional[torch.Tensor] = None,
               last_n: int = 0, mode: str = "sum", hash_ids: bool = True, skip_pad: bool = False,
               pad_id: int = PAD_TOKEN, per_sample_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Build-defined pooled multi-hot lookup (sum / mean / last-N with masks).  Sequential
    in-order fp32 accumulation == CPU EmbeddingBag('sum') bit-for-bit."""
    m, p = ids.shape
    rows = torch.remainder(ids, weight.shape[0]) if hash_ids else ids
    pos = torch.arange(p).unsqueeze(0)
    hi = torch.full((m, 1), p) if lengths is None else lengths.long().clamp(0, p).unsqueeze(1)
    lo = (hi - last_n).clamp(min=0) if last_n > 0 else torch.zeros_like(hi)
    use = (pos >= lo) & (pos < hi)
    if skip_pad:
        use = use & (ids != pad_id)
    acc = torch.zeros((m, weight.shape[1]), dtype=torch.float32)
    w32 = weight.float()
    for j in range(p):  # slot order, fp32
        contrib = w32[rows[:, j]]
        if per_sample_weights is not None:
            contrib = contrib * per_sample_weights[:, j:j + 1].float()
        acc = torch.where(use[:, j:j + 1], acc + contrib, acc)
    if mode == "mean":
        cnt = use.sum(dim=1, keepdim=True).clamp(min=1).float()
        acc = acc / cnt
    return acc.to(weight.dtype)


# ------------------------------------------------------ backward / optimizers ----
def dense_grad(rows: torch.Tensor, grad_rows: torch.Tensor, num_rows: int,
               padding_idx: Optional[int] = None) -> torch.Tensor:
    """What autograd's embedding_dense_backward yields: grad_w[r] = sum of grad rows hitting r."""
    gw = torch.zeros((num_rows, grad_rows.shape[-1]), dtype=torch.float32)
    flat_rows = rows.reshape(-1)
    g = grad_rows.reshape(-1, grad_rows.shape[-1]).float()
    if padding_idx is not None:
        keep = flat_rows != padding_idx
        flat_rows, g = flat_rows[keep], g[keep]
    gw.index_add_(0, flat_rows, g)
    return gw


def adagrad_step(w: torch.Tensor, g: torch.Tensor, state_sum: torch.Tensor, lr: float,
                 eps: float = 1e-10, weight_decay: float = 0.0, lr_decay: float = 0.0,
                 step: int = 1, touched_only: bool = True) -> None:
    """torch.optim.Adagrad single-tensor update (embedding_module_gen.py:97, :137), in place.
    touched_only restricts it to rows with a non-zero gradient row -- identical to the dense
    update because untouched rows have g == 0 => no change (weight_decay == 0)."""
    clr = lr / (1.0 + (step - 1) * lr_decay)
    if weight_decay != 0.0:
        g = g + weight_decay * w
    state_sum.addcmul_(g, g, value=1.0)
    std = state_sum.sqrt().add_(eps)
    w.addcdiv_(g, std, value=-clr)


def rowwise_adagrad_step(w: torch.Tensor, g: torch.Tensor, touched: torch.Tensor,
                         state_row: torch.Tensor, lr: float, eps: float = 1e-10) -> None:
    """Row-wise Adagrad (SURVEY.md section 8d cfg 4; the reference has no code for it, and fbgemm_gpu -- whose
    EXACT_ROWWISE_ADAGRAD this is -- is neither a dependency of the reference nor installed here: PARITY UNPINNED,
    the published algorithm is restated).  fbgemm_gpu's optimizer code generator (rowwise_adagrad in
    codegen/genscript/optimizers.py) emits, per touched row of dimension D:
        g_avg_square        = sum_d(g_d^2) / D
        new_sum_square_grads = momentum1[row] + g_avg_square ;  momentum1[row] = new_sum_square_grads
        multiplier           = learning_rate / (sqrtf(new_sum_square_grads) + eps)
        weight_d            -= multiplier * g_d
    i.e. s_r += mean_d(g_r^2); w_r -= lr * g_r / (sqrt(s_r) + eps), touched rows only (no weight decay / max-norm
    variant here)."""
    gr = g[touched]
    s_new = state_row[touched] + (gr * gr).mean(dim=1)
    state_row[touched] = s_new
    w[touched] = w[touched] - lr * gr / (s_new.sqrt() + eps).unsqueeze(1)


def sgd_step(w: torch.Tensor, g: torch.Tensor, lr: float, weight_decay: float = 0.0) -> None:
    if weight_decay != 0.0:
        g = g + weight_decay * w
    w.add_(g, alpha=-lr)


def lazy_adam_step(w: torch.Tensor, g: torch.Tensor, touched: torch.Tensor, exp_avg: torch.Tensor,
                   exp_avg_sq: torch.Tensor, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                   weight_decay: float = 0.0, step: int = 1, decoupled: bool = False) -> None:
    """torch.optim.Adam / AdamW formula applied to the touched rows only (the semantics of
    torch.optim.SparseAdam; AdamW = models/lthm/sequence/wrapper.py:265)."""
    b1, b2 = betas
    gr, wr = g[touched], w[touched]
    if decoupled:
        wr = wr * (1.0 - lr * weight_decay)
    elif weight_decay != 0.0:
        gr = gr + weight_decay * wr
    m = b1 * exp_avg[touched] + (1 - b1) * gr
    v = b2 * exp_avg_sq[touched] + (1 - b2) * gr * gr
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    w[touched] = wr - (lr / bc1) * (m / denom)
    exp_avg[touched] = m
    exp_avg_sq[touched] = v


# ------------------------------------------------------------ dot interaction ----
def dot_interaction(feats: torch.Tensor) -> torch.Tensor:
    """Canonical DLRM pairwise interaction (no reference code: models/ranker/fdlrm/ is empty):
    strictly-lower triangle of feats @ feats^T, fp32 math on the (bf16-rounded) inputs."""
    b, f, _ = feats.shape
    z = torch.bmm(feats.float(), feats.float().transpose(1, 2))
    li, lj = torch.tril_indices(f, f, offset=-1)
    return z[:, li, lj]


# --------------------------------------------------------------- id pipeline ----
def pad_history(arr, size: int, pad_token: int = PAD_TOKEN) -> np.ndarray:
    """pad_array, commons/feature_utils.py:21-25: truncate to `size`, right-pad with 0."""
    a = np.asarray(arr, dtype=np.int64).reshape(-1)[:size]
    out = np.full((size,), pad_token, dtype=np.int64)
    out[:a.shape[0]] = a
    return out


def hash_feature_name(name: str) -> int:
    """hash_feature_name_to_int, commons/feature_utils.py:36-37."""
    import xxhash
    return xxhash.xxh32(name.lower(), 0).intdigest()


def hash_string_to_id(value, seed: int, lower: bool = False) -> int:
    """hash_string_to_long, commons/feature_utils.py:40-46: xxh64 - 2^63 -> signed int64."""
    import xxhash
    s = str(value)
    if lower:
        s = s.lower()
    return xxhash.xxh64(s, seed).intdigest() - TWO63


# ------------------------------------------------- time patterns / streaming logQ ----
def pattern_index(x: torch.Tensor, div: int, mod: int) -> torch.Tensor:
    """PatternFromTimelocal index, commons/layers.py:39-40."""
    return torch.remainder(torch.floor_divide(x.long(), div), mod)


def logq_hash(products: torch.Tensor, hash_offset: int, num_buckets: int) -> torch.Tensor:
    """StreamingLogQCorrectionModule.hash_fn, commons/layers.py:206-208 (`%` on tensors is floor-mod)."""
    return (products + hash_offset) % num_buckets


def logq_forward(b_tables, hash_offsets, products: torch.Tensor) -> torch.Tensor:
    """CascadedStreamingLogQCorrectionModule.forward, commons/layers.py:224-232 over :202-204."""
    result = None
    for b, off in zip(b_tables, hash_offsets):
        lq = -b[logq_hash(products, off, b.numel())].log().reshape(*products.shape)
        result = lq if result is None else torch.minimum(result, lq)
    return result


def logq_train_step(b_tables, a_tables, hash_offsets, products: torch.Tensor, alpha: float, batch_idx: int) -> None:
    """StreamingLogQCorrectionModule.train_step for every cascaded table, commons/layers.py:210-213 and
    :234-237 with the two evident repairs (`self.a[hash] = batch_idx`; iterate the modules and call
    train_step).  In place."""
    for b, a, off in zip(b_tables, a_tables, hash_offsets):
        h = logq_hash(products, off, b.numel())
        b[h] = ((1 - alpha) * b[h]) + (alpha * (batch_idx - a[h])).float()
        a[h] = batch_idx


# ------------------------------------------------- the reference's train step ----
class FlatTableCPU(torch.nn.Module):
    """nn.Embedding-backed FlatEmbedding restatement used as the timed CPU baseline: the same
    torch calls the reference makes (remainder -> F.embedding -> autograd -> optim.Adagrad)."""

    def __init__(self, weight: torch.Tensor):
        super().__init__()
        self.table = torch.nn.Embedding.from_pretrained(weight.clone(), freeze=False)

    def forward(self, ids):
        return self.table(torch.remainder(ids, self.table.num_embeddings).long())


def cpu_train_step(module: torch.nn.Module, optim: torch.optim.Optimizer, ids: torch.Tensor,
                   grad_out: torch.Tensor) -> torch.Tensor:
    """fwd + bwd + optimizer step as in embedding_module_gen.py:148-153 with an externally
    supplied upstream gradient (the embedding slice of a training step)."""
    optim.zero_grad()
    out = module(ids)
    out.backward(grad_out)
    optim.step()
    return out.detach()
