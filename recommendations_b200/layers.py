"""B200 drop-ins for the reference's embedding modules.

Same constructors, forward signatures and state_dict keys as
commons/layers.py (FlatEmbedding :44-61, KShiftEmbedding :125-185, QREmbedding
:102-123) and commons/transformers/layers.py (CosineVectorEmbedding :443-471),
with every lookup / pooling / backward / optimizer step running in the
hand-written sm_100a kernels behind include/recemb_b200.h.  Keyword-only
extras (dtype, device, fused optimizer) extend, never change, the reference
signatures.  CUDA only: a CPU tensor raises (no fallback).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as N
from . import ops
from .table import EmbeddingTable, FusedOptimizerConfig


# ---------------------------------------------------------- autograd glue ----
# The backward plan (id hashing + radix sort of the lookup slots) is a function of the ids alone.
# When a lookup will need a gradient it is therefore built during FORWARD, on a side stream next
# to the gather / pooling kernel (which is HBM-bound and leaves the sort's latency-bound passes
# room), and backward only waits for it: ~0.5 ms per step off the critical path at cfg 2.
# RECEMB_PLAN_IN_FORWARD=0 (or layers.PLAN_IN_FORWARD = False) builds it in backward instead --
# needed when only the forward pass is captured into a CUDA graph.
import os as _os

PLAN_IN_FORWARD = _os.environ.get("RECEMB_PLAN_IN_FORWARD", "1") != "0"
_side_streams = {}


def _plan_early(ctx, ids: torch.Tensor, wanted: bool, build) -> None:
    ctx.plan, ctx.plan_ready = None, None
    if not (wanted and PLAN_IN_FORWARD and ids.is_cuda):
        return
    dev = ids.device
    main = torch.cuda.current_stream(dev)
    side = _side_streams.get(dev)
    if side is None:
        side = _side_streams[dev] = torch.cuda.Stream(device=dev)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        ctx.plan = build()
        ctx.plan_ready = torch.cuda.Event()
        ctx.plan_ready.record(side)
    ids.record_stream(side)
    ctx.plan.buf.record_stream(main)


def _plan_take(ctx, build):
    """The plan built in forward (after waiting for it on the current stream) or a fresh one."""
    if getattr(ctx, "plan", None) is None:
        return build()
    plan = ctx.plan
    torch.cuda.current_stream(plan.buf.device).wait_event(ctx.plan_ready)
    plan.buf.record_stream(torch.cuda.current_stream(plan.buf.device))
    ctx.plan = None
    return plan


class _GatherFn(torch.autograd.Function):
    """out = epilogue(table[h(ids)] (+ table2[h2(ids)])); backward = plan + segmented reduce."""

    @staticmethod
    def forward(ctx, anchor, anchor2, ids, holder, holder2, hash_mode, hash_mode2, hash_arg,
                epilogue, zero_pad, pad_id, flip_len=0, record=True, window=None):
        # `record` = torch.is_grad_enabled() at the call site (grad mode is always off in here and
        # needs_input_grad ignores no_grad): inference builds no inverse norms and no plan
        need_grad = record and (ctx.needs_input_grad[0] or (anchor2 is not None and ctx.needs_input_grad[1]))
        _check_window(window, ids, holder, holder2)
        out, inv = ops.gather_fwd(
            holder.weight.detach(), ids, hash_mode=hash_mode, hash_arg=hash_arg,
            table2=None if holder2 is None else holder2.weight.detach(), hash_mode2=hash_mode2,
            epilogue=epilogue, zero_pad=zero_pad, pad_id=pad_id, want_inv_norm=need_grad,
            flip_len=flip_len, window=window)
        ctx.holder, ctx.holder2 = holder, holder2
        ctx.flip_len, ctx.window = flip_len, window
        ctx.cfg = (hash_mode, hash_mode2, hash_arg, epilogue, zero_pad, pad_id)
        _plan_early(ctx, ids, holder2 is None and record and ctx.needs_input_grad[0]
                    and not (holder.sparse and holder.fused is None),
                    lambda: holder.build_plan(
                        ids, num_rows=holder.num_embeddings, hash_mode=hash_mode, hash_arg=hash_arg,
                        zero_pad=zero_pad, pad_id=pad_id,
                        pad_row=-1 if holder.padding_idx is None else holder.padding_idx, flip_len=flip_len,
                        window=window))
        if window is not None:  # everything is enqueued: now the host may learn `keep`
            out, inv = _compact(window, ids, out, inv)
        ctx.save_for_backward(ids, inv, out if epilogue == N.EPI_L2NORM else None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        ids, inv, out = ctx.saved_tensors
        hash_mode, hash_mode2, hash_arg, epilogue, zero_pad, pad_id = ctx.cfg
        dim = grad_out.shape[-1]
        g = grad_out.contiguous().view(-1, dim)
        if epilogue == N.EPI_L2NORM:
            g = ops.epilogue_bwd(g, out, inv, N.EPI_L2NORM)
        grads = [None, None]
        for i, (holder, mode, needs) in enumerate(((ctx.holder, hash_mode, ctx.needs_input_grad[0]),
                                                   (ctx.holder2, hash_mode2, ctx.needs_input_grad[1]))):
            if holder is None or not needs:
                continue
            if holder.sparse and holder.fused is None:
                rows = ops.row_index(ids.view(-1), mode, holder.num_embeddings, hash_arg)
                vals = g.to(holder.weight.dtype)
                if ctx.flip_len:  # gradient rows are in mirrored order
                    vals = vals.view(-1, ctx.flip_len, dim).flip(1).reshape(-1, dim)
                # fused pad mask: the forward never read the table at id == pad_id positions, so they
                # carry no gradient (the plan-based modes drop those slots the same way)
                keep = (ids.reshape(-1) != pad_id) if zero_pad else None
                grads[i] = _coo(rows, vals, holder, keep)
                continue
            plan = _plan_take(ctx, lambda: holder.build_plan(
                ids, num_rows=holder.num_embeddings, hash_mode=mode, hash_arg=hash_arg,
                zero_pad=zero_pad, pad_id=pad_id,
                pad_row=-1 if holder.padding_idx is None else holder.padding_idx, flip_len=ctx.flip_len,
                window=ctx.window))
            grads[i] = holder.consume(plan, g)
        return (grads[0], grads[1]) + (None,) * 12


def _check_window(window, ids, *holders) -> None:
    if window is None:
        return
    if ids.dim() != 2 or ids.shape[1] != window.seq_len:
        raise N.NativeError(f"windowed lookup: ids must be [batch, {window.seq_len}], got {tuple(ids.shape)}")
    for h in holders:
        if h is not None and h.sparse and h.fused is None:
            raise N.NativeError("windowed lookups do not support the sparse-COO gradient mode")


def _compact(window, ids, out, inv):
    """Full-size kernel outputs -> the compact [batch, keep, dim] prefix the windowed kernels wrote."""
    b = ids.shape[0]
    out = window.compact(out, b)
    if inv is not None:
        inv = inv[: b * window.keep]
    return out, inv


def _coo(rows: torch.Tensor, vals: torch.Tensor, holder: EmbeddingTable,
         keep: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The uncoalesced COO gradient nn.Embedding(sparse=True) produces (nnz == lookups that
    read the table; `keep` masks positions the forward zero-filled)."""
    if holder.padding_idx is not None:
        k2 = rows != holder.padding_idx
        keep = k2 if keep is None else (keep & k2)
    if keep is not None:
        rows, vals = rows[keep], vals[keep]
    return torch.sparse_coo_tensor(rows.view(1, -1), vals, holder.weight.shape)


class _KShiftFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, ids, holder, num_shifts, epilogue, flip_len=0, record=True, window=None):
        record = record and ctx.needs_input_grad[0]
        _check_window(window, ids, holder)
        out, inv = ops.kshift_fwd(holder.weight.detach(), ids, num_shifts, epilogue,
                                  want_inv_norm=record, flip_len=flip_len, window=window)
        ctx.holder, ctx.k, ctx.epilogue, ctx.flip_len, ctx.window = holder, num_shifts, epilogue, flip_len, window
        _plan_early(ctx, ids, record and not (holder.sparse and holder.fused is None),
                    lambda: holder.build_plan(
                        ids, num_rows=holder.num_embeddings, hash_mode=N.HASH_ROTL_FLOORMOD, slots_per_id=num_shifts,
                        pad_row=-1 if holder.padding_idx is None else holder.padding_idx, flip_len=flip_len,
                        window=window))
        if window is not None:
            out, inv = _compact(window, ids, out, inv)
        ctx.save_for_backward(ids, inv, out if epilogue == N.EPI_L2NORM else None)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        ids, inv, out = ctx.saved_tensors
        holder, k = ctx.holder, ctx.k
        dim = grad_out.shape[-1]
        g2d = grad_out.contiguous().view(-1, dim)
        sparse_coo = holder.sparse and holder.fused is None
        if ctx.epilogue == N.EPI_RSQRT_K and not sparse_coo:
            # x / sqrt(k) (commons/layers.py:170): its backward is a division of every gradient
            # element -- folded into the segmented reduction, dx is never materialised
            plan = _plan_take(ctx, lambda: holder.build_plan(
                ids, num_rows=holder.num_embeddings, hash_mode=N.HASH_ROTL_FLOORMOD, slots_per_id=k,
                pad_row=-1 if holder.padding_idx is None else holder.padding_idx, flip_len=ctx.flip_len,
                window=ctx.window))
            return (holder.consume(plan, g2d, slots_per_grad_row=k, grad_div=math.sqrt(k)),
                    None, None, None, None, None, None, None)
        dx = ops.epilogue_bwd(g2d, out, inv, ctx.epilogue, k)
        if sparse_coo:
            flat = ids.contiguous().view(-1)
            rows = torch.cat([ops.row_index(flat, N.HASH_ROTL_FLOORMOD, holder.num_embeddings, c)
                              for c in range(k)])
            if ctx.flip_len:
                dx = dx.view(-1, ctx.flip_len, dim).flip(1).reshape(-1, dim)
            vals = dx.to(holder.weight.dtype).repeat(k, 1)
            return (_coo(rows, vals, holder), None, None, None, None, None, None, None)
        plan = _plan_take(ctx, lambda: holder.build_plan(
            ids, num_rows=holder.num_embeddings, hash_mode=N.HASH_ROTL_FLOORMOD, slots_per_id=k,
            pad_row=-1 if holder.padding_idx is None else holder.padding_idx, flip_len=ctx.flip_len,
            window=ctx.window))
        return (holder.consume(plan, dx, slots_per_grad_row=k), None, None, None, None, None, None, None)


class _PoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, ids, lengths, per_slot_weight, holder, hash_mode, hash_arg, pool_mode,
                last_n, zero_pad, pad_id, record=True):
        out = ops.pool_fwd(holder.weight.detach(), ids, lengths=lengths, last_n=last_n,
                           per_slot_weight=per_slot_weight, hash_mode=hash_mode, hash_arg=hash_arg,
                           pool_mode=pool_mode, zero_pad=zero_pad, pad_id=pad_id)
        ctx.holder = holder
        ctx.cfg = (hash_mode, hash_arg, pool_mode, last_n, zero_pad, pad_id)
        ctx.save_for_backward(ids, lengths, per_slot_weight)
        _plan_early(ctx, ids, record and ctx.needs_input_grad[0], lambda: holder.build_plan(
            ids, num_rows=holder.num_embeddings, hash_mode=hash_mode, hash_arg=hash_arg,
            zero_pad=zero_pad, pad_id=pad_id,
            pad_row=-1 if holder.padding_idx is None else holder.padding_idx, bag_size=ids.shape[1],
            lengths=lengths, last_n=last_n))
        if ctx.plan is not None and lengths is not None:
            lengths.record_stream(_side_streams[ids.device])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        ids, lengths, per_slot_weight = ctx.saved_tensors
        holder = ctx.holder
        hash_mode, hash_arg, pool_mode, last_n, zero_pad, pad_id = ctx.cfg
        m, p = ids.shape
        plan = _plan_take(ctx, lambda: holder.build_plan(
            ids, num_rows=holder.num_embeddings, hash_mode=hash_mode, hash_arg=hash_arg,
            zero_pad=zero_pad, pad_id=pad_id,
            pad_row=-1 if holder.padding_idx is None else holder.padding_idx, bag_size=p,
            lengths=lengths, last_n=last_n))
        scale = None
        if pool_mode == N.POOL_MEAN:
            scale = 1.0 / pooled_counts(ids, lengths, last_n, zero_pad, pad_id).clamp_(min=1).float()
        gw = holder.consume(plan, grad_out.contiguous().view(m, -1), slots_per_grad_row=p,
                            slot_weight=per_slot_weight, grad_row_scale=scale)
        return (gw,) + (None,) * 11


def pooled_counts(ids, lengths, last_n, zero_pad, pad_id) -> torch.Tensor:
    """Number of slots each bag pools (the MEAN divisor); tiny index arithmetic on the device."""
    m, p = ids.shape
    pos = torch.arange(p, device=ids.device).unsqueeze(0)
    hi = torch.full((m, 1), p, device=ids.device) if lengths is None else \
        lengths.to(torch.int64).clamp(0, p).unsqueeze(1)
    lo = (hi - last_n).clamp(min=0) if last_n > 0 else torch.zeros_like(hi)
    ok = (pos >= lo) & (pos < hi)
    if zero_pad:
        ok &= ids != pad_id
    return ok.sum(dim=1)


def _flip_len(ids: torch.Tensor, flip: bool) -> int:
    """flip_sequences: write the [.., L, D] output mirrored along L (Encoder.flip_all,
    models/lthm/sequence/encoder.py:52-54: right-padded histories -> left-padded) inside the
    gather instead of a torch.flip copy afterwards."""
    if not flip:
        return 0
    if ids.dim() < 2:
        raise N.NativeError("flip_sequences needs ids of shape [..., L]")
    return int(ids.shape[-1])


# ------------------------------------------------------------------ modules ----
class FlatEmbedding(nn.Module):
    """commons/layers.py:44-61: row = floor_mod(id, N); out = table[row]; optional L2 norm.

    state_dict key: `_emb_table.weight`.  `fused_pad_mask=True` additionally zero-fills
    positions whose id is 0 without reading the table (the `ids == 0` mask of
    models/lthm/sequence/product_tower.py:47-59 folded into the gather)."""

    def __init__(self, num_embeddings: int, emb_dim: int, padding_idx: int = None,
                 zero_init: bool = False, normalize_output: bool = False, *,
                 dtype: torch.dtype = torch.float32, device=None, sparse: bool = False,
                 fused_optimizer: Optional[FusedOptimizerConfig] = None,
                 fused_pad_mask: bool = False, flip_sequences: bool = False):
        super().__init__()
        self._flip_sequences = flip_sequences
        self._num_embeddings = num_embeddings
        self._emb_dim = emb_dim
        self.padding_idx = padding_idx
        self._emb_table = EmbeddingTable(num_embeddings, emb_dim, padding_idx=padding_idx,
                                         sparse=sparse, dtype=dtype, device=device)
        self._normalize_output = normalize_output
        self._fused_pad_mask = fused_pad_mask
        if zero_init:
            self._emb_table.weight.data.fill_(0.0)
        if fused_optimizer is not None:
            self._emb_table.enable_fused_optimizer(fused_optimizer)

    def forward(self, x: torch.Tensor, window=None) -> torch.Tensor:
        """window (sequence.SequenceWindow over x's [batch, L]): only the kept columns are looked up and the
        result is [batch, keep, D] -- QueryTower's trim (query_tower.py:73-86) before the rows are moved."""
        t = self._emb_table
        return _GatherFn.apply(t.grad_anchor(), None, x, t, None, N.HASH_FLOORMOD, 0, 0,
                               N.EPI_L2NORM if self._normalize_output else N.EPI_NONE,
                               self._fused_pad_mask, 0, _flip_len(x, self._flip_sequences), torch.is_grad_enabled(),
                               window)


class PatternFromTimelocal(nn.Module):
    """commons/layers.py:13-41: index = floor_mod(floor_divide(t, div), mod); out = emb[index]
    (hour-of-day / hour-of-week / day-of-week tables of QueryTower,
    models/lthm/sequence/query_tower.py:27-33).  The reference chains floor_divide, remainder and an
    embedding lookup; here the index arithmetic runs inside the gather kernel (DIV_FLOORMOD).
    The reference constructor forgets super().__init__() and passes `emb_dim=` to nn.Embedding
    (both TypeErrors); semantics otherwise identical.  state_dict key: `emb.weight`."""

    def __init__(self, div, mod, emb_dim, *, dtype: torch.dtype = torch.float32, device=None):
        super().__init__()
        self.div = div
        self.mod = mod
        self.emb_dim = emb_dim
        self.emb = EmbeddingTable(mod, emb_dim, dtype=dtype, device=device) if emb_dim > 0 else nn.Identity()

    def index(self, x: torch.Tensor) -> torch.Tensor:
        return ops.row_index(x.long(), N.HASH_DIV_FLOORMOD, self.mod, self.div)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.emb_dim <= 0:
            return self.index(x)
        return _GatherFn.apply(self.emb.grad_anchor(), None, x.long(), self.emb, None, N.HASH_DIV_FLOORMOD, 0,
                               self.div, N.EPI_NONE, False, 0, 0, torch.is_grad_enabled())


class KShiftEmbedding(nn.Module):
    """commons/layers.py:125-185: k hashed lookups into one shared table, summed in the
    order c = 0..k-1, then L2-normalised or scaled by 1/sqrt(k).  One kernel instead of
    2k-1 launches.  state_dict key: `emb.weight`."""

    def __init__(self, num_embeddings: int, emb_dim: int, num_shifts: int = 8,
                 normalize_output: bool = False, sparse: bool = False, *,
                 dtype: torch.dtype = torch.float32, device=None,
                 fused_optimizer: Optional[FusedOptimizerConfig] = None, flip_sequences: bool = False):
        super().__init__()
        self._flip_sequences = flip_sequences
        self.emb = EmbeddingTable(num_embeddings, emb_dim, sparse=sparse, dtype=dtype, device=device)
        self._num_embeddings = num_embeddings
        self._num_shifts = num_shifts
        self._num_bits = 64
        self._normalize_output = normalize_output
        if fused_optimizer is not None:
            self.emb.enable_fused_optimizer(fused_optimizer)

    def forward(self, id_: torch.Tensor, window=None) -> torch.Tensor:
        """window: see FlatEmbedding.forward."""
        return _KShiftFn.apply(self.emb.grad_anchor(), id_, self.emb, self._num_shifts,
                               N.EPI_L2NORM if self._normalize_output else N.EPI_RSQRT_K,
                               _flip_len(id_, self._flip_sequences), torch.is_grad_enabled(), window)

    def get_row_idx(self, x: torch.Tensor, col_idx: int) -> torch.Tensor:
        """Bit-exact commons/layers.py:174-185 (wrapping <<, arithmetic >>, floor-mod)."""
        return ops.row_index(x, N.HASH_ROTL_FLOORMOD, self._num_embeddings, col_idx)


class QREmbedding(nn.Module):
    """commons/layers.py:102-123 (the reference constructor forgets super().__init__();
    semantics otherwise identical): d = floor(sqrt(N)); x' = id mod d^2;
    out = emb_q[x' // d] + emb_r[x' mod d].  state_dict keys: `emb_q.weight`, `emb_r.weight`."""

    def __init__(self, num_embeddings: int, emb_dim: int, normalize_output: bool, *,
                 dtype: torch.dtype = torch.float32, device=None,
                 fused_optimizer: Optional[FusedOptimizerConfig] = None):
        super().__init__()
        self._div = int(math.sqrt(num_embeddings))
        self.num_embeddings = self._div * self._div
        self.emb_dim = emb_dim
        self.emb_q = EmbeddingTable(self._div, emb_dim, dtype=dtype, device=device)
        self.emb_r = EmbeddingTable(self._div, emb_dim, dtype=dtype, device=device)
        self.normalize_output = normalize_output
        if fused_optimizer is not None:
            self.emb_q.enable_fused_optimizer(fused_optimizer)
            self.emb_r.enable_fused_optimizer(FusedOptimizerConfig(**vars(fused_optimizer)))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return _GatherFn.apply(self.emb_q.grad_anchor(), self.emb_r.grad_anchor(), x, self.emb_q,
                               self.emb_r, N.HASH_QR_QUOTIENT, N.HASH_QR_REMAINDER, self._div,
                               N.EPI_L2NORM if self.normalize_output else N.EPI_NONE, False, 0, 0, torch.is_grad_enabled())


class PooledEmbeddingBag(nn.Module):
    """Multi-hot pooled lookup (north_star item 1; the op behind nn.EmbeddingBag(mode='sum')
    at commons/transformers/layers.py:457).  ids [num_bags, bag_size] int64, optional
    per-bag valid `lengths`, `last_n` window, per_sample_weights; sum or mean; padding ids
    (`pad_id`, default 0 = CATEGORICAL_VAR_HASH_PAD_TOKEN, commons/feature_utils.py:8) are
    skipped when `skip_pad`.  state_dict key: `emb.weight`."""

    def __init__(self, num_embeddings: int, emb_dim: int, mode: str = "sum", *, last_n: int = 0,
                 hash_ids: bool = True, skip_pad: bool = False, pad_id: int = 0,
                 padding_idx: Optional[int] = None, dtype: torch.dtype = torch.float32, device=None,
                 fused_optimizer: Optional[FusedOptimizerConfig] = None, validate_ids: bool = True):
        super().__init__()
        if mode not in ("sum", "mean"):
            raise ValueError("mode must be 'sum' or 'mean' (use last_n for the last-N window)")
        self.emb = EmbeddingTable(num_embeddings, emb_dim, padding_idx=padding_idx, dtype=dtype,
                                  device=device)
        self.mode, self.last_n = mode, int(last_n)
        self.hash_ids, self.skip_pad, self.pad_id = hash_ids, skip_pad, pad_id
        # hash_ids=False: ids are rows.  nn.EmbeddingBag raises on an out-of-range index; the kernels
        # only DROP such slots (never read or update out of bounds), so the module checks the range
        # itself (one device reduction + sync per call; validate_ids=False skips it).
        self.validate_ids = validate_ids
        if fused_optimizer is not None:
            self.emb.enable_fused_optimizer(fused_optimizer)

    def forward(self, ids: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                per_sample_weights: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not self.hash_ids and self.validate_ids and ids.numel():
            lo, hi = torch.aminmax(ids)
            if int(lo) < 0 or int(hi) >= self.emb.num_embeddings:
                raise IndexError(f"index out of range in PooledEmbeddingBag(hash_ids=False): ids span "
                                 f"[{int(lo)}, {int(hi)}], table has {self.emb.num_embeddings} rows")
        return _PoolFn.apply(self.emb.grad_anchor(), ids, lengths, per_sample_weights, self.emb,
                             N.HASH_FLOORMOD if self.hash_ids else N.HASH_IDENTITY, 0,
                             N.POOL_SUM if self.mode == "sum" else N.POOL_MEAN, self.last_n,
                             self.skip_pad, self.pad_id, torch.is_grad_enabled())


class CosineVectorEmbedding(nn.Module):
    """commons/transformers/layers.py:443-471: project -> bucketize -> fixed-size-bag sum.
    The projection matmul and bucketize stay on torch (dense math, out of scope); the bag
    sum over n_proj rows and its backward run in the pooled-bag kernels.
    state_dict: `projection_mat`, `grid`, `pos_offset`, `emb.weight`."""

    def __init__(self, inp_dim: int, emb_dim: int, n_proj: int = 16, num_bins: int = 20, *,
                 device=None, fused_optimizer: Optional[FusedOptimizerConfig] = None):
        super().__init__()
        proj = F.normalize(torch.randn((inp_dim, n_proj)), p=2.0, dim=0)
        self.register_buffer("projection_mat", proj.to(device), persistent=True)
        resolution = 2.0 / float(num_bins)
        grid = torch.linspace(-1.0, 1.0, steps=num_bins + 1)[:-1] + 0.5 * resolution
        self.register_buffer("grid", grid.to(device), persistent=True)
        pos_offset = ((num_bins + 1) * torch.arange(0, n_proj, dtype=torch.long)).reshape(n_proj)
        self.register_buffer("pos_offset", pos_offset.to(device), persistent=True)
        # nn.EmbeddingBag initialises N(0, 1) like nn.Embedding
        self.emb = EmbeddingTable((num_bins + 1) * n_proj, emb_dim, device=device)
        self.emb_dim, self.n_proj, self.num_bins = emb_dim, n_proj, num_bins
        if fused_optimizer is not None:
            self.emb.enable_fused_optimizer(fused_optimizer)

    def bucket_indices(self, x: torch.Tensor) -> torch.Tensor:
        z = F.normalize(x, p=2.0, dim=-1) @ self.projection_mat
        return (torch.bucketize(z, self.grid).view(-1, self.n_proj) + self.pos_offset.unsqueeze(0))

    def bag(self, idxs: torch.Tensor) -> torch.Tensor:
        """EmbeddingBag(mode='sum') half (commons/transformers/layers.py:469): idxs [M, n_proj] -> [M, emb_dim]."""
        return _PoolFn.apply(self.emb.grad_anchor(), idxs, None, None, self.emb, N.HASH_IDENTITY, 0,
                             N.POOL_SUM, 0, False, 0, torch.is_grad_enabled())

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        bs, seq_len, _ = x.size()
        return self.bag(self.bucket_indices(x)).view(bs, seq_len, self.emb_dim)
