"""Sequence window: QueryTower's batch-wide trim (models/lthm/sequence/query_tower.py:73-86) decided on the
device BEFORE the embedding rows are moved.

The reference gathers [B, L, D] for the whole padded history, flips it (encoder.py:52-54) and only then drops
the columns that are padding in every row of the batch, keeping at least `export_span` of them
(query_tower.py:75-79, one host synchronisation).  Here one small kernel reduces the ids (or a mask) over the
batch and leaves the number of kept columns in device memory; the gather / k-shift kernels and the backward
plan read it (recemb_layout.window_keep): rows outside the window are neither read nor written and the kept
ones are stored compactly as [B, keep, D].  The host learns `keep` with one 8-byte read -- after the gather is
already enqueued -- where the reference synchronised anyway.

Pre-trimming on `ids == pad_id` alone is exact even though the reference's mask also contains
`||x|| < norm_threshold` (product_tower.py:49): the id-only mask is a subset of the full one, so every column
it drops is all-pad under the full mask too, and the reference rule applied afterwards to the window gives
(true trim - pre-trim) -- see tests/test_gpu_window.py.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _native as N


class SequenceWindow:
    """`keep` columns of sequences of length `seq_len`, decided on the device.

    drop="tail": all-pad columns are at the END of the sequences (right-padded histories, the layout the
    reference's Encoder receives before flip_all): the first `keep` positions stay.
    drop="head": at the START (left-padded, i.e. after the flip): the last `keep` positions stay."""

    def __init__(self, seq_len: int, drop: str, dev_pair: torch.Tensor):
        self.seq_len, self.drop = int(seq_len), drop
        self.side = 0 if drop == "tail" else 1
        self.dev = dev_pair                     # int32 [2] on the device: (keep, trim)
        self._host = torch.empty(2, dtype=torch.int32).pin_memory()
        self._host.copy_(dev_pair, non_blocking=True)
        self._ready = torch.cuda.Event()
        self._ready.record(torch.cuda.current_stream(dev_pair.device))
        self._keep: Optional[int] = None

    # ------------------------------------------------------------ builders ----
    @staticmethod
    def _build(data: torch.Tensor, kind: int, export_span: int, pad_id: int, drop: str) -> "SequenceWindow":
        if drop not in ("tail", "head"):
            raise ValueError("drop must be 'tail' or 'head'")
        if data.dim() != 2:
            raise N.NativeError("a sequence window is computed over [batch, seq_len]")
        data = data.contiguous()
        dev = N.require_cuda(data)
        b, length = data.shape
        lib = N.load()
        ws = torch.empty((int(lib.recemb_sequence_window_workspace_bytes(length)),), dtype=torch.uint8,
                         device=data.device)
        out = torch.empty(2, dtype=torch.int32, device=data.device)
        N.check(lib.recemb_sequence_window(N.ptr(data), kind, b, length, pad_id, int(export_span),
                                           0 if drop == "tail" else 1, N.ptr(ws), ws.numel(), N.ptr(out), dev,
                                           N.stream_ptr(dev)), "recemb_sequence_window")
        return SequenceWindow(length, drop, out)

    @staticmethod
    def from_ids(ids: torch.Tensor, export_span: int, pad_id: int = 0, drop: str = "tail") -> "SequenceWindow":
        if ids.dtype != torch.int64:
            raise N.NativeError(f"ids must be int64, got {ids.dtype}")
        return SequenceWindow._build(ids, 0, export_span, pad_id, drop)

    @staticmethod
    def from_mask(mask: torch.Tensor, export_span: int, drop: str = "head") -> "SequenceWindow":
        """mask: bool / uint8 [B, L], non-zero = padded (QueryTower's mask_inp)."""
        if mask.dtype == torch.bool:
            mask = mask.view(torch.uint8) if mask.is_contiguous() else mask.contiguous().view(torch.uint8)
        if mask.dtype != torch.uint8:
            raise N.NativeError(f"mask must be bool or uint8, got {mask.dtype}")
        return SequenceWindow._build(mask, 1, export_span, 0, drop)

    # ---------------------------------------------------------------- host ----
    @property
    def keep(self) -> int:
        """Kept columns (host value; the one synchronisation, on an event recorded right after the kernel)."""
        if self._keep is None:
            self._ready.synchronize()
            self._keep = int(self._host[0])
        return self._keep

    @property
    def trim(self) -> int:
        return self.seq_len - self.keep

    def keep_ptr(self) -> int:
        return self.dev.data_ptr()

    def narrow(self, t: torch.Tensor) -> torch.Tensor:
        """[B, L, ...] -> the kept columns [B, keep, ...] (a view), un-flipped orientation."""
        return t[:, :self.keep] if self.side == 0 else t[:, self.seq_len - self.keep:]

    def compact(self, full: torch.Tensor, batch: int) -> torch.Tensor:
        """The windowed kernels write [batch, keep, D] compactly at the start of a [batch * L, D] buffer."""
        dim = full.shape[-1]
        return full.view(-1, dim)[: batch * self.keep].view(batch, self.keep, dim)


def sequence_trim(mask_inp: torch.Tensor, export_span: int) -> int:
    """QueryTower's `trim` (query_tower.py:73-79) for a left-padded mask [B, L] (True = padded): drop-in for
    the two torch reductions + nonzero of the reference, one kernel + one 8-byte read."""
    return SequenceWindow.from_mask(mask_inp, export_span, drop="head").trim


# ------------------------------------------------------------ QueryTower's input sum ----
class _LookupSumFn(torch.autograd.Function):
    """out = where(mask, masked_row, base + sum_k table_k[h_k(ids_k)]) as ONE forward kernel
    (recemb_multi_gather_add_fwd); backward: the masked gradient is the gradient of `base`, every table gets it
    through the sort-based segmented reduction (tiny tables: a handful of very long runs, closed by the
    multi-level records), masked_row gets the column sums of the masked positions."""

    @staticmethod
    def forward(ctx, base, masked_row, mask, specs, record, *anchors):
        from . import ops
        terms = [(h.weight.detach(), ids, mode, arg) for (h, ids, mode, arg) in specs]
        out = ops.multi_gather_add_fwd(base, terms, mask=mask, masked_row=masked_row)
        ctx.specs, ctx.mask, ctx.has_base = specs, mask, base is not None
        ctx.row_shape = None if masked_row is None else tuple(masked_row.shape)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        from . import ops
        g = grad_out.contiguous()
        dim = g.shape[-1]
        gm, g_row = g, None
        if ctx.mask is not None:
            m = ctx.mask.bool().unsqueeze(-1)
            if ctx.needs_input_grad[1]:
                g_row = (g * m).reshape(-1, dim).sum(0).reshape(ctx.row_shape)
            gm = g.masked_fill(m, 0.0)
        grads = []
        for k, (holder, ids, mode, arg) in enumerate(ctx.specs):
            if not ctx.needs_input_grad[5 + k]:
                grads.append(None)
                continue
            plan = ops.BackwardPlan.build(ids, num_rows=holder.num_embeddings, hash_mode=mode, hash_arg=arg)
            grads.append(holder.consume(plan, gm.reshape(-1, dim)))
        return (gm if ctx.has_base and ctx.needs_input_grad[0] else None, g_row, None, None, None) + tuple(grads)


def fused_lookup_sum(base: Optional[torch.Tensor], lookups, mask: Optional[torch.Tensor] = None,
                     masked_row: Optional[torch.Tensor] = None) -> torch.Tensor:
    """QueryTower's input (models/lthm/sequence/query_tower.py:89-104) in one kernel:

        x = fused_lookup_sum(self.inp_proj(input),
                             [(self.action_embedding, labels), (self.time_embedding.hod, timestamp),
                              (self.time_embedding.how, timestamp), (self.time_embedding.dow, timestamp)],
                             mask=mask.squeeze(-1), masked_row=self.pad)

    == where(mask, pad, inp_proj(input) + action(labels) + hod(ts) + how(ts) + dow(ts)), bit-exact in fp32 (the
    adds run in the listed order).  `lookups`: (module, ids) pairs, module a recommendations_b200 FlatEmbedding
    (without output normalisation) or PatternFromTimelocal; ids of base's leading shape."""
    from . import layers as L
    specs, anchors = [], []
    for mod, ids in lookups:
        if isinstance(mod, L.PatternFromTimelocal):
            holder, mode, arg = mod.emb, N.HASH_DIV_FLOORMOD, int(mod.div)
        elif isinstance(mod, L.FlatEmbedding) and not mod._normalize_output and not mod._fused_pad_mask:
            holder, mode, arg = mod._emb_table, N.HASH_FLOORMOD, 0
        else:
            raise N.NativeError("fused_lookup_sum takes FlatEmbedding (plain) and PatternFromTimelocal modules")
        specs.append((holder, ids.long(), mode, arg))
        anchors.append(holder.grad_anchor())
    return _LookupSumFn.apply(base, masked_row, mask, specs, torch.is_grad_enabled(), *anchors)
