"""Streaming logQ correction on the B200 path.

Drop-ins for commons/layers.py:189-237 -- same constructors, forward / hash_fn / train_step and
state_dict keys (`b`, `a`; `models.<i>.b`, `models.<i>.a`) -- over the D = 1 gather / scatter kernels
of csrc/logq.cu.  The cascaded module serves all its bucket tables with ONE launch per call
(forward: 7 hashed reads + min per id; train_step: gather-new-values, then scatter).

The reference's train_step cannot run as written (commons/layers.py:213 indexes the Python float
`self.alpha`; :236-237 iterates `enumerate(...)` and calls `train_Step`); the semantics kept here are
the evident ones: `a[hash] = batch_idx`, every sub-module updated.  Duplicate ids: the new bucket
value is a function of the bucket alone, so all duplicates write the same number (no last-writer
ambiguity), and the right-hand side is evaluated from the old tables before anything is written.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch
import torch.nn as nn

from . import _native as N


def _ptr_array(tensors: Sequence[torch.Tensor]):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _check_tables(tables: Sequence[torch.Tensor], num_buckets: int) -> None:
    for t in tables:
        if t.dtype != torch.float32 or t.numel() != num_buckets:
            raise N.NativeError("logQ bucket tables must be float32 [num_buckets]")


def logq_fwd(b_tables: Sequence[torch.Tensor], hash_offsets: Sequence[int], products: torch.Tensor) -> torch.Tensor:
    """min over tables of -log(b_m[(id + offset_m) mod num_buckets]); same shape as `products`."""
    if products.dtype != torch.int64:
        raise N.NativeError(f"ids must be int64, got {products.dtype}")
    flat = products.contiguous().view(-1)
    dev = N.require_cuda(flat, *b_tables)
    nb = b_tables[0].numel()
    _check_tables(b_tables, nb)
    out = torch.empty(flat.shape, dtype=torch.float32, device=flat.device)
    offs = (C.c_int64 * len(hash_offsets))(*[int(o) for o in hash_offsets])
    N.check(N.load().recemb_logq_fwd(_ptr_array(b_tables), len(b_tables), offs, nb, N.ptr(flat), flat.numel(),
                                     N.ptr(out), dev, N.stream_ptr(dev)), "recemb_logq_fwd")
    return out.view(products.shape)


def logq_update(b_tables: Sequence[torch.Tensor], a_tables: Sequence[torch.Tensor], hash_offsets: Sequence[int],
                products: torch.Tensor, alpha: float, batch_idx: int,
                skip_mask: Optional[torch.Tensor] = None) -> None:
    """In-place streaming update of every table for the ids whose skip_mask is 0."""
    if products.dtype != torch.int64:
        raise N.NativeError(f"ids must be int64, got {products.dtype}")
    flat = products.contiguous().view(-1)
    if skip_mask is not None:
        skip_mask = skip_mask.contiguous().view(-1)
        if skip_mask.dtype == torch.bool:
            skip_mask = skip_mask.view(torch.uint8)
        if skip_mask.dtype != torch.uint8 or skip_mask.numel() != flat.numel():
            raise N.NativeError("skip_mask must be bool / uint8 with one entry per id")
    dev = N.require_cuda(flat, skip_mask, *b_tables, *a_tables)
    nb = b_tables[0].numel()
    _check_tables(list(b_tables) + list(a_tables), nb)
    scratch = torch.empty((flat.numel() * len(b_tables),), dtype=torch.float32, device=flat.device)
    offs = (C.c_int64 * len(hash_offsets))(*[int(o) for o in hash_offsets])
    N.check(N.load().recemb_logq_update(_ptr_array(b_tables), _ptr_array(a_tables), len(b_tables), offs, nb,
                                        N.ptr(flat), flat.numel(), N.ptr(skip_mask), float(alpha), int(batch_idx),
                                        N.ptr(scratch), dev, N.stream_ptr(dev)), "recemb_logq_update")


class StreamingLogQCorrectionModule(nn.Module):
    """commons/layers.py:189-213.  Buffers `b` (estimated inter-arrival steps, 1 / p_init at start)
    and `a` (last step a bucket was seen)."""

    def __init__(self, num_buckets, hash_offset, alpha: float = 0.05, p_init: float = 0.01, *, device=None):
        super().__init__()
        self.num_buckets = num_buckets
        self.hash_offset = hash_offset
        self.alpha = alpha
        self.p_init = p_init
        self.register_buffer("b", (1.0 / p_init) * torch.ones((num_buckets,), dtype=torch.float32, device=device))
        self.register_buffer("a", torch.zeros((num_buckets,), dtype=torch.float, device=device))

    def forward(self, products: torch.Tensor) -> torch.Tensor:
        return logq_fwd([self.b], [self.hash_offset], products)

    def hash_fn(self, products: torch.Tensor) -> torch.Tensor:
        from . import ops
        shifted = products + self.hash_offset  # wraps like the reference's int64 add
        return ops.row_index(shifted, N.HASH_FLOORMOD, self.num_buckets)

    @torch.no_grad()
    def train_step(self, products: torch.Tensor, batch_idx: int, skip_mask: Optional[torch.Tensor] = None):
        logq_update([self.b], [self.a], [self.hash_offset], products, self.alpha, batch_idx, skip_mask)


class CascadedStreamingLogQCorrectionModule(nn.Module):
    """commons/layers.py:217-237: elementwise minimum over sub-modules with different hash offsets.
    One kernel launch serves all of them."""

    def __init__(self, num_buckets, hash_offsets, alpha: float = 0.05, p_init: float = 0.01, *, device=None):
        super().__init__()
        self.models = nn.ModuleList([
            StreamingLogQCorrectionModule(num_buckets, offset, alpha, p_init, device=device)
            for offset in hash_offsets
        ])

    def forward(self, products: torch.Tensor) -> torch.Tensor:
        if len(self.models) == 0:
            return torch.empty((0,), device=products.device)
        return logq_fwd([m.b for m in self.models], [m.hash_offset for m in self.models], products)

    @torch.no_grad()
    def train_step(self, products: torch.Tensor, batch_idx: int, skip_mask: Optional[torch.Tensor] = None):
        """`skip_mask` (bool, one per id) replaces the caller's boolean compaction
        `product_ids.view(-1)[mask.view(-1) == 0]` (models/lthm/sequence/wrapper.py:133), which costs a
        host synchronisation for the compacted size."""
        if len(self.models) == 0:
            return
        ms = list(self.models)
        logq_update([m.b for m in ms], [m.a for m in ms], [m.hash_offset for m in ms], products,
                    ms[0].alpha, batch_idx, skip_mask)
