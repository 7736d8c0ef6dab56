"""Tensor-level wrappers over the C ABI (include/recemb_b200.h).

Each function takes CUDA tensors, launches on torch's current stream of the
tensors' device and returns CUDA tensors.  No function here computes on the
CPU or falls back to torch ops: a missing library or a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import torch

from . import _native as N


def _flat_ids(ids: torch.Tensor) -> torch.Tensor:
    if ids.dtype != torch.int64:
        # the reference asserts int64 history ids (models/lthm/sequence/wrapper.py:52)
        raise N.NativeError(f"ids must be int64, got {ids.dtype}")
    return ids.contiguous().view(-1)


# ------------------------------------------------------------------ hashing ----
def row_index(ids: torch.Tensor, hash_mode: int, num_rows: int, hash_arg: int = 0) -> torch.Tensor:
    """rows = transform(ids); same shape, int64.  (commons/layers.py:174-185, :57)"""
    flat = _flat_ids(ids)
    dev = N.require_cuda(flat)
    out = torch.empty_like(flat)
    N.check(N.load().recemb_row_index(N.ptr(flat), flat.numel(), hash_mode, num_rows, hash_arg,
                                      N.ptr(out), dev, N.stream_ptr(dev)), "recemb_row_index")
    return out.view(ids.shape)


# ------------------------------------------------------------------ forward ----
def gather_fwd(table: torch.Tensor, ids: torch.Tensor, *, hash_mode: int = N.HASH_FLOORMOD,
               hash_arg: int = 0, table2: Optional[torch.Tensor] = None,
               hash_mode2: int = N.HASH_FLOORMOD, epilogue: int = N.EPI_NONE,
               zero_pad: bool = False, pad_id: int = 0, want_inv_norm: bool = False,
               out: Optional[torch.Tensor] = None, ids_per_table: int = 0,
               flip_len: int = 0, window=None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """window (sequence.SequenceWindow): only the kept columns of every sequence are looked up; the result is
    the FULL-size buffer viewed as [*ids.shape, dim] whose first batch * keep rows hold [batch, keep, dim]
    (window.compact(out, batch) once the host may know `keep`)."""
    flat = _flat_ids(ids)
    dev = N.require_cuda(table, table2, flat, out)
    n, dim = flat.numel(), table.shape[1]
    if out is None:
        out = torch.empty((n, dim), dtype=table.dtype, device=table.device)
    elif out.numel() != n * dim or out.dtype != table.dtype:
        raise N.NativeError("preallocated `out` has the wrong size / dtype")
    inv = None
    if want_inv_norm and epilogue == N.EPI_L2NORM:
        inv = torch.empty((n,), dtype=torch.float32, device=table.device)
    rows_per_table = table.shape[0]
    layout = N.make_layout(ids_per_table=ids_per_table, flip_len=flip_len, window=window)
    if ids_per_table:
        n_tables = -(-n // ids_per_table)
        if table.shape[0] % n_tables:
            raise N.NativeError("stacked table rows are not a multiple of the number of tables")
        rows_per_table = table.shape[0] // n_tables
    N.check(N.load().recemb_gather_fwd(
        N.ptr(table), rows_per_table, N.ptr(table2), 0 if table2 is None else table2.shape[0], dim,
        N.dtype_code(table.dtype), N.ptr(flat), n, layout, hash_mode, hash_mode2, hash_arg, epilogue,
        int(zero_pad), pad_id, N.ptr(out), N.ptr(inv), dev, N.stream_ptr(dev)), "recemb_gather_fwd")
    return out.view(*ids.shape, dim), inv


def kshift_fwd(table: torch.Tensor, ids: torch.Tensor, num_shifts: int, epilogue: int,
               want_inv_norm: bool = False, flip_len: int = 0,
               window=None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    flat = _flat_ids(ids)
    dev = N.require_cuda(table, flat)
    n, dim = flat.numel(), table.shape[1]
    out = torch.empty((n, dim), dtype=table.dtype, device=table.device)
    inv = None
    if want_inv_norm and epilogue == N.EPI_L2NORM:
        inv = torch.empty((n,), dtype=torch.float32, device=table.device)
    N.check(N.load().recemb_kshift_fwd_layout(
        N.ptr(table), table.shape[0], dim, N.dtype_code(table.dtype), N.ptr(flat), n, num_shifts,
        epilogue, N.make_layout(flip_len=flip_len, window=window), N.ptr(out), N.ptr(inv), dev, N.stream_ptr(dev)),
        "recemb_kshift_fwd")
    return out.view(*ids.shape, dim), inv


def multi_gather_add_fwd(base: Optional[torch.Tensor], terms, *, mask: Optional[torch.Tensor] = None,
                         masked_row: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i] = mask[i] ? masked_row : base[i] + sum_k table_k[transform_k(ids_k[i])] in one pass (added in
    order k = 0, 1, ...).  terms: sequence of (table [rows, dim], ids [n...] int64, hash_mode, hash_arg)."""
    if not terms and base is None:
        raise N.NativeError("nothing to add")
    first_ids = terms[0][1] if terms else None
    shape = tuple(base.shape[:-1]) if base is not None else tuple(first_ids.shape)
    dim = base.shape[-1] if base is not None else terms[0][0].shape[1]
    dtype = base.dtype if base is not None else terms[0][0].dtype
    n = 1
    for d in shape:
        n *= d
    keep = []   # keeps the contiguous copies alive until the launch is enqueued
    if base is not None:
        base = base.contiguous()
    arr = (N.GatherTerm * max(len(terms), 1))()
    for k, (table, ids, hash_mode, hash_arg) in enumerate(terms):
        flat = _flat_ids(ids)
        if flat.numel() != n or table.shape[1] != dim or table.dtype != dtype:
            raise N.NativeError(f"term {k}: ids / table do not match the base shape [{n}, {dim}] {dtype}")
        keep.append(flat)
        arr[k].table, arr[k].num_rows, arr[k].ids = table.data_ptr(), table.shape[0], flat.data_ptr()
        arr[k].hash_mode, arr[k].hash_arg = hash_mode, hash_arg
    if mask is not None:
        mask = mask.contiguous().view(torch.uint8) if mask.dtype == torch.bool else mask.contiguous()
        if mask.numel() != n or mask.dtype != torch.uint8 or masked_row is None:
            raise N.NativeError("mask must be bool / uint8 of the base's leading shape, with a masked_row")
        masked_row = masked_row.contiguous().view(-1).to(dtype)
        if masked_row.numel() != dim:
            raise N.NativeError("masked_row must hold one row")
    dev = N.require_cuda(base, mask, masked_row, *[t[0] for t in terms], *keep)
    ref = base if base is not None else terms[0][0]
    out = torch.empty(shape + (dim,), dtype=dtype, device=ref.device)
    N.check(N.load().recemb_multi_gather_add_fwd(N.ptr(base), arr, len(terms), n, dim, N.dtype_code(dtype), N.ptr(mask),
                                                 N.ptr(masked_row), N.ptr(out), dev, N.stream_ptr(dev)),
            "recemb_multi_gather_add_fwd")
    return out


def pool_fwd(table: torch.Tensor, ids: torch.Tensor, *, lengths: Optional[torch.Tensor] = None,
             last_n: int = 0, per_slot_weight: Optional[torch.Tensor] = None,
             hash_mode: int = N.HASH_FLOORMOD, hash_arg: int = 0, pool_mode: int = N.POOL_SUM,
             zero_pad: bool = False, pad_id: int = 0, num_rows: Optional[int] = None,
             shard_world: int = 1, shard_rank: int = 0, bags_per_table: int = 0,
             num_tables: int = 0, out: Optional[torch.Tensor] = None, out_features: int = 0,
             out_feature_offset: int = 0) -> torch.Tensor:
    """`num_rows` = GLOBAL rows of one table (default: table.shape[0]).  Sharded (shard_world > 1):
    `table` is this rank's row-wise shard and the result is this owner's partial pool.
    Table-batched (bags_per_table > 0): bag g uses table (g // bags_per_table) [% num_tables] of
    the stacked (local) table.
    out_features = F > 0 (table-batched, `out` = a [bags_per_table, F, dim] tensor): the pooled row of bag g is
    written at out[g % bags_per_table, g // bags_per_table + out_feature_offset] -- the dot interaction's input
    layout, no concatenation afterwards; returns `out`."""
    if ids.dim() != 2:
        raise N.NativeError("pooled bags take ids of shape [num_bags, bag_size]")
    ids = ids.contiguous()
    if ids.dtype != torch.int64:
        raise N.NativeError(f"ids must be int64, got {ids.dtype}")
    if lengths is not None:
        lengths = lengths.to(torch.int32).contiguous()
    if per_slot_weight is not None:
        per_slot_weight = per_slot_weight.to(torch.float32).contiguous()
    dev = N.require_cuda(table, ids, lengths, per_slot_weight)
    m, p = ids.shape
    dim = table.shape[1]
    if out_features:
        if out is None or not out.is_contiguous() or out.dtype != table.dtype or bags_per_table <= 0 or \
                tuple(out.shape) != (bags_per_table, out_features, dim):
            raise N.NativeError("out_features needs table-batched bags and a contiguous `out` of shape "
                                "[bags_per_table, out_features, dim] in the table dtype")
        N.require_cuda(out)
    elif out is None:
        out = torch.empty((m, dim), dtype=table.dtype, device=table.device)
    elif out.numel() != m * dim or out.dtype != table.dtype or not out.is_contiguous():
        raise N.NativeError("preallocated `out` has the wrong size / dtype")
    N.check(N.load().recemb_pool_fwd(
        N.ptr(table), table.shape[0] if num_rows is None else num_rows, dim,
        N.dtype_code(table.dtype), N.ptr(ids), m, p, N.ptr(lengths), last_n, N.ptr(per_slot_weight),
        hash_mode, hash_arg, pool_mode, int(zero_pad), pad_id,
        N.make_layout(bags_per_table * p, num_tables, shard_world, shard_rank, out_features=out_features,
                      out_feature_offset=out_feature_offset), N.ptr(out), dev,
        N.stream_ptr(dev)), "recemb_pool_fwd")
    return out


# ----------------------------------------------------------------- backward ----
@dataclass
class BackwardPlan:
    """Sorted (row, slot) pairs of one forward call: the dedup half of backward."""
    buf: torch.Tensor  # uint8 device buffer holding the plan
    n_slots: int
    num_rows: int
    slots_per_id: int
    slots_cover_grad: bool = True  # False: slots are explicit gradient-row indices (routed entries)
    # (ids, build keyword arguments) of a plan built from ids: lets table.commit_step() rebuild ONE
    # plan over the concatenated ids of several lookups of a table (accumulate mode)
    recipe: Optional[tuple] = None

    @staticmethod
    def build(ids: torch.Tensor, *, num_rows: int, hash_mode: int = N.HASH_FLOORMOD,
              hash_arg: int = 0, slots_per_id: int = 1, zero_pad: bool = False, pad_id: int = 0,
              pad_row: int = -1, bag_size: int = 0, lengths: Optional[torch.Tensor] = None,
              last_n: int = 0, buf: Optional[torch.Tensor] = None, ids_per_table: int = 0,
              num_tables: int = 0, shard_world: int = 1, shard_rank: int = 0,
              flip_len: int = 0, window=None, out_features: int = 0,
              out_feature_offset: int = 0) -> "BackwardPlan":
        """num_rows is rows PER TABLE; with ids_per_table > 0 the plan covers the stacked table
        of ceil(n_ids / ids_per_table) tables and `self.num_rows` is the stacked total."""
        flat = _flat_ids(ids)
        if lengths is not None:
            lengths = lengths.to(torch.int32).contiguous()
        dev = N.require_cuda(flat, lengths)
        n_slots = flat.numel() * slots_per_id
        lib = N.load()
        layout = N.make_layout(ids_per_table, num_tables, shard_world, shard_rank, flip_len, window, out_features,
                               out_feature_offset)
        total_rows = int(lib.recemb_layout_total_rows(num_rows, layout, flat.numel()))
        need = int(lib.recemb_bwd_plan_bytes(n_slots, total_rows))
        if need == 0:
            N.check(-2, "recemb_bwd_plan_bytes")
        if buf is None or buf.numel() < need:
            buf = torch.empty((need,), dtype=torch.uint8, device=flat.device)
        N.check(lib.recemb_bwd_plan(N.ptr(flat), flat.numel(), layout, slots_per_id, hash_mode, num_rows,
                                    hash_arg, int(zero_pad), pad_id, pad_row, bag_size,
                                    N.ptr(lengths), last_n, N.ptr(buf), buf.numel(), dev,
                                    N.stream_ptr(dev)), "recemb_bwd_plan")
        recipe = (ids, dict(num_rows=num_rows, hash_mode=hash_mode, hash_arg=hash_arg, slots_per_id=slots_per_id,
                            zero_pad=zero_pad, pad_id=pad_id, pad_row=pad_row, bag_size=bag_size, lengths=lengths,
                            last_n=last_n, ids_per_table=ids_per_table, num_tables=num_tables,
                            shard_world=shard_world, shard_rank=shard_rank, flip_len=flip_len, window=window,
                            out_features=out_features, out_feature_offset=out_feature_offset))
        # windowed / feature-interleaved: the slots map into a gradient tensor of another row count
        return BackwardPlan(buf=buf, n_slots=n_slots, num_rows=total_rows, slots_per_id=slots_per_id,
                            slots_cover_grad=window is None and not out_features, recipe=recipe)

    def _arr(self, which: int) -> torch.Tensor:
        arr_bytes = (self.n_slots * 4 + 255) // 256 * 256
        off = 256 + which * arr_bytes
        return self.buf[off:off + self.n_slots * 4].view(torch.int32)

    @property
    def sorted_rows(self) -> torch.Tensor:
        """int64 copy of the sorted row keys (num_rows marks a dropped slot)."""
        return self._arr(2).to(torch.int64) & 0xFFFFFFFF

    @property
    def sorted_slots(self) -> torch.Tensor:
        return self._arr(3).to(torch.int64) & 0xFFFFFFFF

    @property
    def counters(self) -> torch.Tensor:
        """device int64 [2] = (valid slots, distinct rows); counted on demand."""
        dev = N.require_cuda(self.buf)
        N.check(N.load().recemb_plan_count(N.ptr(self.buf), self.buf.numel(), self.n_slots,
                                           self.num_rows, dev, N.stream_ptr(dev)), "recemb_plan_count")
        return self.buf[:16].view(torch.int64)


def make_optim_params(lr: float = 0.0, eps: float = 0.0, weight_decay: float = 0.0,
                      beta1: float = 0.9, beta2: float = 0.999, step: int = 1,
                      grad_div: float = 0.0) -> N.OptimParams:
    return N.OptimParams(lr=lr, eps=eps, weight_decay=weight_decay, beta1=beta1, beta2=beta2,
                         bias_correction1=1.0 - beta1 ** step, bias_correction2=1.0 - beta2 ** step,
                         grad_div=grad_div)


def bwd_apply(plan: BackwardPlan, grad: torch.Tensor, *, table: torch.Tensor, update: int,
              slots_per_grad_row: int = 1, state1: Optional[torch.Tensor] = None,
              state2: Optional[torch.Tensor] = None, hp: Optional[N.OptimParams] = None,
              slot_weight: Optional[torch.Tensor] = None,
              grad_row_scale: Optional[torch.Tensor] = None,
              workspace: Optional[torch.Tensor] = None, grad_div: float = 0.0,
              guard: Optional[torch.Tensor] = None) -> None:
    """Segmented reduction of `grad` rows over the plan + `update` on `table` (in place).
    grad_div > 0: gradient elements are divided by it on the fly (k-shift 1/sqrt(k) backward).
    guard (int32 / uint32 device word): the update is skipped entirely when it is non-zero at kernel start
    (the peer exchange's status word: overflowed inbox / timed-out barrier -> the shard stays untouched)."""
    dim = table.shape[1]
    grad2d = grad.contiguous().view(-1, dim)
    if plan.slots_cover_grad and grad2d.shape[0] * slots_per_grad_row != plan.n_slots:
        raise N.NativeError(
            f"grad has {grad2d.shape[0]} rows x {slots_per_grad_row} slots, plan has {plan.n_slots}")
    dev = N.require_cuda(plan.buf, grad2d, table, state1, state2, slot_weight, grad_row_scale, guard)
    lib = N.load()
    need = int(lib.recemb_bwd_apply_workspace_bytes(plan.n_slots, dim))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty((need,), dtype=torch.uint8, device=table.device)
    hp = hp or make_optim_params()
    if grad_div:
        if slot_weight is not None or grad_row_scale is not None:
            raise N.NativeError("grad_div cannot be combined with slot_weight / grad_row_scale")
        hp = N.OptimParams.from_buffer_copy(hp)
        hp.grad_div = grad_div
    N.check(lib.recemb_bwd_apply_guarded(
        N.ptr(plan.buf), plan.buf.numel(), plan.n_slots, N.ptr(grad2d), N.dtype_code(grad2d.dtype),
        grad2d.shape[0], dim, slots_per_grad_row, N.ptr(slot_weight), N.ptr(grad_row_scale), update,
        N.ptr(table), N.dtype_code(table.dtype), table.shape[0], N.ptr(state1), N.ptr(state2),
        C.byref(hp), N.ptr(workspace), workspace.numel(), N.ptr(guard), dev, N.stream_ptr(dev)),
        "recemb_bwd_apply")


def epilogue_bwd(grad_out: torch.Tensor, out: Optional[torch.Tensor],
                 inv_norm: Optional[torch.Tensor], epilogue: int, num_shifts: int = 1) -> torch.Tensor:
    """fp32 [n, dim] gradient w.r.t. the pre-epilogue sum (autograd of commons/layers.py:167-170)."""
    dim = grad_out.shape[-1]
    g = grad_out.contiguous().view(-1, dim)
    o = None if out is None else out.contiguous().view(-1, dim)
    dev = N.require_cuda(g, o, inv_norm)
    dx = torch.empty(g.shape, dtype=torch.float32, device=g.device)
    N.check(N.load().recemb_epilogue_bwd(N.ptr(g), N.ptr(o), N.dtype_code(g.dtype), N.ptr(inv_norm),
                                         g.shape[0], dim, epilogue, num_shifts, N.ptr(dx), dev,
                                         N.stream_ptr(dev)), "recemb_epilogue_bwd")
    return dx


def shard_bucket(ids: torch.Tensor, *, num_rows: int, world: int, rank: int, bags_total: int,
                 lengths: Optional[torch.Tensor] = None, last_n: int = 0, zero_pad: bool = False,
                 pad_id: int = 0, bags_per_table: int = 0, num_tables: int = 0,
                 hash_mode: int = N.HASH_FLOORMOD) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sender side of the routed exchange: ids [bags, P] -> (entries int64 [bags * P] grouped by
    owner, counts int64 [world]); only the first counts.sum() entries are meaningful."""
    ids = ids.contiguous()
    if lengths is not None:
        lengths = lengths.to(torch.int32).contiguous()
    dev = N.require_cuda(ids, lengths)
    m, p = ids.shape
    lib = N.load()
    layout = N.Layout(ids_per_table=bags_per_table * p, num_tables=num_tables, shard_world=world,
                      shard_rank=rank, flip_len=0)
    entries = torch.empty((m * p,), dtype=torch.int64, device=ids.device)
    counts = torch.empty((world,), dtype=torch.int64, device=ids.device)
    ws = torch.empty((int(lib.recemb_shard_bucket_workspace_bytes(m * p, world)),), dtype=torch.uint8,
                     device=ids.device)
    N.check(lib.recemb_shard_bucket(N.ptr(ids), m * p, layout, hash_mode, num_rows, 0, int(zero_pad), pad_id,
                                    p, N.ptr(lengths), last_n, bags_total, N.ptr(entries), N.ptr(counts),
                                    N.ptr(ws), ws.numel(), dev, N.stream_ptr(dev)), "recemb_shard_bucket")
    return entries, counts


def pool_entries(table: torch.Tensor, entries: torch.Tensor, out_rows: int) -> torch.Tensor:
    """Owner side: received entries -> partial pools [out_rows, dim] (zero where no entry)."""
    entries = entries.contiguous()
    dev = N.require_cuda(table, entries)
    out = torch.zeros((out_rows, table.shape[1]), dtype=table.dtype, device=table.device)
    N.check(N.load().recemb_pool_entries(N.ptr(table), table.shape[1], N.dtype_code(table.dtype),
                                         N.ptr(entries), entries.numel(), N.ptr(out), dev,
                                         N.stream_ptr(dev)), "recemb_pool_entries")
    return out


def plan_from_entries(entries: torch.Tensor, total_rows: int) -> "BackwardPlan":
    """Owner side: received entries -> plan (key = local row, slot = gathered-gradient row)."""
    entries = entries.contiguous()
    dev = N.require_cuda(entries)
    lib = N.load()
    n = entries.numel()
    need = int(lib.recemb_bwd_plan_bytes(n, total_rows))
    if need == 0:
        N.check(-2, "recemb_bwd_plan_bytes")
    buf = torch.empty((need,), dtype=torch.uint8, device=entries.device)
    N.check(lib.recemb_bwd_plan_entries(N.ptr(entries), n, total_rows, N.ptr(buf), buf.numel(), dev,
                                        N.stream_ptr(dev)), "recemb_bwd_plan_entries")
    return BackwardPlan(buf=buf, n_slots=n, num_rows=total_rows, slots_per_id=1, slots_cover_grad=False)


def sum_partials(parts: torch.Tensor, row_scale: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[world, rows, dim] per-owner partial pools -> [rows, dim], fixed owner order, fp32 accumulate."""
    parts = parts.contiguous()
    dev = N.require_cuda(parts, row_scale, out)
    w, rows, dim = parts.shape
    if out is None:
        out = torch.empty((rows, dim), dtype=parts.dtype, device=parts.device)
    elif tuple(out.shape) != (rows, dim) or out.dtype != parts.dtype:
        raise N.NativeError("preallocated `out` has the wrong shape / dtype")
    N.check(N.load().recemb_sum_partials(N.ptr(parts), w, rows, dim, N.dtype_code(parts.dtype),
                                         N.ptr(row_scale), N.ptr(out), dev, N.stream_ptr(dev)),
            "recemb_sum_partials")
    return out


# ------------------------------------------------------------ peer exchange ----
def peer_pool_fwd(group, ids: torch.Tensor, *, num_rows: int, dim: int, dtype: torch.dtype,
                  lengths: Optional[torch.Tensor] = None, last_n: int = 0,
                  per_slot_weight: Optional[torch.Tensor] = None, hash_mode: int = N.HASH_FLOORMOD,
                  hash_arg: int = 0, pool_mode: int = N.POOL_SUM, zero_pad: bool = False, pad_id: int = 0,
                  bags_per_table: int = 0, num_tables: int = 0) -> torch.Tensor:
    """Pooled bags over a row-wise sharded table: every row is read from its owner's shard through
    peer-mapped memory (group: peer.PeerGroup).  Complete pools of THIS rank's bags, slot order."""
    if ids.dim() != 2 or ids.dtype != torch.int64:
        raise N.NativeError("pooled bags take int64 ids of shape [num_bags, bag_size]")
    ids = ids.contiguous()
    if lengths is not None:
        lengths = lengths.to(torch.int32).contiguous()
    if per_slot_weight is not None:
        per_slot_weight = per_slot_weight.to(torch.float32).contiguous()
    dev = N.require_cuda(ids, lengths, per_slot_weight)
    m, p = ids.shape
    out = torch.empty((m, dim), dtype=dtype, device=ids.device)
    N.check(N.load().recemb_peer_pool_fwd(
        C.byref(group.struct), num_rows, dim, N.dtype_code(dtype), N.ptr(ids), m, p, N.ptr(lengths), last_n,
        N.ptr(per_slot_weight), hash_mode, hash_arg, pool_mode, int(zero_pad), pad_id,
        N.make_layout(bags_per_table * p, num_tables), N.ptr(out), dev, N.stream_ptr(dev)),
        "recemb_peer_pool_fwd")
    return out


def peer_bucket_push(group, ids: torch.Tensor, *, num_rows: int, lengths: Optional[torch.Tensor] = None,
                     last_n: int = 0, zero_pad: bool = False, pad_id: int = 0, bags_per_table: int = 0,
                     num_tables: int = 0, hash_mode: int = N.HASH_FLOORMOD, hash_arg: int = 0,
                     tablewise: bool = False) -> None:
    """Sender side of the peer backward: my (local row, gradient row) entries land in the owners'
    inboxes, my bucket sizes in their count slots.  tablewise: table t lives whole on rank t % world."""
    ids = ids.contiguous()
    if lengths is not None:
        lengths = lengths.to(torch.int32).contiguous()
    dev = N.require_cuda(ids, lengths)
    m, p = ids.shape
    lib = N.load()
    layout = N.Layout(ids_per_table=bags_per_table * p, num_tables=num_tables, shard_world=group.world,
                      shard_rank=group.rank, flip_len=0, partition=int(tablewise))
    ws = torch.empty((int(lib.recemb_shard_bucket_workspace_bytes(m * p, group.world)),), dtype=torch.uint8,
                     device=ids.device)
    N.check(lib.recemb_peer_bucket_push(C.byref(group.struct), C.byref(group.layout), N.ptr(ids), m * p, layout,
                                        hash_mode, num_rows, hash_arg, int(zero_pad), pad_id, p, N.ptr(lengths),
                                        last_n, N.ptr(ws), ws.numel(), dev, N.stream_ptr(dev)),
            "recemb_peer_bucket_push")


def peer_bucket_push_rows(group, ids: torch.Tensor, *, num_rows: int, zero_pad: bool = False, pad_id: int = 0,
                          ids_per_table: int = 0, num_tables: int = 0, hash_mode: int = N.HASH_FLOORMOD,
                          hash_arg: int = 0) -> torch.Tensor:
    """Sequence mode, sender side (needs only the ids): my (local row, gradient slot) entries land in the
    owners' inboxes; returns dest int64 [n] = (owner << 32 | position) per lookup, -1 = dropped."""
    flat = _flat_ids(ids)
    dev = N.require_cuda(flat)
    n = flat.numel()
    lib = N.load()
    layout = N.Layout(ids_per_table=ids_per_table, num_tables=num_tables, shard_world=group.world,
                      shard_rank=group.rank, flip_len=0)
    ws = torch.empty((int(lib.recemb_shard_bucket_workspace_bytes(n, group.world)),), dtype=torch.uint8,
                     device=flat.device)
    dest = torch.empty((n,), dtype=torch.int64, device=flat.device)
    N.check(lib.recemb_peer_bucket_push_rows(C.byref(group.struct), C.byref(group.layout), N.ptr(flat), n, layout,
                                             hash_mode, num_rows, hash_arg, int(zero_pad), pad_id, N.ptr(dest),
                                             N.ptr(ws), ws.numel(), dev, N.stream_ptr(dev)),
            "recemb_peer_bucket_push_rows")
    return dest


def peer_rows_scatter_push(group, rows: torch.Tensor, dest: torch.Tensor) -> None:
    """Sequence mode, backward: gradient row i -> the slot of its owner's gradient buffer named by dest[i]."""
    rows = rows.contiguous()
    dev = N.require_cuda(rows, dest)
    if rows.dim() != 2 or rows.shape[0] != dest.numel():
        raise N.NativeError(f"rows {tuple(rows.shape)} do not match dest [{dest.numel()}]")
    N.check(N.load().recemb_peer_rows_scatter_push(C.byref(group.struct), C.byref(group.layout), N.ptr(rows),
                                                   rows.shape[0], rows.shape[1], N.dtype_code(rows.dtype),
                                                   N.ptr(dest), dev, N.stream_ptr(dev)),
            "recemb_peer_rows_scatter_push")


def peer_pool_push(group, dim: int, dtype: torch.dtype, tablewise_bags_per_table: int = 0) -> None:
    """Owner side of the push forward: pool the (sender, bag) runs of my inbox from my shard and
    store every partial row into the sender's parts region over NVLink.  tablewise_bags_per_table > 0:
    table-wise partitioning (zero rows only for empty bags of the tables this rank owns)."""
    if tablewise_bags_per_table:
        N.check(N.load().recemb_peer_pool_push_tablewise(
            C.byref(group.struct), C.byref(group.layout), dim, N.dtype_code(dtype), tablewise_bags_per_table,
            group.device, N.stream_ptr(group.device)), "recemb_peer_pool_push_tablewise")
        return
    N.check(N.load().recemb_peer_pool_push(C.byref(group.struct), C.byref(group.layout), dim, N.dtype_code(dtype),
                                           group.device, N.stream_ptr(group.device)), "recemb_peer_pool_push")


def peer_allgather_push(group, src: torch.Tensor, dst_offset: int) -> None:
    """src -> slice `rank` at dst_offset of every rank's arena (push all-gather over NVLink)."""
    src = src.contiguous()
    dev = N.require_cuda(src)
    N.check(N.load().recemb_peer_allgather_push(C.byref(group.struct), N.ptr(src), src.numel() * src.element_size(),
                                                dst_offset, dev, N.stream_ptr(dev)), "recemb_peer_allgather_push")


def peer_barrier(group, channel: int = 0) -> None:
    """Device-side barrier on the current stream; `channel` = independent flag set (one per stream)."""
    N.check(N.load().recemb_peer_barrier(C.byref(group.struct), C.byref(group.layout), channel, group.device,
                                         N.stream_ptr(group.device)), "recemb_peer_barrier")


def peer_bwd_apply_fused(plan: "BackwardPlan", my_grad: torch.Tensor, *, group, table: torch.Tensor, update: int,
                         state1: Optional[torch.Tensor], hp, tables: int, bags_per_table: int, rows_per_table: int,
                         push_ctas: int = 32, workspace: Optional[torch.Tensor] = None,
                         tablewise: bool = False) -> None:
    """Sharded backward, sender and owner side in ONE level-0 launch: the first push_ctas CTAs store my pooled
    gradients into every rank's buffer table by table, the others reduce + update my rows, each chunk gated on
    the flags of the last table it touches (recemb_peer_bwd_apply_fused)."""
    dim = table.shape[1]
    my_grad = my_grad.contiguous().view(-1, dim)
    if my_grad.dtype != table.dtype:
        raise N.NativeError("fused push: gradients must have the table's dtype")
    if my_grad.shape[0] != tables * bags_per_table:
        raise N.NativeError(f"my_grad has {my_grad.shape[0]} rows, expected {tables} x {bags_per_table}")
    dev = N.require_cuda(plan.buf, my_grad, table, state1)
    lib = N.load()
    need = int(lib.recemb_bwd_apply_workspace_bytes(plan.n_slots, dim))
    if workspace is None or workspace.numel() < need:
        workspace = torch.empty((need,), dtype=torch.uint8, device=table.device)
    fn = lib.recemb_peer_bwd_apply_fused_tablewise if tablewise else lib.recemb_peer_bwd_apply_fused
    N.check(fn(
        C.byref(group.struct), C.byref(group.layout), N.ptr(plan.buf), plan.buf.numel(), N.ptr(my_grad), tables,
        bags_per_table, dim, N.dtype_code(table.dtype), update, N.ptr(table), table.shape[0], rows_per_table,
        N.ptr(state1), C.byref(hp), N.ptr(workspace), workspace.numel(), push_ctas, dev, N.stream_ptr(dev)),
        "recemb_peer_bwd_apply_fused")


def peer_signal(group, channel: int) -> None:
    """First half of a barrier round on the current stream (never blocks): see recemb_peer_signal."""
    N.check(N.load().recemb_peer_signal(C.byref(group.struct), C.byref(group.layout), channel, group.device,
                                        N.stream_ptr(group.device)), "recemb_peer_signal")


def peer_wait(group, channel: int) -> None:
    """Second half: the current stream blocks until every rank has signalled this round of `channel`."""
    N.check(N.load().recemb_peer_wait(C.byref(group.struct), C.byref(group.layout), channel, group.device,
                                      N.stream_ptr(group.device)), "recemb_peer_wait")


def peer_plan(group, total_rows: int, buf: Optional[torch.Tensor] = None) -> "BackwardPlan":
    """Owner side (after the barrier): plan over my inbox, world * cap pairs, unused ones = sentinel.
    `buf`: plan memory kept by the caller from step to step (reallocated when too small)."""
    lib = N.load()
    n = group.world * int(group.layout.cap)
    need = int(lib.recemb_bwd_plan_bytes(n, total_rows))
    if need == 0:
        N.check(-2, "recemb_bwd_plan_bytes")
    if buf is None or buf.numel() < need or buf.device != group.arena.device:
        buf = torch.empty((need,), dtype=torch.uint8, device=group.arena.device)
    N.check(lib.recemb_peer_plan(C.byref(group.struct), C.byref(group.layout), total_rows, N.ptr(buf), buf.numel(),
                                 group.device, N.stream_ptr(group.device)), "recemb_peer_plan")
    return BackwardPlan(buf=buf, n_slots=n, num_rows=total_rows, slots_per_id=1, slots_cover_grad=False)


# --------------------------------------------------------- dot interaction ----
def dot_interaction_fwd(feats: torch.Tensor) -> torch.Tensor:
    if feats.dim() != 3 or feats.dtype != torch.bfloat16:
        raise N.NativeError("dot interaction takes bf16 [batch, num_feats, dim]")
    feats = feats.contiguous()
    dev = N.require_cuda(feats)
    b, f, d = feats.shape
    out = torch.empty((b, f * (f - 1) // 2), dtype=torch.bfloat16, device=feats.device)
    N.check(N.load().recemb_dot_interaction_fwd(N.ptr(feats), b, f, d, N.ptr(out), dev,
                                                N.stream_ptr(dev)), "recemb_dot_interaction_fwd")
    return out


def dot_interaction_bwd(feats: torch.Tensor, grad_out: torch.Tensor) -> torch.Tensor:
    feats = feats.contiguous()
    grad_out = grad_out.contiguous()
    dev = N.require_cuda(feats, grad_out)
    b, f, d = feats.shape
    gf = torch.empty_like(feats)
    N.check(N.load().recemb_dot_interaction_bwd(N.ptr(feats), N.ptr(grad_out), b, f, d, N.ptr(gf),
                                                dev, N.stream_ptr(dev)), "recemb_dot_interaction_bwd")
    return gf
