"""EmbeddingCollection: T equal-shaped tables behind ONE launch per phase.

The reference holds its tables as independent nn.Modules (FlatEmbedding: commons/layers.py:44-61,
callers models/lthm/sequence/query_tower.py:24, :53; one nn.EmbeddingBag per sparse feature in a
ranker) and therefore pays one gather, one sort and one optimizer pass per table.  Here the T
tables live stacked in one [T * N, D] allocation; a forward is ONE gather (or pooled-bag) kernel
over all tables, a backward ONE plan + ONE segmented reduction (table-batched mode of the C ABI:
recemb_layout.ids_per_table).  Nothing changes for checkpoints: every table is still a child module
with the reference's key (`<name>._emb_table.weight` / `<name>.emb.weight`), whose tensor is a VIEW
of the stacked storage, so state_dict() / load_state_dict() see T independent tables.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Union

import torch
import torch.nn as nn

from . import _native as N
from . import ops
from .layers import _plan_early, _plan_take, _side_streams, pooled_counts
from .table import EmbeddingTable, FusedOptimizerConfig


class _CollectionFn(torch.autograd.Function):
    """ids [T, ...] -> out [T, ..., D] (flat) or [T, B, D] (pooled); weights are the T per-table views
    (torch-compatible mode) or nothing but the anchor (fused mode)."""

    @staticmethod
    def forward(ctx, ids, lengths, coll, record, *anchors):
        t, n_rows = coll.num_tables, coll.num_embeddings
        stacked = coll._all.weight.detach()
        needs = record and any(ctx.needs_input_grad[4:])  # record = torch.is_grad_enabled() at the call site
        if coll.kind == "flat":
            per_table = ids[0].numel()
            flip_len = int(ids.shape[-1]) if coll.flip_sequences else 0
            out, _ = ops.gather_fwd(stacked, ids, zero_pad=coll.fused_pad_mask, pad_id=0, ids_per_table=per_table,
                                    flip_len=flip_len)
            build = lambda: coll._all.build_plan(  # noqa: E731
                ids, num_rows=n_rows, zero_pad=coll.fused_pad_mask, pad_id=0, ids_per_table=per_table,
                flip_len=flip_len)
        else:
            b, p = ids.shape[1], ids.shape[2]
            flat_len = None if lengths is None else lengths.reshape(-1)
            out = ops.pool_fwd(stacked, ids.reshape(t * b, p), lengths=flat_len, last_n=coll.last_n,
                               hash_mode=coll.hash_mode, pool_mode=coll.pool_mode, zero_pad=coll.skip_pad,
                               pad_id=coll.pad_id, num_rows=n_rows, bags_per_table=b).view(t, b, -1)
            build = lambda: coll._all.build_plan(  # noqa: E731
                ids.reshape(t * b, p), num_rows=n_rows, hash_mode=coll.hash_mode, zero_pad=coll.skip_pad,
                pad_id=coll.pad_id, bag_size=p, lengths=flat_len, last_n=coll.last_n, ids_per_table=b * p)
        ctx.coll, ctx.build = coll, build
        ctx.save_for_backward(ids, lengths)
        _plan_early(ctx, ids, needs, build)
        if ctx.plan is not None and lengths is not None:
            lengths.record_stream(_side_streams[ids.device])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        coll = ctx.coll
        ids, lengths = ctx.saved_tensors
        plan = _plan_take(ctx, ctx.build)
        dim = grad_out.shape[-1]
        g2d = grad_out.contiguous().view(-1, dim)
        if coll.kind == "flat":
            gw = coll._all.consume(plan, g2d)
        else:
            t, b, p = ids.shape
            scale = None
            if coll.pool_mode == N.POOL_MEAN:
                scale = 1.0 / pooled_counts(ids.reshape(t * b, p), None if lengths is None else lengths.reshape(-1),
                                            coll.last_n, coll.skip_pad, coll.pad_id).clamp_(min=1).float()
            gw = coll._all.consume(plan, g2d, slots_per_grad_row=p, grad_row_scale=scale)
        if gw is None:  # fused: updated in place
            return (None, None, None, None) + (None,) * coll.num_anchors
        n = coll.num_embeddings
        return (None, None, None, None) + tuple(gw[i * n:(i + 1) * n] for i in range(coll.num_tables))


class _Member(nn.Module):
    """One table of the collection under the reference's module layout: FlatEmbedding keeps its table
    as `_emb_table` (commons/layers.py:51), the bag modules as `emb` (commons/transformers/layers.py:457)."""

    def __init__(self, attr: str, table: EmbeddingTable):
        super().__init__()
        setattr(self, attr, table)


class EmbeddingCollection(nn.Module):
    """T tables of [num_embeddings, emb_dim] served by table-batched launches.

    kind="flat"    FlatEmbedding semantics per table (row = floor_mod(id, N); sequence gather):
                   forward(ids [T, ...]) -> [T, ..., emb_dim]; state_dict keys `<name>._emb_table.weight`
    kind="pooled"  PooledEmbeddingBag semantics per table (sum / mean, per-bag lengths, last_n):
                   forward(ids [T, B, P], lengths [T, B]) -> [T, B, emb_dim]; keys `<name>.emb.weight`
    `ids` may also be a list of T equal-shaped tensors.  Gradient modes as in EmbeddingTable:
    torch-compatible (T Parameters, each gets its dense .grad) or fused (one in-kernel update)."""

    def __init__(self, names: Union[int, Sequence[str]], num_embeddings: int, emb_dim: int, *, kind: str = "flat",
                 mode: str = "sum", last_n: int = 0, hash_ids: bool = True, skip_pad: bool = False, pad_id: int = 0,
                 fused_pad_mask: bool = False, flip_sequences: bool = False, dtype: torch.dtype = torch.float32,
                 device=None, fused_optimizer: Optional[FusedOptimizerConfig] = None):
        super().__init__()
        if kind not in ("flat", "pooled"):
            raise ValueError("kind must be 'flat' or 'pooled'")
        if mode not in ("sum", "mean"):
            raise ValueError("mode must be 'sum' or 'mean'")
        self.names: List[str] = [f"table_{i}" for i in range(names)] if isinstance(names, int) else list(names)
        self.num_tables, self.num_embeddings, self.emb_dim, self.kind = len(self.names), int(num_embeddings), emb_dim, kind
        self.fused_pad_mask, self.flip_sequences = fused_pad_mask, flip_sequences
        self.last_n, self.skip_pad, self.pad_id = int(last_n), skip_pad, pad_id
        self.hash_mode = N.HASH_FLOORMOD if hash_ids else N.HASH_IDENTITY
        self.pool_mode = N.POOL_SUM if mode == "sum" else N.POOL_MEAN
        attr = "_emb_table" if kind == "flat" else "emb"
        t, n = self.num_tables, self.num_embeddings
        # the stacked storage; every member table is initialised exactly like nn.Embedding would be
        stacked = torch.empty((t * n, emb_dim), dtype=dtype, device=device)
        members = []
        for i in range(t):
            tab = EmbeddingTable(n, emb_dim, dtype=dtype, device=device)
            stacked[i * n:(i + 1) * n].copy_(tab.weight.detach())
            members.append(tab)
        self.tables = nn.ModuleDict({name: _Member(attr, tab) for name, tab in zip(self.names, members)})
        self._attr = attr
        object.__setattr__(self, "_all", EmbeddingTable(t * n, emb_dim, dtype=dtype, device=device, _weight=stacked))
        self._restack(stacked)
        if fused_optimizer is not None:
            self.enable_fused_optimizer(fused_optimizer)

    # ------------------------------------------------------------- storage ----
    def members(self) -> List[EmbeddingTable]:
        return [getattr(self.tables[name], self._attr) for name in self.names]

    def _restack(self, stacked: torch.Tensor) -> None:
        """Point the stacked holder and every member table at `stacked` (members become views)."""
        n = self.num_embeddings
        all_ = self._all
        fused = all_.fused is not None
        if fused:
            all_._buffers["weight"] = stacked
        else:
            all_._parameters["weight"] = nn.Parameter(stacked, requires_grad=False)
        for i, tab in enumerate(self.members()):
            view = stacked[i * n:(i + 1) * n]
            if fused:
                tab._parameters.pop("weight", None)
                tab._buffers["weight"] = view
            else:
                tab._buffers.pop("weight", None)
                tab._parameters["weight"] = nn.Parameter(view)

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() move every member tensor separately, which would un-stack them:
        # move the stacked storage once and re-create the views
        stacked = fn(self._all.weight.detach())
        states = {k: fn(v) for k, v in self._all._buffers.items() if k.startswith("opt_state") and v is not None}
        super()._apply(fn, *args, **kwargs)
        self._restack(stacked)
        for k, v in states.items():
            self._all._buffers[k] = v
        self._share_state()
        return self

    def enable_fused_optimizer(self, config: Optional[FusedOptimizerConfig] = None, **kw) -> "EmbeddingCollection":
        cfg = config or FusedOptimizerConfig(**kw)
        stacked = self._all.weight.detach()
        self._all.enable_fused_optimizer(cfg)
        for tab in self.members():
            tab.fused = cfg
        self._restack(stacked)
        self._all._ensure_state()
        self._share_state()
        return self

    def _share_state(self) -> None:
        """Per-table optimizer state = views of the stacked state (checkpoints stay per table)."""
        n = self.num_embeddings
        for name in ("opt_state1", "opt_state2"):
            st = self._all._buffers.get(name)
            if st is None:
                continue
            for i, tab in enumerate(self.members()):
                tab._buffers[name] = st[i * n:(i + 1) * n]
                tab._non_persistent_buffers_set.add(name)

    @property
    def table(self) -> EmbeddingTable:
        """The stacked table (hand this to FusedEmbeddingOptimizer in fused mode)."""
        return self._all

    @property
    def num_anchors(self) -> int:
        return 1 if self._all.fused is not None else self.num_tables

    # -------------------------------------------------------------- forward ----
    def forward(self, ids, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        if isinstance(ids, (list, tuple)):
            ids = torch.stack(list(ids))
        if ids.shape[0] != self.num_tables:
            raise N.NativeError(f"ids carry {ids.shape[0]} tables, the collection has {self.num_tables}")
        if self.kind == "pooled" and ids.dim() != 3:
            raise N.NativeError("pooled collections take ids of shape [T, num_bags, bag_size]")
        if isinstance(lengths, (list, tuple)):
            lengths = torch.stack(list(lengths))
        ids = ids.contiguous()
        anchors = (self._all.grad_anchor(),) if self._all.fused is not None else \
            tuple(tab.weight for tab in self.members())
        return _CollectionFn.apply(ids, lengths, self, torch.is_grad_enabled(), *anchors)
