"""DLRM-style pairwise dot interaction for the ranker (BASELINE.json north_star item 3).

The reference names a `factorized_dlrm` ranker (models/ranker/config.py:16-61) but ships no
interaction code (models/ranker/fdlrm/* are empty files), so this is the canonical DLRM op:
T = [dense, e_1 .. e_F] (bf16 [B, F', D]); Z = T T^T; output = concat(dense, Z[tril(-1)]).
The Gram product runs on the tcgen05 tensor cores (csrc/interaction.cu)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _native as N
from . import ops
from .layers import _plan_early, _plan_take, _side_streams


class _DotInteractionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats):
        ctx.save_for_backward(feats)
        return ops.dot_interaction_fwd(feats)

    @staticmethod
    def backward(ctx, grad_out):
        (feats,) = ctx.saved_tensors
        return ops.dot_interaction_bwd(feats, grad_out.to(torch.bfloat16))


def dot_interaction(feats: torch.Tensor) -> torch.Tensor:
    """bf16 [B, F, D] -> bf16 [B, F(F-1)/2]: strictly-lower triangle of feats @ feats^T."""
    return _DotInteractionFn.apply(feats)


class DotInteraction(nn.Module):
    """forward(dense [B, D], sparse [B, F, D]) -> [B, D + (F+1)F/2] (dense passthrough first)."""

    def forward(self, dense: torch.Tensor, sparse: torch.Tensor) -> torch.Tensor:
        feats = torch.cat([dense.unsqueeze(1), sparse], dim=1).to(torch.bfloat16)
        return torch.cat([dense.to(torch.bfloat16), dot_interaction(feats)], dim=1)


class _PooledInteractionFn(torch.autograd.Function):
    """dense [B, D] + T pooled features -> interaction, without a copy in between: the table-batched pooled
    lookup writes its rows feature-interleaved straight into the interaction's [B, T + 1, D] input
    (recemb_layout.out_features) and the pooled backward reads the interaction's [B, T + 1, D] gradient in
    place through the same row mapping -- no torch.cat, no permute, no 113 MB (cfg 3) round trips."""

    @staticmethod
    def forward(ctx, dense, ids, lengths, coll, record, *anchors):
        t, n_rows = coll.num_tables, coll.num_embeddings
        stacked = coll._all.weight.detach()
        b, p = ids.shape[1], ids.shape[2]
        f = t + 1
        feats = torch.empty((b, f, coll.emb_dim), dtype=stacked.dtype, device=stacked.device)
        feats[:, 0].copy_(dense)
        flat_len = None if lengths is None else lengths.reshape(-1)
        ops.pool_fwd(stacked, ids.reshape(t * b, p), lengths=flat_len, last_n=coll.last_n, hash_mode=coll.hash_mode,
                     pool_mode=coll.pool_mode, zero_pad=coll.skip_pad, pad_id=coll.pad_id, num_rows=n_rows,
                     bags_per_table=b, out=feats, out_features=f, out_feature_offset=1)
        build = lambda: coll._all.build_plan(  # noqa: E731
            ids.reshape(t * b, p), num_rows=n_rows, hash_mode=coll.hash_mode, zero_pad=coll.skip_pad,
            pad_id=coll.pad_id, bag_size=p, lengths=flat_len, last_n=coll.last_n, ids_per_table=b * p,
            out_features=f, out_feature_offset=1)
        needs = record and any(ctx.needs_input_grad[5:])
        ctx.coll, ctx.build = coll, build
        ctx.save_for_backward(feats, ids, lengths)
        _plan_early(ctx, ids, needs, build)
        if ctx.plan is not None and lengths is not None:
            lengths.record_stream(_side_streams[ids.device])
        return ops.dot_interaction_fwd(feats)

    @staticmethod
    def backward(ctx, grad_out):
        coll = ctx.coll
        feats, ids, lengths = ctx.saved_tensors
        gfeats = ops.dot_interaction_bwd(feats, grad_out.to(torch.bfloat16))      # [B, T + 1, D]
        gdense = gfeats[:, 0] if ctx.needs_input_grad[0] else None
        gw = None
        if any(ctx.needs_input_grad[5:]):
            t, b, p = ids.shape
            plan = _plan_take(ctx, ctx.build)
            scale = None
            if coll.pool_mode == N.POOL_MEAN:
                from .layers import pooled_counts
                cnt = pooled_counts(ids.reshape(t * b, p), None if lengths is None else lengths.reshape(-1),
                                    coll.last_n, coll.skip_pad, coll.pad_id).clamp_(min=1).float()
                # grad_row_scale is indexed by the (interleaved) gradient row
                scale = torch.ones((b, t + 1), dtype=torch.float32, device=ids.device)
                scale[:, 1:] = (1.0 / cnt).view(t, b).t()
                scale = scale.reshape(-1)
            gw = coll._all.consume(plan, gfeats.view(-1, gfeats.shape[-1]), slots_per_grad_row=p, grad_row_scale=scale)
        if gw is None:
            return (gdense, None, None, None, None) + (None,) * coll.num_anchors
        n = coll.num_embeddings
        return (gdense, None, None, None, None) + tuple(gw[i * n:(i + 1) * n] for i in range(coll.num_tables))


class PooledInteraction(nn.Module):
    """The ranker's sparse front end as one pipeline: forward(dense [B, D] bf16, ids [T, B, P], lengths [T, B])
    -> [B, D + (T+1)T/2] = concat(dense, lower triangle of [dense, pooled_1 .. pooled_T] Gram matrix), the same
    result as DotInteraction()(dense, collection(ids, lengths).permute(1, 0, 2)) without the intermediate
    [T, B, D] tensor, its permute and the concatenation.  `collection`: a bf16 EmbeddingCollection(kind="pooled")."""

    def __init__(self, collection):
        super().__init__()
        if collection.kind != "pooled" or collection._all.weight.dtype != torch.bfloat16:
            raise ValueError("PooledInteraction needs a bf16 pooled EmbeddingCollection")
        if collection.num_tables + 1 > 32 or collection.emb_dim % 64:
            raise ValueError("the tcgen05 interaction takes <= 32 features of a dim that is a multiple of 64")
        self.collection = collection

    def forward(self, dense: torch.Tensor, ids, lengths=None) -> torch.Tensor:
        coll = self.collection
        if isinstance(ids, (list, tuple)):
            ids = torch.stack(list(ids))
        if isinstance(lengths, (list, tuple)):
            lengths = torch.stack(list(lengths))
        if ids.dim() != 3 or ids.shape[0] != coll.num_tables:
            raise N.NativeError(f"ids must be [T = {coll.num_tables}, B, P]")
        dense = dense.to(torch.bfloat16)
        anchors = (coll._all.grad_anchor(),) if coll._all.fused is not None else \
            tuple(tab.weight for tab in coll.members())
        inter = _PooledInteractionFn.apply(dense, ids.contiguous(), lengths, coll, torch.is_grad_enabled(), *anchors)
        return torch.cat([dense, inter], dim=1)
