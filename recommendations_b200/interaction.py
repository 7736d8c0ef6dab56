"""DLRM-style pairwise dot interaction for the ranker (BASELINE.json north_star item 3).

The reference names a `factorized_dlrm` ranker (models/ranker/config.py:16-61) but ships no
interaction code (models/ranker/fdlrm/* are empty files), so this is the canonical DLRM op:
T = [dense, e_1 .. e_F] (bf16 [B, F', D]); Z = T T^T; output = concat(dense, Z[tril(-1)]).
The Gram product runs on the tcgen05 tensor cores (csrc/interaction.cu)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


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
