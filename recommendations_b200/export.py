"""TorchScript export of the embedding modules.

The reference exports its product-embedding module with torch.jit.script + jit.save
(embedding_module_gen.py:186-196: ModelWrapper(model, mask_model)) and loads it back with torch.jit.load
inside the LTHM encoder (models/lthm/sequence/encoder.py:25-29).  The training modules of this package call
the C ABI through ctypes inside autograd.Functions, which TorchScript cannot compile; their SCRIPTABLE twins
below call the same kernels through operators registered with TORCH_LIBRARY
(csrc_torch/torch_ops.cpp -> lib/librecemb_torch_ops.so: recemb_b200::kshift_fwd / gather_fwd / pool_fwd).
The twins keep the reference's state_dict keys, so `scriptable(module)` shares the trained weights.

A process that torch.jit.load()s such an archive must load the operator library first:
    import recommendations_b200.export as X; X.load_ops()      # or torch.ops.load_library(X.OPS_PATH)
"""
from __future__ import annotations

from pathlib import Path

import torch
import torch.nn as nn

from . import _native as N

OPS_PATH = Path(__file__).resolve().parent / "lib" / "librecemb_torch_ops.so"
_loaded = False


def load_ops() -> None:
    """Registers the recemb_b200::* operators with this process's torch dispatcher (idempotent)."""
    global _loaded
    if _loaded:
        return
    if not OPS_PATH.exists():
        raise N.NativeLibraryMissing(f"{OPS_PATH} is missing: build it with `python -m recommendations_b200.build_native`")
    torch.ops.load_library(str(OPS_PATH))
    _loaded = True


class _Weight(nn.Module):
    """Holder that keeps the reference's `<name>.weight` state_dict key."""

    def __init__(self, weight: torch.Tensor):
        super().__init__()
        self.register_buffer("weight", weight)


class ScriptableKShiftEmbedding(nn.Module):
    """KShiftEmbedding.forward (commons/layers.py:152-172) as one registered operator; key `emb.weight`."""

    def __init__(self, weight: torch.Tensor, num_shifts: int, normalize_output: bool, flip_sequences: bool = False):
        super().__init__()
        load_ops()
        self.emb = _Weight(weight)
        self._num_shifts = int(num_shifts)
        self._normalize_output = bool(normalize_output)
        self._flip_sequences = bool(flip_sequences)

    def forward(self, id_: torch.Tensor) -> torch.Tensor:
        flip_len = id_.size(-1) if self._flip_sequences else 0
        return torch.ops.recemb_b200.kshift_fwd(self.emb.weight, id_, self._num_shifts, self._normalize_output, flip_len)


class ScriptableFlatEmbedding(nn.Module):
    """FlatEmbedding.forward (commons/layers.py:56-61); key `_emb_table.weight`."""

    def __init__(self, weight: torch.Tensor, normalize_output: bool, fused_pad_mask: bool = False,
                 flip_sequences: bool = False):
        super().__init__()
        load_ops()
        self._emb_table = _Weight(weight)
        self._normalize_output = bool(normalize_output)
        self._fused_pad_mask = bool(fused_pad_mask)
        self._flip_sequences = bool(flip_sequences)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        flip_len = x.size(-1) if self._flip_sequences else 0
        return torch.ops.recemb_b200.gather_fwd(self._emb_table.weight, x, self._normalize_output, self._fused_pad_mask,
                                                flip_len)


class ScriptablePooledEmbeddingBag(nn.Module):
    """Fixed-size pooled bags (the op behind nn.EmbeddingBag, commons/transformers/layers.py:457); key `emb.weight`."""

    def __init__(self, weight: torch.Tensor, mean: bool, hash_ids: bool):
        super().__init__()
        load_ops()
        self.emb = _Weight(weight)
        self._mean = bool(mean)
        self._hash_ids = bool(hash_ids)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        return torch.ops.recemb_b200.pool_fwd(self.emb.weight, ids, self._hash_ids, self._mean)


def scriptable(module: nn.Module) -> nn.Module:
    """The scriptable inference twin of a training module of this package, SHARING its weight tensor."""
    from . import layers as L
    if isinstance(module, L.KShiftEmbedding):
        return ScriptableKShiftEmbedding(module.emb.weight.detach(), module._num_shifts, module._normalize_output,
                                         module._flip_sequences)
    if isinstance(module, L.FlatEmbedding):
        return ScriptableFlatEmbedding(module._emb_table.weight.detach(), module._normalize_output,
                                       module._fused_pad_mask, module._flip_sequences)
    if isinstance(module, L.PooledEmbeddingBag):
        if module.last_n or module.skip_pad:
            raise N.NativeError("export covers fixed-size bags (no last_n window / pad skipping)")
        return ScriptablePooledEmbeddingBag(module.emb.weight.detach(), module.mode == "mean", module.hash_ids)
    if isinstance(module, nn.Sequential):
        return nn.Sequential(*[scriptable(m) for m in module])
    return module  # dense torch modules script as they are


class ModelWrapper(nn.Module):
    """embedding_module_gen.py:31-41: emb * sigmoid(mask_model(ids)) -- the module the reference scripts."""

    def __init__(self, model: nn.Module, mask_model: nn.Module):
        super().__init__()
        self.model = scriptable(model)
        self.mask_model = scriptable(mask_model)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        emb = self.model(x)
        mask = self.mask_model(x).sigmoid()
        return mask * emb
