"""YAML config surface of the embedding path.

Keeps the reference's field names verbatim so an existing `hydra-configs/model/*.yaml` drives the
B200 modules: `features.defaults.categorical_features.embedding.{num_embeddings, emb_dim, use_qr}`,
`proj_dim`, `features.embedding_tables.<name>`, `features.defaults.embedding_table_config.
{shared,query,item}.<name>`, `categorical_history_features[].{name, history_length,
history_id_feature_name, emb_table_name, remove_history_id_from_history}`,
`product_tower.latent_model_config.{vocab_size_latent, num_shifts_latent, normalize_embedding}`,
`product_tower.cosine_lsh_config[].{num_bins, num_proj}`, `sparse`
(commons/configs/feature_config.py:12-16, 54-59, 107-125, 325-361, 456-472;
hydra-configs/model/lthm.yaml:5, 19-36, 68-94; models/lthm/sequence/encoder.py:32-37).

hydra / omegaconf are not needed: the YAML is read with PyYAML and `${eval: ...}` interpolations are
resolved with Python `eval`, exactly the semantics of the reference's resolver
(commons/configs/trainer_pipeline_config.py:63-65).  Registry / loader mechanics of the reference
(pydantic polymorphic dispatch, Ray, trackers) are out of scope.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import torch
import yaml

from .layers import CosineVectorEmbedding, FlatEmbedding, KShiftEmbedding, QREmbedding
from .table import FusedOptimizerConfig

_EVAL = re.compile(r"^\$\{eval:\s*(.*)\}$")


def resolve(node: Any) -> Any:
    """Resolve `${eval: expr}` strings (OmegaConf resolver `eval` == Python eval) recursively."""
    if isinstance(node, dict):
        return {k: resolve(v) for k, v in node.items()}
    if isinstance(node, list):
        return [resolve(v) for v in node]
    if isinstance(node, str):
        m = _EVAL.match(node.strip())
        if m:
            return eval(m.group(1), {"__builtins__": {}}, {})  # arithmetic only, as in the reference's YAML
    return node


@dataclass
class EmbeddingTable:
    """commons/configs/feature_config.py:12-16."""
    num_embeddings: int
    emb_dim: int
    use_qr: bool = False


@dataclass
class CategoricalHistoryFeature:
    """commons/configs/feature_config.py:325-361 (fields of the embedding path only)."""
    name: str
    history_length: int = 20
    history_id_feature_name: Optional[str] = None
    emb_table_name: Optional[str] = None
    remove_history_id_from_history: bool = False


@dataclass
class LatentModelConfig:
    """product_tower.latent_model_config, read at models/lthm/sequence/encoder.py:32-37."""
    vocab_size_latent: int
    num_shifts_latent: int = 8
    normalize_embedding: bool = False


@dataclass
class EmbeddingPathConfig:
    default_table: Optional[EmbeddingTable] = None
    proj_dim: int = 0
    embedding_tables: Dict[str, EmbeddingTable] = field(default_factory=dict)
    table_groups: Dict[str, Dict[str, EmbeddingTable]] = field(default_factory=dict)  # shared / query / item
    history_features: List[CategoricalHistoryFeature] = field(default_factory=list)
    latent: Optional[LatentModelConfig] = None
    cosine_lsh: List[Dict[str, int]] = field(default_factory=list)
    inp_emb_dim: int = 32
    out_emb_dim: int = 512
    sparse: bool = False
    sparse_learning_rate: float = 0.25  # commons/configs/trainer_config.py:110

    @staticmethod
    def from_dict(cfg: Dict[str, Any]) -> "EmbeddingPathConfig":
        cfg = resolve(cfg)
        model = cfg.get("model", cfg)
        feats = model.get("features", {}) or {}
        defaults = feats.get("defaults", {}) or {}
        cat = defaults.get("categorical_features", {}) or {}
        out = EmbeddingPathConfig(sparse=bool(model.get("sparse", False)))
        if cat.get("embedding"):
            out.default_table = EmbeddingTable(**cat["embedding"])
        out.proj_dim = int(cat.get("proj_dim", 0) or 0)
        for name, t in (feats.get("embedding_tables") or {}).items():
            out.embedding_tables[name] = EmbeddingTable(**t)
        for group, tables in (defaults.get("embedding_table_config") or {}).items():
            if tables:
                out.table_groups[group] = {n: EmbeddingTable(**t) for n, t in tables.items()}
        keep = CategoricalHistoryFeature.__dataclass_fields__
        for f in feats.get("categorical_history_features") or []:
            out.history_features.append(CategoricalHistoryFeature(**{k: v for k, v in f.items() if k in keep}))
        tower = model.get("product_tower", {}) or {}
        if tower.get("latent_model_config"):
            out.latent = LatentModelConfig(**tower["latent_model_config"])
        out.cosine_lsh = list(tower.get("cosine_lsh_config") or [])
        out.inp_emb_dim = int(tower.get("inp_emb_dim", out.inp_emb_dim))
        out.out_emb_dim = int(tower.get("out_emb_dim", out.out_emb_dim))
        train = cfg.get("train", {}) or {}
        if "sparse_learning_rate" in train:
            out.sparse_learning_rate = float(train["sparse_learning_rate"])
        return out

    @staticmethod
    def from_yaml(path_or_text: str) -> "EmbeddingPathConfig":
        text = path_or_text
        if "\n" not in path_or_text and path_or_text.endswith((".yaml", ".yml")):
            with open(path_or_text) as fh:
                text = fh.read()
        return EmbeddingPathConfig.from_dict(yaml.safe_load(text))

    # ------------------------------------------------------------- builders ----
    def table(self, name: Optional[str] = None) -> EmbeddingTable:
        if name:
            if name in self.embedding_tables:
                return self.embedding_tables[name]
            for group in self.table_groups.values():
                if name in group:
                    return group[name]
            raise KeyError(f"embedding table {name!r} is not configured")
        if self.default_table is None:
            raise KeyError("no default categorical embedding configured")
        return self.default_table

    def build_table_module(self, name: Optional[str] = None, *, device=None, dtype=torch.float32,
                           fused: Optional[FusedOptimizerConfig] = None, normalize_output: bool = False):
        """use_qr -> QREmbedding (commons/layers.py:102-123) else FlatEmbedding (:44-61)."""
        t = self.table(name)
        if t.use_qr:
            return QREmbedding(t.num_embeddings, t.emb_dim, normalize_output, device=device, dtype=dtype,
                               fused_optimizer=fused)
        return FlatEmbedding(t.num_embeddings, t.emb_dim, normalize_output=normalize_output, device=device,
                             dtype=dtype, sparse=self.sparse, fused_optimizer=fused)

    def build_product_embedding(self, *, device=None, dtype=torch.float32,
                                fused: Optional[FusedOptimizerConfig] = None) -> KShiftEmbedding:
        """The fallback product embedding of Encoder (models/lthm/sequence/encoder.py:32-37), built with
        inp_emb_dim (the dimension ProductTower.emb_mapper consumes, product_tower.py:19, :52)."""
        if self.latent is None:
            raise KeyError("product_tower.latent_model_config is not configured")
        return KShiftEmbedding(self.latent.vocab_size_latent, self.inp_emb_dim,
                               num_shifts=self.latent.num_shifts_latent,
                               normalize_output=self.latent.normalize_embedding, sparse=self.sparse,
                               device=device, dtype=dtype, fused_optimizer=fused)

    def build_direction_embeddings(self, *, device=None):
        """ProductTower.direction_emb (product_tower.py:21-29): one CosineVectorEmbedding per entry."""
        return torch.nn.ModuleList([
            CosineVectorEmbedding(self.inp_emb_dim, self.out_emb_dim, n_proj=int(c["num_proj"]),
                                  num_bins=int(c["num_bins"]), device=device) for c in self.cosine_lsh])
