"""Repaired harness of the reference's LTHM training step (BASELINE.json configs[0]) -- TEST
INFRASTRUCTURE.

`main_training.py + hydra-configs/lthm_train.yaml` cannot run as shipped (about 30
wiring defects, ray / hydra / accelerate absent).  This module restates ONE training step of it --
Encoder.forward (models/lthm/sequence/encoder.py:44-61), ProductTower.forward (product_tower.py:43-62),
QueryTower.forward (query_tower.py:60-137), LTHMModelWrapper._train_or_val_step_helper (wrapper.py:114-245),
the loop body `loss += 0.0 * sum |p|; backward; AdamW.step` (accelerate_training_strategy.py:353-368,
wrapper.py:255-275) -- parameterised by a LAYER FLAVOUR:

  "reference"  the reference's own classes imported from /root/reference (build container only):
               commons.layers.{KShiftEmbedding, FlatEmbedding}, commons.transformers.layers.
               {CosineVectorEmbedding, TransformerBlock}, models.lthm.sequence.query_tower.QueryTower,
               commons.layers.CascadedStreamingLogQCorrectionModule
  "oracle"     plain-torch restatements of the same classes (run anywhere, CPU)
  "b200"       recommendations_b200 modules swapped in (CUDA)

All flavours share parameter names, so one state_dict loads into each: that is the drop-in claim.

Repairs (every one is needed only to make the reference execute; none changes behaviour):
  R1  import paths: encoder.py:9-11 imports `lthm.*` (package is models.lthm); product_tower.py:6 imports
      the non-existent commons.layers.HistogramEmbedding -> Encoder / ProductTower / wrapper are restated
      here line by line instead of imported; norm_bins = 1 disables the histogram branch
      (product_tower.py:31, :56).
  R2  config: plain dataclasses stand in for LTHMModelConfig / ProductTowerConfig, which lack the fields
      the towers read (inp_emb_dim, out_emb_dim, norm_threshold, cosine_lsh_config, log_q_config, ...).
  R3  product_tower.py:25 passes `num_proj=`; the constructor keyword is `n_proj`
      (commons/transformers/layers.py:444).
  R4  encoder.py:32-37 builds the fallback KShiftEmbedding with out_emb_dim, but ProductTower.emb_mapper
      expects inp_emb_dim (product_tower.py:19, :52) -> inp_emb_dim.
  R5  commons/layers.py:28-37 PatternFromTimelocal: no super().__init__(), nn.Embedding(emb_dim=...)
      -> repaired subclass patched into query_tower before QueryTower is constructed.
  R6  query_tower.py:17 sets `self.ememb_dim`, :108 reads `self.emb_dim` -> attribute added.
  R7  query_tower.py:129 returns `current_token_ids`, wrapper.py:132 reads `current_token_id` -> the
      plural key is used.  encoder.py:53 reads batch["timestamp"]: the batch carries that key.
  R8  wrapper.py:19 sets `self.model_config`, :82 reads `self._model_config` -> train_mini_batch_size < 0
      path (whole batch at once).
  R9  commons/layers.py:213 `self.alpha[hash] = batch_idx` (alpha is a float) -> `self.a[hash] = ...`;
      :236-237 `for mod in enumerate(...)` / `train_Step` -> iterate the modules, call train_step.
  R10 dropout = 0 (deterministic comparison); python `random` is seeded before the lookahead offsets are
      drawn (wrapper.py:162).
"""
from __future__ import annotations

import math
import random
import sys
from dataclasses import dataclass, field
from pathlib import Path
from types import SimpleNamespace
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

REF = Path("/root/reference")


# ------------------------------------------------------------------------------- config ----
@dataclass
class HarnessConfig:
    batch: int = 256
    hist: int = 50
    max_valid: int = 42          # valid lengths ~ U{1..max_valid}: leaves all-pad columns so the trim acts
    vocab: int = 10_000          # KShiftEmbedding(10 000, 32, num_shifts=8)  (SURVEY section 8d cfg 1)
    num_shifts: int = 8
    normalize_embedding: bool = False
    inp_emb_dim: int = 32
    out_emb_dim: int = 64
    product_emb_dim: int = 16
    norm_threshold: float = 0.05
    cosine: Tuple[Tuple[int, int], ...] = ((4, 8), (12, 8), (20, 8))  # (num_bins, n_proj)
    n_embd: int = 64
    n_head: int = 4
    num_layers: int = 2
    ff_mult: int = 4
    is_causal: bool = True
    context_width: int = 64
    lookahead: Tuple[int, ...] = (0, 5)
    softmax_temperature: float = 1.0
    logq_buckets: int = 4096
    logq_offsets: Tuple[int, ...] = (0, 34144, 7465477)
    logq_alpha: float = 0.05
    logq_p_init: float = 0.001
    logq_beta: float = 0.5
    lr: float = 6e-4
    weight_decay: float = 0.01
    betas: Tuple[float, float] = (0.9, 0.95)
    seed: int = 7

    @property
    def export_tokens(self) -> int:
        return len(self.lookahead)

    @property
    def export_span(self) -> int:
        return max(self.lookahead) + 1

    def model_config(self):
        """The attribute tree the reference modules read (R2)."""
        attn = SimpleNamespace(n_embd=self.n_embd, n_head=self.n_head, attn_dropout=0.0, dropout=0.0, bias=True,
                               pos_bias=None)
        tcfg = SimpleNamespace(rotator_config=self.ff_mult, is_causal=self.is_causal, max_block_size=None,
                               is_sparse_attn=False, sparsity_factor=0.5, enable_gradient_checkpointing=False,
                               attn_config=attn, dropout=0.0, num_layers=self.num_layers)
        tower = SimpleNamespace(inp_emb_dim=self.inp_emb_dim, out_emb_dim=self.out_emb_dim,
                                product_emb_dim=self.product_emb_dim, norm_threshold=self.norm_threshold,
                                norm_bins=1, model_init_metadata=None,
                                cosine_lsh_config=[SimpleNamespace(num_bins=b, num_proj=p) for b, p in self.cosine],
                                latent_model_config=SimpleNamespace(vocab_size_latent=self.vocab,
                                                                    num_shifts_latent=self.num_shifts,
                                                                    normalize_embedding=self.normalize_embedding))
        return SimpleNamespace(transformer_config=tcfg, emb_dim=self.n_embd, context_width=self.context_width,
                               product_tower=tower, export_tokens=self.export_tokens, export_span=self.export_span)


# ------------------------------------------------------------------ synthetic batch ----
def make_batch(cfg: HarnessConfig) -> Dict[str, torch.Tensor]:
    """SURVEY section 8d cfg 1: per row a valid length, ids uniform over the signed 64-bit range for the
    valid positions, 0 right-padding; labels in {0..3}; timestamps uniform epoch seconds."""
    g = torch.Generator().manual_seed(cfg.seed)
    lens = torch.randint(1, cfg.max_valid + 1, (cfg.batch,), generator=g)
    ids = torch.randint(-2 ** 63, 2 ** 63 - 1, (cfg.batch, cfg.hist), generator=g, dtype=torch.int64)
    ids[ids == 0] = 1
    ids = torch.where(torch.arange(cfg.hist).unsqueeze(0) < lens.unsqueeze(1), ids, torch.zeros_like(ids))
    labels = torch.randint(0, 4, (cfg.batch, cfg.hist), generator=g, dtype=torch.int64)
    ts = torch.randint(1_600_000_000, 1_760_000_000, (cfg.batch, cfg.hist), generator=g, dtype=torch.int64)
    return {"product_ids": ids, "labels": labels, "timestamp": ts}


def kshift_table(cfg: HarnessConfig) -> torch.Tensor:
    """The one big tensor of the model is regenerated, not stored: numpy's PCG64 stream is
    platform-independent; the fixture keeps its float64 checksum."""
    rng = np.random.default_rng(1234 + cfg.seed)
    return torch.from_numpy(rng.standard_normal((cfg.vocab, cfg.inp_emb_dim)).astype(np.float32))


# ------------------------------------------------------------------ layer flavours ----
class _OracleFlat(nn.Module):
    """commons/layers.py:44-61 restated."""

    def __init__(self, num_embeddings, emb_dim, **_):
        super().__init__()
        self._num_embeddings = num_embeddings
        self._emb_table = nn.Embedding(num_embeddings, emb_dim)

    def forward(self, x):
        return self._emb_table(torch.remainder(x, self._num_embeddings).long())


class _OraclePattern(nn.Module):
    """commons/layers.py:13-41 with R5."""

    def __init__(self, div, mod, emb_dim, **_):
        super().__init__()
        self.div, self.mod = div, mod
        self.emb = nn.Embedding(mod, emb_dim)

    def forward(self, x):
        return self.emb(torch.remainder(torch.floor_divide(x.long(), self.div), self.mod))


class _OracleKShift(nn.Module):
    """commons/layers.py:125-185 restated (via the oracle's function)."""

    def __init__(self, num_embeddings, emb_dim, num_shifts=8, normalize_output=False, **_):
        super().__init__()
        self.emb = nn.Embedding(num_embeddings, emb_dim)
        self._num_shifts, self._normalize_output = num_shifts, normalize_output

    def forward(self, ids):
        from oracle import embedding_oracle as O
        return O.kshift_embedding(self.emb.weight, ids, self._num_shifts, self._normalize_output)


class _OracleCosine(nn.Module):
    """commons/transformers/layers.py:443-471 restated."""

    def __init__(self, inp_dim, emb_dim, n_proj=16, num_bins=20, **_):
        super().__init__()
        self.register_buffer("projection_mat", F.normalize(torch.randn((inp_dim, n_proj)), p=2.0, dim=0))
        self.register_buffer("grid", torch.linspace(-1.0, 1.0, steps=num_bins + 1)[:-1] + 1.0 / float(num_bins))
        self.register_buffer("pos_offset", (num_bins + 1) * torch.arange(0, n_proj, dtype=torch.long))
        self.emb = nn.EmbeddingBag((num_bins + 1) * n_proj, emb_dim, mode="sum")
        self.emb_dim, self.n_proj = emb_dim, n_proj

    def forward(self, x):
        from oracle import embedding_oracle as O
        bs, seq_len, _ = x.size()
        idxs = O.cosine_bucket_indices(x, self.projection_mat, self.grid, self.pos_offset)
        return self.emb(idxs).view(bs, seq_len, self.emb_dim)


class _OracleLogQ(nn.Module):
    """commons/layers.py:189-237 restated with R9 (state_dict keys models.<i>.b / .a)."""

    def __init__(self, num_buckets, hash_offsets, alpha=0.05, p_init=0.01, **_):
        super().__init__()
        self.offsets, self.alpha = list(hash_offsets), alpha
        self.models = nn.ModuleList()
        for _o in hash_offsets:
            m = nn.Module()
            m.register_buffer("b", (1.0 / p_init) * torch.ones((num_buckets,), dtype=torch.float32))
            m.register_buffer("a", torch.zeros((num_buckets,), dtype=torch.float))
            self.models.append(m)

    def forward(self, products):
        from oracle import embedding_oracle as O
        return O.logq_forward([m.b for m in self.models], self.offsets, products)

    @torch.no_grad()
    def train_step(self, products, batch_idx, skip_mask=None):
        from oracle import embedding_oracle as O
        if skip_mask is not None:
            products = products.reshape(-1)[~skip_mask.reshape(-1)]
        O.logq_train_step([m.b for m in self.models], [m.a for m in self.models], self.offsets, products,
                          self.alpha, batch_idx)


class _Block(nn.Module):
    """TransformerBlock restated for the dense, non-sparse, MLP-rotator case
    (commons/transformers/layers.py:236-258, :300-420, :40-63): pre-LN attention + MLP, both residual.
    Dense torch math outside the embedding path; needed only so that the step can run without
    /root/reference.  Parameter names follow the reference (ln_1, attn.c_attn, attn.c_proj, ln_2,
    mlp.c_fc, mlp.c_proj, and the two empty index buffers)."""

    def __init__(self, tcfg):
        super().__init__()
        a = tcfg.attn_config
        self.is_causal, self.n_head = tcfg.is_causal, a.n_head
        self.ln_1 = nn.LayerNorm(a.n_embd, eps=1e-5)
        self.attn = nn.Module()
        self.attn.c_attn = nn.Linear(a.n_embd, 3 * a.n_embd, bias=a.bias)
        self.attn.c_proj = nn.Linear(a.n_embd, a.n_embd, bias=a.bias)
        self.ln_2 = nn.LayerNorm(a.n_embd, eps=1e-5)
        self.mlp = nn.Module()
        hidden = int((tcfg.rotator_config if isinstance(tcfg.rotator_config, (int, float)) else 4) * a.n_embd)
        self.mlp.c_fc = nn.Linear(a.n_embd, hidden, bias=a.bias)
        self.mlp.c_proj = nn.Linear(hidden, a.n_embd, bias=a.bias)
        self.register_buffer("input_mask_idx", torch.empty(0, dtype=torch.long))
        self.register_buffer("input_mask_not_idx", torch.empty(0, dtype=torch.long))

    def forward(self, x):
        B, T, C = x.shape
        h = self.ln_1(x)
        q, k, v = self.attn.c_attn(h).split(C, dim=2)
        q, k, v = (t.view(B, T, self.n_head, C // self.n_head).transpose(1, 2) for t in (q, k, v))
        qk = (q @ k.transpose(-2, -1)) / math.sqrt(float(C // self.n_head))
        if self.is_causal:
            keep = torch.ones((T, T), device=x.device, dtype=torch.bool).tril(diagonal=0)
            qk = qk + keep.float().masked_fill(~keep, -float("inf")).unsqueeze(0).unsqueeze(1)
        y = (F.softmax(qk, dim=-1) @ v).transpose(1, 2).contiguous().view(B, T, C)
        x = x + self.attn.c_proj(y)
        return x + self.mlp.c_proj(F.gelu(self.mlp.c_fc(self.ln_2(x)), approximate="tanh"))


class QueryTowerH(nn.Module):
    """QueryTower restated (models/lthm/sequence/query_tower.py:14-137) over a layer flavour; R5-R7."""

    def __init__(self, mc, L):
        super().__init__()
        emb_dim = self.emb_dim = mc.emb_dim
        self.inp_proj = nn.Linear(mc.product_tower.out_emb_dim, emb_dim)
        self.action_embedding = L.FlatEmbedding(4, emb_dim)
        self.time_embedding = nn.ModuleDict(dict(
            hod=L.PatternFromTimelocal(60 * 60, 24, emb_dim),
            how=L.PatternFromTimelocal(60 * 60, 24 * 7, emb_dim),
            dow=L.PatternFromTimelocal(60 * 60 * 24, 7, emb_dim)))
        self.transformer = nn.ModuleDict(dict(
            dropout=nn.Dropout(mc.transformer_config.dropout),
            residual_attn=nn.ModuleList([_Block(mc.transformer_config)
                                         for _ in range(mc.transformer_config.num_layers)])))
        self.wpe = nn.Embedding(mc.context_width + 1, emb_dim)
        self.pad = nn.Parameter(torch.randn((1, 1, emb_dim)) / math.sqrt(emb_dim))
        self.export_tokens, self.export_span = mc.export_tokens, mc.export_span
        self.outcome_conditioning = L.FlatEmbedding(4, emb_dim)
        self.emb_heads = nn.ModuleList([nn.Linear(emb_dim, mc.product_tower.product_emb_dim, bias=False)
                                        for _ in range(self.export_tokens)])
        self.trim_fn = getattr(L, "trim_fn", None)
        self.lookup_sum = getattr(L, "lookup_sum", None)

    @staticmethod
    def reference_trim(mask_inp: torch.Tensor, export_span: int) -> int:
        """query_tower.py:73-79: leading columns that are padding in EVERY row are dropped, but at
        least export_span columns stay."""
        seq_len = mask_inp.shape[1]
        all_pad = mask_inp.all(dim=0)
        if int(all_pad.sum()) > seq_len - export_span:
            return seq_len - export_span
        return int(torch.nonzero((~all_pad).cumsum(dim=0) > 0)[0, 0])

    def forward(self, input, target, mask_inp, labels, timestamp, ids, future_outcome=torch.zeros((1, 1))):
        bsz = input.size(0)
        device = input.device
        trim = self.trim_fn(mask_inp, self.export_span) if self.trim_fn else \
            self.reference_trim(mask_inp, self.export_span)
        mask = mask_inp.unsqueeze(-1)[:, trim:].contiguous()
        input = input[:, trim:].contiguous()
        labels = labels[:, trim:].contiguous().long()
        timestamp = timestamp[:, trim:].contiguous().long()
        target = target[:, trim:].contiguous()
        ids = ids[:, trim:].contiguous()
        if self.lookup_sum is not None:
            # query_tower.py:89-104 as one kernel: four tiny-table lookups + the dense term + the pad select
            x = self.lookup_sum(self.inp_proj(input),
                                [(self.action_embedding, labels), (self.time_embedding.hod, timestamp),
                                 (self.time_embedding.how, timestamp), (self.time_embedding.dow, timestamp)],
                                mask=mask.squeeze(-1), masked_row=self.pad)
            seq_len = x.size(1)
        else:
            x = self.inp_proj(input) + self.action_embedding(labels) + self.time_embedding.hod(timestamp) \
                + self.time_embedding.how(timestamp) + self.time_embedding.dow(timestamp)
            seq_len = x.size(1)
            x = torch.where(mask, self.pad.expand(bsz, seq_len, -1), x)
        pos = seq_len - torch.arange(0, seq_len + 1, device=device).unsqueeze(0)
        x = torch.cat((torch.zeros(1, 1, self.emb_dim, device=device).expand(bsz, -1, -1), x), dim=1)
        x = x + self.wpe(pos)
        x = self.transformer.dropout(x)
        for mod in self.transformer.residual_attn:
            x = x + mod(x)  # query_tower.py:134-135 adds the block's (already residual) output again
        outcomes = torch.cat((labels, future_outcome.to(device=device, dtype=torch.long).expand(bsz, -1)), dim=-1)
        x = x + self.outcome_conditioning(outcomes)
        y = torch.stack([mod(x) for mod in self.emb_heads], dim=2)
        return {"current_token_emb": target, "next_token_emb": y, "current_token_mask": mask,
                "current_token_ids": ids}


def oracle_layers():
    return SimpleNamespace(name="oracle", FlatEmbedding=_OracleFlat, PatternFromTimelocal=_OraclePattern,
                           KShiftEmbedding=_OracleKShift, CosineVectorEmbedding=_OracleCosine, LogQ=_OracleLogQ,
                           QueryTower=lambda mc: QueryTowerH(mc, oracle_layers()))


def reference_layers():
    """The reference's own classes (needs /root/reference).  QueryTower is the IMPORTED class."""
    if not REF.exists():
        raise RuntimeError("/root/reference is not mounted")
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    import commons.layers as cl
    import commons.transformers.layers as tl
    import models.lthm.sequence.query_tower as qt

    class PatternFixed(cl.PatternFromTimelocal):  # R5
        def __init__(self, div, mod, emb_dim):
            nn.Module.__init__(self)
            self.div, self.mod, self.emb_dim = div, mod, emb_dim
            self.emb = nn.Embedding(num_embeddings=mod, embedding_dim=emb_dim)

    class LogQFixed(cl.CascadedStreamingLogQCorrectionModule):  # R9
        def train_step(self, products, batch_idx, skip_mask=None):
            if skip_mask is not None:
                products = products.reshape(-1)[~skip_mask.reshape(-1)]
            with torch.no_grad():
                for mod in self.models:
                    h = mod.hash_fn(products)
                    mod.b[h] = ((1 - mod.alpha) * mod.b[h]) + (mod.alpha * (batch_idx - mod.a[h])).float()
                    mod.a[h] = batch_idx

    def query_tower(mc):
        qt.PatternFromTimelocal = PatternFixed
        m = qt.QueryTower(mc)
        m.emb_dim = m.ememb_dim  # R6
        return m

    return SimpleNamespace(name="reference", FlatEmbedding=cl.FlatEmbedding, PatternFromTimelocal=PatternFixed,
                           KShiftEmbedding=cl.KShiftEmbedding, CosineVectorEmbedding=tl.CosineVectorEmbedding,
                           LogQ=LogQFixed, QueryTower=query_tower)


def b200_layers(device="cuda:0", trim_fn=None, pretrim=False, fused_front_end=False):
    """recommendations_b200 modules swapped in (torch-compatible gradient mode: the reference's AdamW
    drives every parameter, wrapper.py:255-275).  pretrim: the product lookup is windowed BEFORE the rows
    are moved (recommendations_b200.sequence.SequenceWindow on ids == 0; QueryTower's own trim then acts on
    the window).  fused_front_end: QueryTower's input sum (query_tower.py:89-104) through
    recommendations_b200.fused_lookup_sum and its trim through sequence_trim."""
    import recommendations_b200 as R
    from functools import partial
    L = SimpleNamespace(name="b200",
                        FlatEmbedding=partial(R.FlatEmbedding, device=device),
                        PatternFromTimelocal=partial(R.PatternFromTimelocal, device=device),
                        KShiftEmbedding=partial(R.KShiftEmbedding, device=device),
                        CosineVectorEmbedding=partial(R.CosineVectorEmbedding, device=device),
                        LogQ=partial(R.CascadedStreamingLogQCorrectionModule, device=device),
                        trim_fn=trim_fn, pretrim=pretrim,
                        lookup_sum=R.fused_lookup_sum if fused_front_end else None)
    L.QueryTower = lambda mc: QueryTowerH(mc, L)
    return L


# ---------------------------------------------------------------- the model + the step ----
class LTHMStep(nn.Module):
    """LTHMModelWrapper + Encoder + ProductTower restated over a layer flavour (R1-R4, R7-R9)."""

    def __init__(self, cfg: HarnessConfig, L):
        super().__init__()
        self.cfg = cfg
        mc = cfg.model_config()
        t = mc.product_tower
        enc = self._model = nn.Module()
        enc.product_emb_module = L.KShiftEmbedding(t.latent_model_config.vocab_size_latent, t.inp_emb_dim,  # R4
                                                   num_shifts=t.latent_model_config.num_shifts_latent,
                                                   normalize_output=t.latent_model_config.normalize_embedding)
        pt = enc.product_tower = nn.Module()
        pt.emb_mapper = nn.Linear(t.inp_emb_dim, t.out_emb_dim)
        pt.direction_emb = nn.ModuleList([
            L.CosineVectorEmbedding(t.inp_emb_dim, t.out_emb_dim, n_proj=c.num_proj, num_bins=c.num_bins)  # R3
            for c in t.cosine_lsh_config])
        pt.product_mapper = nn.Linear(t.out_emb_dim, t.product_emb_dim, bias=False)
        enc.query_tower = L.QueryTower(mc)
        self._log_q_calc = L.LogQ(num_buckets=cfg.logq_buckets, hash_offsets=list(cfg.logq_offsets),
                                  alpha=cfg.logq_alpha, p_init=cfg.logq_p_init)
        self.batch_idx = 0
        self.fused_logq_mask = L.name == "b200"
        self.pretrim = bool(getattr(L, "pretrim", False))
        self.last_window_keep = None

    # product_tower.py:43-62
    def product_tower(self, ids, x):
        pt = self._model.product_tower
        x = x.detach()
        x_norm = x.norm(p=2.0, dim=-1)
        mask = torch.logical_or(x_norm < self.cfg.norm_threshold, ids == 0)
        x = F.normalize(x, p=2.0, dim=-1)
        emb = pt.emb_mapper(x)
        for mod in pt.direction_emb:
            emb = emb + mod(x)
        emb = emb.masked_fill(mask.unsqueeze(-1), 0.0)
        return emb, pt.product_mapper(emb), mask

    # encoder.py:44-61
    def forward(self, batch):
        ids = batch["product_ids"]
        assert ids.dtype == torch.int64  # wrapper.py:52
        labels, timestamp = batch["labels"], batch["timestamp"]
        if self.pretrim:
            from recommendations_b200.sequence import SequenceWindow
            w = SequenceWindow.from_ids(ids, self.cfg.model_config().export_span)
            embs = self._model.product_emb_module(ids, window=w)   # [B, keep, D]: trimmed columns never moved
            ids, labels, timestamp = w.narrow(ids), w.narrow(labels), w.narrow(timestamp)
            self.last_window_keep = w.keep
        else:
            embs = self._model.product_emb_module(ids)
        inp, target, mask = self.product_tower(ids, embs)
        inp, target, mask, labels, timestamp, ids = [torch.flip(t, dims=[1]) for t in
                                                     (inp, target, mask, labels, timestamp, ids)]
        return self._model.query_tower(inp, target, mask, labels, timestamp, ids)

    # wrapper.py:114-245 (metrics dropped; loss path line by line)
    def train_step(self, output):
        cfg = self.cfg
        output_emb = F.normalize(output["next_token_emb"], p=2.0, dim=-1)
        input_emb = F.normalize(output["current_token_emb"], p=2.0, dim=-1)
        mask = output["current_token_mask"]
        device = output_emb.device
        batch_size, emb_dim, seq_len = output_emb.size(0), output_emb.size(-1), input_emb.size(-2)
        assert input_emb.size(-1) == emb_dim and output_emb.size(1) == seq_len + 1
        assert output_emb.size(2) == cfg.export_tokens
        product_ids = output["current_token_ids"]  # R7
        if self.fused_logq_mask:  # the boolean compaction folded into the update kernel
            self._log_q_calc.train_step(product_ids, self.batch_idx, skip_mask=mask.view(product_ids.shape))
        else:
            self._log_q_calc.train_step(product_ids.view(-1)[mask.view(-1) == 0], self.batch_idx)
        log_q_correction = self._log_q_calc(product_ids)
        self.batch_idx += 1
        loss = torch.zeros((1,), device=device)
        previous_offset = 0
        for i, max_offset in enumerate(cfg.lookahead):
            offset = max_offset if i == 0 else random.randint(previous_offset + 1, max_offset)
            previous_offset = offset
            mask_ = mask[:, offset:].contiguous()
            log_q_ = log_q_correction[:, offset:].contiguous()
            this_seq_len = seq_len - offset
            if this_seq_len <= 0:
                continue
            input_emb_ = input_emb[:, offset:].reshape(-1, emb_dim)
            output_emb_ = output_emb[:, :this_seq_len, i].reshape(-1, emb_dim)
            bs_ = output_emb_.size(0)
            labels = torch.arange(0, bs_, device=device)
            log_q_ = log_q_.reshape(1, -1).repeat(bs_, 1)
            log_q_.index_put_((labels, labels), torch.zeros_like(labels, dtype=log_q_.dtype), accumulate=False)
            pos = torch.arange(0, batch_size, dtype=torch.long, device=device).unsqueeze(1) \
                .repeat(1, this_seq_len).view(-1, 1)
            pos_matrix = torch.eq(pos, pos.T)
            eye = torch.eye(bs_, dtype=torch.bool, device=device)
            mask_ = mask_.view(-1)
            logits = (output_emb_ @ input_emb_.T) / cfg.softmax_temperature
            logits = torch.where(pos_matrix & ~eye, -float("inf"), logits)
            logits = torch.where(mask_.unsqueeze(0), -float("inf"), logits)
            logits = torch.where(mask_.unsqueeze(1), -float("inf"), logits)
            num_negatives = (~torch.isinf(logits)).sum(dim=-1) - 1
            not_use = torch.logical_or(mask_, num_negatives <= 0)
            if bool(not_use.all()):
                continue
            logits, labels, log_q_ = logits[~not_use], labels[~not_use], log_q_[~not_use]
            unreduced = F.cross_entropy(logits - cfg.logq_beta * log_q_, labels, reduction="none")
            unreduced = unreduced[~unreduced.isnan()]
            if unreduced.size(0) == 0:
                continue
            loss = loss + unreduced.mean()
        return loss


def run_step(model: LTHMStep, batch: Dict[str, torch.Tensor], steps: int = 1):
    """accelerate_training_strategy.py:353-368 + wrapper.py:255-275: forward, loss, the `0.0 * sum |p|`
    term that touches every parameter, backward, AdamW."""
    cfg = model.cfg
    opt = torch.optim.AdamW(model.parameters(), lr=cfg.lr, weight_decay=cfg.weight_decay, betas=cfg.betas)
    random.seed(cfg.seed)  # R10
    res, losses = None, []
    for _ in range(steps):
        opt.zero_grad()
        output = model(batch)
        loss = model.train_step(output)
        loss = loss + 0.0 * sum(p.abs().sum() for p in model.parameters())
        loss.backward()
        grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
        opt.step()
        losses.append(float(loss.detach().cpu()))
        res = {"loss": loss.detach(), "losses": losses, "output": {k: v.detach() for k, v in output.items()},
               "grads": grads}
    return res


EMBEDDING_PARAMS = ("_model.product_emb_module.emb.weight",
                    "_model.query_tower.action_embedding._emb_table.weight",
                    "_model.query_tower.outcome_conditioning._emb_table.weight",
                    "_model.query_tower.time_embedding.hod.emb.weight",
                    "_model.query_tower.time_embedding.how.emb.weight",
                    "_model.query_tower.time_embedding.dow.emb.weight")


def embedding_param_names(model: LTHMStep) -> List[str]:
    names = list(EMBEDDING_PARAMS)
    names += [f"_model.product_tower.direction_emb.{i}.emb.weight" for i in range(len(model.cfg.cosine))]
    return names


def fix_margins(model: LTHMStep, batch, min_margin: float = 2e-5, seed: int = 99) -> int:
    """Re-draws (in place) the valid ids whose k-shift embedding projects closer than `min_margin` to a
    bucket boundary of any CosineVectorEmbedding: fp32 matmul noise (~3e-7 here) then cannot move a
    bucket index between CPU and GPU, so the comparison below is about the kernels, not about
    torch.bucketize tie-breaking.  Returns the number of ids replaced."""
    g = torch.Generator().manual_seed(seed)
    ids = batch["product_ids"]
    replaced = 0
    for _ in range(50):
        with torch.no_grad():
            x = model._model.product_emb_module(ids).double()
            nrm = x.norm(p=2.0, dim=-1)
            bad = (nrm - model.cfg.norm_threshold).abs() < min_margin
            xn = x / nrm.clamp_min(1e-12).unsqueeze(-1)
            for mod in model._model.product_tower.direction_emb:
                z = xn @ mod.projection_mat.double()
                d = (z.unsqueeze(-1) - mod.grid.double()).abs().min(dim=-1).values
                bad |= (d < min_margin).any(dim=-1)
            bad &= ids != 0
        n_bad = int(bad.sum())
        if n_bad == 0:
            return replaced
        new = torch.randint(-2 ** 63, 2 ** 63 - 1, (n_bad,), generator=g, dtype=torch.int64)
        new[new == 0] = 1
        ids[bad] = new
        replaced += n_bad
    raise RuntimeError("could not find ids with the requested bucket margin")


def bucket_margin(model: LTHMStep, batch) -> float:
    """Smallest distance of any cosine projection to a bucket boundary and of any norm to the mask
    threshold (float64): inputs whose margin is above fp32 matmul noise make the integer part of the
    step (bucket indices, masks) identical on every device."""
    with torch.no_grad():
        ids = batch["product_ids"]
        x = model._model.product_emb_module(ids).double()
        nrm = x.norm(p=2.0, dim=-1)
        margin = float((nrm - model.cfg.norm_threshold).abs().min())
        xn = x / nrm.clamp_min(1e-12).unsqueeze(-1)
        for mod in model._model.product_tower.direction_emb:
            z = xn @ mod.projection_mat.double()
            d = (z.unsqueeze(-1) - mod.grid.double()).abs().min(dim=-1).values  # [B, L, n_proj]
            d = d[ids != 0]
            margin = min(margin, float(d.min()))
    return margin


# ------------------------------------------------------------------ fixture helpers ----
def model_from_golden(g, L, device="cpu") -> LTHMStep:
    """A model of flavour L carrying the fixture's initial state (the k-shift table is regenerated
    and checked against the stored checksum)."""
    cfg = HarnessConfig()
    model = LTHMStep(cfg, L)
    w0 = kshift_table(cfg)
    chk = np.array([w0.double().sum().item(), (w0.double() ** 2).sum().item()])
    if not np.array_equal(chk, np.asarray(g["kshift_checksum"])):
        raise RuntimeError("regenerated k-shift table differs from the one the fixture was made with "
                           "(numpy Generator stream changed?): regenerate tests/golden/lthm_step.npz")
    sd = {k[4:]: torch.from_numpy(np.asarray(g[k])) for k in g.files if k.startswith("sd0/")}
    sd["_model.product_emb_module.emb.weight"] = w0
    model = model.to(device)
    model.load_state_dict(sd, strict=True)
    return model


def batch_from_golden(g, device="cpu") -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(np.asarray(g[k])).to(device) for k in ("product_ids", "labels", "timestamp")}


def compare_with_golden(g, model: LTHMStep, res, *, loss_rtol, out_atol, out_rtol, grad_tol, w_tol,
                        report=None) -> None:
    """Loss, outputs, embedding-table gradients and the tables after the 2nd step vs the fixture.
    grad_tol / w_tol are relative to the largest magnitude of the tensor compared."""
    def note(name, val):
        if report is not None:
            report[name] = val

    losses = np.asarray(g["losses"])
    note("loss_rel", float(np.max(np.abs(np.array(res["losses"]) - losses) / np.abs(losses))))
    np.testing.assert_allclose(res["losses"], losses, rtol=loss_rtol)
    out = {k: v.cpu() for k, v in res["output"].items()}
    assert torch.equal(out["current_token_ids"], torch.from_numpy(g["current_token_ids"]))  # trim + flip
    assert torch.equal(out["current_token_mask"], torch.from_numpy(g["current_token_mask"]))
    for key, sub, rs in (("next_token_emb", "next_token_emb", "next_token_rowsum"),
                         ("current_token_emb", "current_token_emb", "current_token_rowsum")):
        want = torch.from_numpy(g[sub])
        note(key, float((out[key][::8] - want).abs().max()))
        torch.testing.assert_close(out[key][::8], want, rtol=out_rtol, atol=out_atol)
        torch.testing.assert_close(out[key].sum(-1), torch.from_numpy(g[rs]), rtol=out_rtol,
                                   atol=out_atol * out[key].shape[-1])
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    big = "_model.product_emb_module.emb.weight"
    assert res["grads"][big].abs().max().item() == 0.0 == float(g["kshift_grad_absmax"])  # detached
    torch.testing.assert_close(sd[big][::97], torch.from_numpy(g["kshift_after"]), rtol=1e-6, atol=1e-7)
    for n in embedding_param_names(model):
        if n == big:
            continue
        want_g = torch.from_numpy(g[f"grad/{n}"])
        got_g = res["grads"][n].cpu()
        scale = want_g.abs().max().item()
        note(f"grad:{n}", float((got_g - want_g).abs().max()) / max(scale, 1e-30))
        assert (got_g - want_g).abs().max().item() <= grad_tol * scale, n
        # AdamW from a zero state moves every element by ~lr * sign(g): only elements whose gradient is
        # far above the summation noise are comparable at 1e-5; the others are counted and bounded
        want_w, got_w = torch.from_numpy(g[f"sd2/{n}"]), sd[n]
        solid = want_g.abs() > 1e-4 * scale
        err = (got_w - want_w).abs()
        note(f"w:{n}", float(err[solid].max()) if solid.any() else 0.0)
        assert err[solid].max().item() <= w_tol * max(1.0, want_w.abs().max().item()), n
        assert err.max().item() <= 2.5 * model.cfg.lr * 2, n  # a sign flip of a noise-level gradient at most
        assert solid.float().mean().item() > 0.5 or scale == 0.0, n
    for k in g.files:
        if k.startswith("sd2/_log_q_calc"):
            torch.testing.assert_close(sd[k[4:]], torch.from_numpy(g[k]), rtol=1e-6, atol=0.0)
