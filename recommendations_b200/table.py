"""EmbeddingTable: the weight holder behind every embedding module of this package.

It keeps the reference's state_dict key (`<name>.weight`, as nn.Embedding /
nn.EmbeddingBag would) and offers the two gradient modes of SURVEY.md section 8(b):

  torch-compatible   weight is an nn.Parameter; backward produces the dense
                     (or, sparse=True, uncoalesced COO) gradient nn.Embedding
                     would, so the reference's optim_group /
                     optimizers_for_param_groups flow
                     (commons/base_model_wrapper.py:51-72) keeps working.
  fused              weight is a buffer (hidden from parameters(), so the loops at
                     accelerate_training_strategy.py:355, :360-362, :378 do not
                     scan it); backward applies the optimizer to the touched rows
                     inside the segmented-reduction kernel.
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.nn as nn

from . import _native as N
from . import ops


@dataclass
class FusedOptimizerConfig:
    """Hyper-parameters of the in-kernel update.  Defaults follow torch.optim."""
    kind: str = "adagrad"  # sgd | adagrad | rowwise_adagrad | adam | adamw
    lr: float = 0.5        # embedding_module_gen.py:97, :137 use Adagrad(lr=5e-1)
    eps: float = 1e-10
    weight_decay: float = 0.0
    betas: Tuple[float, float] = (0.9, 0.999)
    lr_decay: float = 0.0
    initial_accumulator_value: float = 0.0
    # A table looked up more than once per optimizer step (a shared table serving history and
    # target lookups, gradient accumulation) must see ONE update with the SUMMED gradient, as
    # torch.optim does on .grad.  accumulate=False (default) applies the update inside backward
    # and refuses a second backward of the same table before FusedEmbeddingOptimizer.step();
    # accumulate=True keeps every lookup's (ids, gradient rows) until step(), which merges them
    # into one plan and applies one update.
    accumulate: bool = False

    def __post_init__(self):
        if self.kind not in ("sgd", "adagrad", "rowwise_adagrad", "adam", "adamw"):
            raise ValueError(f"unknown fused optimizer kind {self.kind!r}")


class EmbeddingTable(nn.Module):
    def __init__(self, num_embeddings: int, embedding_dim: int, padding_idx: Optional[int] = None,
                 sparse: bool = False, dtype: torch.dtype = torch.float32, device=None,
                 _weight: Optional[torch.Tensor] = None):
        super().__init__()
        self.num_embeddings = int(num_embeddings)
        self.embedding_dim = int(embedding_dim)
        if padding_idx is not None and padding_idx < 0:
            padding_idx = self.num_embeddings + padding_idx
        self.padding_idx = padding_idx
        self.sparse = bool(sparse)
        if _weight is None:
            # nn.Embedding.reset_parameters: N(0, 1), padding row zeroed.  On the CPU generator
            # (device None / cpu) the values equal nn.Embedding's under the same seed; a table
            # created straight on a CUDA device is filled there, in its own dtype, so that
            # tens-of-GB shards never exist on the host or in fp32.
            if device is not None and torch.device(device).type == "cuda":
                w = torch.empty((self.num_embeddings, self.embedding_dim), dtype=dtype, device=device)
                w.normal_()
            else:
                w = torch.empty((self.num_embeddings, self.embedding_dim), dtype=torch.float32)
                nn.init.normal_(w)
                w = w.to(dtype=dtype, device=device)
            if padding_idx is not None:
                with torch.no_grad():
                    w[padding_idx].fill_(0)
        else:
            w = _weight
        self.weight = nn.Parameter(w)
        self.fused: Optional[FusedOptimizerConfig] = None
        self.fused_step = 0          # completed optimizer steps
        self._anchor: Optional[torch.Tensor] = None
        self._has_facade = False     # a FusedEmbeddingOptimizer drives the step boundaries
        self._applied_in_step = 0    # updates applied since the last step boundary
        self._pending = []           # accumulate mode: lookups waiting for step()
        self._plan_buf: Optional[torch.Tensor] = None   # plan memory reused from step to step
        self._plan_owner = None      # weakref to the plan that currently lives in _plan_buf

    # ---------------------------------------------------------------- modes ----
    def enable_fused_optimizer(self, config: Optional[FusedOptimizerConfig] = None, **kw) -> "EmbeddingTable":
        """Turn the Parameter into a buffer and apply `config` inside backward."""
        cfg = config or FusedOptimizerConfig(**kw)
        w = self.weight.detach()
        if "weight" in self._parameters:
            del self._parameters["weight"]
            self.register_buffer("weight", w, persistent=True)
        self.fused = cfg
        self.fused_step = 0
        for name in ("opt_state1", "opt_state2"):
            if name in self._buffers:
                del self._buffers[name]
        if w.is_cuda:
            # create the optimizer state now, on the stream the module is built on, not lazily inside the first
            # backward (autograd thread): no fill kernel right in front of the first update, fixed addresses
            # before any CUDA graph is captured
            self._ensure_state()
        return self

    def _ensure_state(self) -> None:
        cfg = self.fused
        w = self.weight
        if cfg.kind == "adagrad" and "opt_state1" not in self._buffers:
            self.register_buffer("opt_state1", torch.full(w.shape, cfg.initial_accumulator_value,
                                                          dtype=torch.float32, device=w.device),
                                 persistent=False)
        elif cfg.kind == "rowwise_adagrad" and "opt_state1" not in self._buffers:
            self.register_buffer("opt_state1", torch.full((w.shape[0],), cfg.initial_accumulator_value,
                                                          dtype=torch.float32, device=w.device),
                                 persistent=False)
        elif cfg.kind in ("adam", "adamw") and "opt_state1" not in self._buffers:
            self.register_buffer("opt_state1", torch.zeros(w.shape, dtype=torch.float32, device=w.device),
                                 persistent=False)
            self.register_buffer("opt_state2", torch.zeros(w.shape, dtype=torch.float32, device=w.device),
                                 persistent=False)

    def grad_anchor(self) -> torch.Tensor:
        """The tensor a lookup Function differentiates against."""
        if self.fused is None:
            return self.weight
        dev = self.weight.device
        if self._anchor is None or self._anchor.device != dev:
            self._anchor = torch.zeros((), device=dev, requires_grad=True)
        return self._anchor

    # ----------------------------------------------------------------- plan ----
    def build_plan(self, ids: torch.Tensor, **kw) -> ops.BackwardPlan:
        """ops.BackwardPlan.build into a buffer this table keeps from step to step.  A fresh torch
        allocation per step would be made on the side stream the plan is built on and handed to the main
        stream with record_stream(): the caching allocator can then only recycle it after a stream event,
        and an un-synchronised training loop falls back to cudaMalloc every step (measured: cfg 2 through
        the modules 8.2 instead of 3.9 ms).  The cached buffer is handed out to ONE live plan at a time
        (a second lookup of the table before the first one's backward gets a fresh allocation)."""
        busy = self._plan_owner is not None and self._plan_owner() is not None
        buf = None
        if not busy and self._plan_buf is not None and self._plan_buf.device == ids.device:
            buf = self._plan_buf
        plan = ops.BackwardPlan.build(ids, buf=buf, **kw)
        if not busy:
            self._plan_buf = plan.buf
            self._plan_owner = weakref.ref(plan)
        return plan

    def _release_plan(self, plan) -> None:
        if self._plan_owner is not None and self._plan_owner() is plan:
            self._plan_owner = None

    # ------------------------------------------------------------- backward ----
    def consume(self, plan: ops.BackwardPlan, grad2d: torch.Tensor, slots_per_grad_row: int = 1,
                slot_weight: Optional[torch.Tensor] = None,
                grad_row_scale: Optional[torch.Tensor] = None, grad_div: float = 0.0,
                guard: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """Reduce grad rows over `plan`; fused: update in place and return None,
        torch-compatible: return the dense gradient of `weight`.  `guard`: see ops.bwd_apply."""
        w = self.weight
        if w.dtype == torch.float32 and grad2d.dtype != torch.float32:
            grad2d = grad2d.float()
        if self.fused is None:
            gw = torch.zeros_like(w)
            ops.bwd_apply(plan, grad2d, table=gw, update=N.UPD_DENSE_GRAD,
                          slots_per_grad_row=slots_per_grad_row, slot_weight=slot_weight,
                          grad_row_scale=grad_row_scale, grad_div=grad_div)
            self._release_plan(plan)
            return gw
        if self.fused.accumulate:
            if not self._has_facade:
                raise N.NativeError("FusedOptimizerConfig(accumulate=True) needs a FusedEmbeddingOptimizer: its "
                                    "step() applies the accumulated lookups")
            self._pending.append((plan, grad2d, slots_per_grad_row, slot_weight, grad_row_scale, grad_div))
            return None
        self.begin_update()
        self._apply_fused(plan, grad2d, slots_per_grad_row, slot_weight, grad_row_scale, grad_div, guard=guard)
        self.end_update()
        self._release_plan(plan)
        return None

    def begin_update(self) -> None:
        """Refuses a second in-backward update of this table inside one optimizer step."""
        if self._has_facade and self._applied_in_step >= 1:
            raise N.NativeError(
                "this fused table already applied an update in the current optimizer step: a second backward "
                "through it before FusedEmbeddingOptimizer.step() would apply two separate updates instead of "
                "one with the summed gradient (Adagrad would accumulate g1^2 + g2^2, not (g1 + g2)^2).  Use "
                "FusedOptimizerConfig(accumulate=True) for tables looked up more than once per step.")

    def end_update(self) -> None:
        if self._has_facade:
            self._applied_in_step += 1
        else:
            self.fused_step += 1  # no facade: every backward is one optimizer step

    def _apply_fused(self, plan, grad2d, slots_per_grad_row, slot_weight, grad_row_scale, grad_div,
                     rows: Optional[Tuple[int, int]] = None, guard: Optional[torch.Tensor] = None,
                     peer_push: Optional[dict] = None) -> None:
        """`rows` = (r0, r1): the plan's keys are relative to that row range of the table (one table group
        of a stacked shard); the update touches nothing outside it.
        `peer_push` (dict(group, tables, bags_per_table, rows_per_table, push_ctas)): grad2d are THIS rank's
        pooled gradients and the launch pushes them to every rank itself (ops.peer_bwd_apply_fused)."""
        cfg = self.fused
        self._ensure_state()
        step = self.fused_step + 1
        lr = cfg.lr
        if cfg.kind in ("adagrad", "rowwise_adagrad"):
            lr = cfg.lr / (1.0 + (step - 1) * cfg.lr_decay)
        hp = ops.make_optim_params(lr=lr, eps=cfg.eps, weight_decay=cfg.weight_decay,
                                   beta1=cfg.betas[0], beta2=cfg.betas[1], step=step)
        w, s1, s2 = self.weight.data, self._buffers.get("opt_state1"), self._buffers.get("opt_state2")
        if rows is not None:
            r0, r1 = rows
            w = w[r0:r1]
            s1 = None if s1 is None else s1[r0:r1]
            s2 = None if s2 is None else s2[r0:r1]
        with torch.no_grad():
            if peer_push is not None:
                ops.peer_bwd_apply_fused(plan, grad2d, table=w, update=N.UPDATE_BY_NAME[cfg.kind], state1=s1, hp=hp,
                                         **peer_push)
                return
            ops.bwd_apply(plan, grad2d, table=w, update=N.UPDATE_BY_NAME[cfg.kind],
                          slots_per_grad_row=slots_per_grad_row, state1=s1, state2=s2, hp=hp,
                          slot_weight=slot_weight, grad_row_scale=grad_row_scale, grad_div=grad_div, guard=guard)

    def commit_step(self) -> None:
        """Step boundary (FusedEmbeddingOptimizer.step): applies the lookups accumulated since the
        last boundary as ONE update with the summed gradient, then advances the step count."""
        pending, self._pending = self._pending, []
        if pending:
            if len(pending) == 1:
                self._apply_fused(*pending[0])
            else:
                self._apply_fused(*_merge_lookups(pending))
            self._applied_in_step += 1
        if self._applied_in_step:
            self.fused_step += 1
        self._applied_in_step = 0

    def extra_repr(self) -> str:
        mode = "torch-grad" if self.fused is None else f"fused-{self.fused.kind}"
        return (f"{self.num_embeddings}, {self.embedding_dim}, padding_idx={self.padding_idx}, "
                f"dtype={self.weight.dtype}, mode={mode}")


def _merge_lookups(pending):
    """Several lookups of one table inside one optimizer step -> one plan over the concatenated ids
    and one concatenated gradient: the segmented reduction then sums every row's gradient over all
    lookups before the single update (what autograd's accumulation into .grad + one optim.step()
    does in the reference)."""
    recipes = [p[0].recipe for p in pending]
    if any(r is None for r in recipes):
        raise N.NativeError("accumulate=True: a lookup's plan was not built from ids and cannot be merged")
    kw0 = recipes[0][1]
    spg0, gdiv0 = pending[0][2], pending[0][5]
    for (_, kw), p in zip(recipes, pending):
        if kw.get("lengths") is not None or kw.get("ids_per_table") or kw.get("shard_world", 1) > 1 \
                or kw != kw0 or p[2] != spg0 or p[5] != gdiv0:
            raise N.NativeError("accumulate=True merges lookups of one kind only (same hashing, pad mask, flip and "
                                "bag shape; no per-bag lengths, table batching or sharding)")
    if any((p[3] is None) != (pending[0][3] is None) or (p[4] is None) != (pending[0][4] is None) for p in pending):
        raise N.NativeError("accumulate=True: lookups with and without per-slot weights cannot be merged")
    ids = torch.cat([r[0].reshape(-1) for r in recipes])
    grads = [p[1] for p in pending]
    dt = torch.float32 if any(g.dtype == torch.float32 for g in grads) else grads[0].dtype
    grad = torch.cat([g.to(dt) for g in grads])
    sw = None if pending[0][3] is None else torch.cat([p[3].reshape(-1) for p in pending])
    gs = None if pending[0][4] is None else torch.cat([p[4].reshape(-1) for p in pending])
    plan = ops.BackwardPlan.build(ids, **kw0)
    return plan, grad, spg0, sw, gs, gdiv0


class FusedEmbeddingOptimizer(torch.optim.Optimizer):
    """Facade returned from optimizers_for_param_groups
    (commons/base_model_wrapper.py:64-72) for tables in fused mode: the update happens
    inside backward (or, accumulate=True, inside step() with the gradient summed over
    every lookup of the step); step() marks the step boundary, the object carries the
    hyper-parameters and exposes the optimizer state.  Without a facade every backward
    through a fused table counts as one optimizer step."""

    def __init__(self, tables, **overrides):
        self.tables = [t for t in tables if isinstance(t, EmbeddingTable)]
        if not self.tables:
            raise ValueError("FusedEmbeddingOptimizer needs at least one EmbeddingTable")
        for t in self.tables:
            if t.fused is None:
                t.enable_fused_optimizer(FusedOptimizerConfig(**overrides))
            else:
                for k, v in overrides.items():
                    setattr(t.fused, k, v)
        for t in self.tables:
            t._has_facade = True
        params = [t.grad_anchor() for t in self.tables]
        super().__init__(params, dict(lr=self.tables[0].fused.lr))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        # a scheduler may have rewritten param_groups[*]['lr']: push it to the tables
        for group in self.param_groups:
            for t in self.tables:
                t.fused.lr = float(group["lr"])
        # step boundary: accumulate-mode tables apply their summed update now, every table that
        # was updated advances its step count (bias correction / lr_decay count optimizer steps,
        # not backward calls)
        for t in self.tables:
            t.commit_step()
        return loss

    def zero_grad(self, set_to_none: bool = True):
        for t in self.tables:
            if t._anchor is not None:
                t._anchor.grad = None

    def fused_state_dict(self):
        return {i: {"step": t.fused_step,
                    "state1": t._buffers.get("opt_state1"),
                    "state2": t._buffers.get("opt_state2")} for i, t in enumerate(self.tables)}

    # Checkpoint / resume through the standard optimizer API (the reference saves optimizers with
    # accelerator.save_state, accelerate_training_strategy.py:260-266; resume was absent there):
    # the in-kernel optimizer state of every table travels inside state_dict()["fused"].
    def state_dict(self):
        sd = super().state_dict()
        sd["fused"] = {i: {"kind": t.fused.kind, "step": t.fused_step,
                           "state1": None if st["state1"] is None else st["state1"].detach().clone(),
                           "state2": None if st["state2"] is None else st["state2"].detach().clone()}
                       for (i, t), st in zip(enumerate(self.tables), self.fused_state_dict().values())}
        return sd

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        state_dict = dict(state_dict)
        fused = state_dict.pop("fused", None)
        super().load_state_dict(state_dict)
        for group in self.param_groups:
            for t in self.tables:
                t.fused.lr = float(group["lr"])
        if fused is None:
            return
        if len(fused) != len(self.tables):
            raise ValueError(f"checkpoint holds {len(fused)} fused tables, optimizer has {len(self.tables)}")
        for i, t in enumerate(self.tables):
            rec = fused[i] if i in fused else fused[str(i)]
            if rec["kind"] != t.fused.kind:
                raise ValueError(f"table {i}: checkpoint optimizer {rec['kind']!r} != {t.fused.kind!r}")
            t.fused_step = int(rec["step"])
            t._ensure_state()
            for name, key in (("opt_state1", "state1"), ("opt_state2", "state2")):
                buf = t._buffers.get(name)
                if buf is None:
                    continue
                if rec[key] is None or tuple(rec[key].shape) != tuple(buf.shape):
                    raise ValueError(f"table {i}: optimizer state {key} missing or of the wrong shape")
                buf.copy_(rec[key].to(buf.device, buf.dtype))
