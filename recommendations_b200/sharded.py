"""Row-wise sharded pooled embedding bag over the GPUs of one box (north_star item 4, cfg 5).

Global row r of the table lives on rank `r % W` at local row `r // W` (hashed ids make this
uniformly balanced and it also spreads the k-shift collapse rows round-robin).  One process per
GPU; torch.distributed (NCCL over NVLink / NVSwitch) is the plumbing.  The reference itself has no
model-parallel embeddings (tables replicated): sharded == unsharded is the parity
statement, checked on one GPU (W emulated), on CPU/gloo (host logic) and on N GPUs.

Two exchange strategies (`exchange=`):

"route" (default for W > 1): every rank buckets its slots by owner (stable counting sort), one
  counts all-to-all + ONE host read of the bucket sizes, then a variable-split all-to-all moves
  each lookup only to the rank that owns its row.  An owner pools / sorts / updates just the
  ~n lookups it serves, independent of W.
      forward   ids --bucket--> entries(owner-major) --all_to_all_v--> owner: run-pool by (sender, bag)
                partial [W, B, D] --all_to_all--> requester: sum over owners in order 0..W-1
      backward  grad_out --all_gather--> owner: plan from its received entries + segmented
                reduction + fused update of ITS rows (no gradient all-reduce)

"gather" (fixed-size collectives only, no host sync; owner-side work grows with W):

  forward   ids [b, P]  --all_gather-->  [W, b, P]
            owner: partial pool of EVERY rank's bags over the rows it owns
                   (recemb_pool_fwd with shard_world / shard_rank: non-owned slots are skipped)
            partial [W, b, D]  --all_to_all-->  recv [W, b, D] (owner s's partial of MY bags)
            out = sum over owners in order 0..W-1 (recemb_sum_partials, fp32 accumulate)
  backward  grad_out [b, D]  --all_gather-->  [W, b, D]
            owner: sort-based plan over the gathered ids (non-owned slots dropped, keys = local
                   rows) + segmented reduction + fused update of ITS rows (no gradient all-reduce)

"peer" (CUDA only; NVLink / NVSwitch peer memory, no collective-library call and no host
  synchronisation on the data path; fixed shapes, CUDA-graph capturable):

  forward   ONE kernel: each lookup loads its row straight from the owner's shard over NVLink and
            is pooled here in slot order -> bit-identical to the unsharded pooled bag
  backward  bucket-by-owner whose scatter stores the entries into the owners' inboxes, push
            all-gather of the pooled gradients, device-side barrier, owner: sort of its inbox +
            segmented reduction + fused update of ITS rows; one more barrier before the next
            forward reads the updated rows

NVLink bytes per GPU per step (W = 8, b = 8192, P = 20, T tables, R = row bytes): ids in
(W-1) b T P 8, partials in/out (W-1) b T R each way, grads in (W-1) b T R.
"""
from __future__ import annotations

import os

from typing import Callable, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from . import _native as N
from . import ops
from .layers import pooled_counts
from .peer import PeerGroup, arena_layout
from .table import EmbeddingTable, FusedOptimizerConfig


def local_rows_of(num_embeddings: int, world: int, rank: int) -> int:
    return (num_embeddings - rank + world - 1) // world


# ------------------------------------------------------------ collectives ----
class Collectives:
    """all_gather / all_to_all over a process group.  NCCL uses the native collectives; other
    backends (gloo, for the CPU tests of the host logic) emulate all_to_all with all_gather."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.native_a2a = dist.get_backend(group) == "nccl"

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        flat = t.contiguous().view(-1)
        out = torch.empty((self.world * flat.numel(),), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, flat, group=self.group)
        return out.view((self.world,) + tuple(t.shape))

    def all_gather_start(self, t: torch.Tensor):
        """Asynchronous all_gather: returns (output view, wait()).  The collective runs on the
        backend's own stream; kernels enqueued before wait() overlap with it."""
        flat = t.contiguous().view(-1)
        out = torch.empty((self.world * flat.numel(),), dtype=t.dtype, device=t.device)
        work = dist.all_gather_into_tensor(out, flat, group=self.group, async_op=True)
        return out.view((self.world,) + tuple(t.shape)), work.wait

    def exchange_counts(self, counts: torch.Tensor):
        """counts [W] (device): how many entries I send to each rank -> (send, recv) host lists.
        The one host synchronisation of the routed exchange."""
        if self.native_a2a:
            recv = torch.empty_like(counts)
            dist.all_to_all_single(recv, counts, group=self.group)
            both = torch.stack([counts, recv]).cpu()
            return both[0].tolist(), both[1].tolist()
        matrix = self.all_gather(counts).cpu()            # [src, dst]
        return matrix[self.rank].tolist(), matrix[:, self.rank].tolist()

    def all_to_all_v(self, send: torch.Tensor, send_counts, recv_counts) -> torch.Tensor:
        """send = buckets for rank 0, 1, ... back to back; returns the buckets received from rank
        0, 1, ... back to back."""
        n_send, n_recv = int(sum(send_counts)), int(sum(recv_counts))
        out = torch.empty((n_recv,) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        if self.native_a2a:
            dist.all_to_all_single(out, send[:n_send].contiguous(), output_split_sizes=list(recv_counts),
                                   input_split_sizes=list(send_counts), group=self.group)
            return out
        # emulation for backends without all_to_all (gloo): all_gather padded sends, slice mine
        matrix = self.all_gather(torch.tensor(send_counts, dtype=torch.int64, device=send.device)).cpu()
        cap = int(matrix.sum(dim=1).max())
        padded = torch.zeros((cap,) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        padded[:n_send] = send[:n_send]
        everyone = self.all_gather(padded)
        pieces = []
        for src in range(self.world):
            off = int(matrix[src, :self.rank].sum())
            pieces.append(everyone[src, off:off + int(matrix[src, self.rank])])
        return torch.cat(pieces) if pieces else out

    def all_to_all(self, t: torch.Tensor) -> torch.Tensor:
        """t [W, ...]: slice s goes to rank s; returns [W, ...] with slice s received from rank s."""
        t = t.contiguous()
        if self.native_a2a:
            out = torch.empty_like(t)
            dist.all_to_all_single(out, t, group=self.group)
            return out
        everyone = self.all_gather(t)          # [W(src), W(dst), ...]
        return everyone[:, self.rank].contiguous()


class SingleProcess(Collectives):
    """W == 1: identity collectives (no process group needed)."""

    def __init__(self):
        self.group, self.world, self.rank, self.native_a2a = None, 1, 0, False

    def all_gather(self, t):
        return t.unsqueeze(0)

    def all_gather_start(self, t):
        return t.unsqueeze(0), (lambda: None)

    def all_to_all(self, t):
        return t

    def exchange_counts(self, counts):
        c = counts.cpu().tolist()
        return c, c

    def all_to_all_v(self, send, send_counts, recv_counts):
        return send[:int(sum(send_counts))]


# ---------------------------------------------------------------- autograd ----
class _ShardedPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, ids, lengths, module):
        # ids [B, P] with B = T * b bags of this rank (table-major when T > 1)
        comm: Collectives = module.comm
        w, b, p = comm.world, ids.shape[0], ids.shape[1]
        ids_all = comm.all_gather(ids).view(w * b, p)
        len_all = None if lengths is None else comm.all_gather(lengths.to(torch.int32)).view(w * b)
        partial = module.local_pool(module.emb.weight.detach(), ids_all, len_all)      # [W*B, D]
        recv = comm.all_to_all(partial.view(w, b, -1))                                 # [W, B, D]
        scale = None
        if module.mode == "mean":
            scale = 1.0 / pooled_counts(ids, lengths, module.last_n, module.skip_pad,
                                        module.pad_id).clamp_(min=1).float()
        ctx.module = module
        ctx.save_for_backward(ids_all, len_all, scale)
        return module.reduce_partials(recv, scale)

    @staticmethod
    def backward(ctx, grad_out):
        ids_all, len_all, scale = ctx.saved_tensors
        module = ctx.module
        g = grad_out.contiguous()
        if scale is not None:
            g = g * scale.unsqueeze(1).to(g.dtype)
        g_all = module.comm.all_gather(g).view(ids_all.shape[0], -1)                   # [W*b, D]
        return module.local_backward(ids_all, len_all, g_all), None, None, None


def _mark(module, name):
    """Optional phase timing (module.phase_events = []): CUDA events at phase boundaries."""
    ev = getattr(module, "phase_events", None)
    if ev is not None and torch.cuda.is_available():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        ev.append((name, e))


class _RoutedPoolFn(torch.autograd.Function):
    """exchange="route": each lookup travels only to the rank that owns its row."""

    @staticmethod
    def forward(ctx, anchor, ids, lengths, module):
        comm: Collectives = module.comm
        w, b = comm.world, ids.shape[0]
        _mark(module, "start")
        entries, counts = module.bucket(ids, lengths)                    # owner-major, device
        _mark(module, "bucket")
        send_c, recv_c = comm.exchange_counts(counts)                    # host sync (bucket sizes)
        _mark(module, "counts+sync")
        recv = comm.all_to_all_v(entries, send_c, recv_c)                # int64 [n_recv]
        _mark(module, "a2a_entries")
        partial = module.pool_entries(module.emb.weight.detach(), recv, w * b)   # [W*B, D]
        _mark(module, "pool_entries")
        back = comm.all_to_all(partial.view(w, b, -1))                   # owner s's partial of MY bags
        _mark(module, "a2a_partials")
        scale = None
        if module.mode == "mean":
            scale = 1.0 / pooled_counts(ids, lengths, module.last_n, module.skip_pad,
                                        module.pad_id).clamp_(min=1).float()
        ctx.module = module
        ctx.save_for_backward(recv, scale)
        out = module.reduce_partials(back, scale)
        _mark(module, "reduce")
        return out

    @staticmethod
    def backward(ctx, grad_out):
        recv, scale = ctx.saved_tensors
        module = ctx.module
        g = grad_out.contiguous()
        if scale is not None:
            g = g * scale.unsqueeze(1).to(g.dtype)
        _mark(module, "bwd_start")
        # the sort of the received entries does not need the gradients: it runs while the
        # all-gather of the pooled gradients is in flight on the NCCL stream
        g_all, wait = module.comm.all_gather_start(g)                    # [W, B, D]
        g_all = g_all.view(-1, g_all.shape[-1])
        res = module.entries_backward(recv, g_all, wait)
        _mark(module, "ag_grads||plan+apply")
        return res, None, None, None


class _PeerPoolFn(torch.autograd.Function):
    """exchange="peer": rows pulled / entries and gradients pushed through peer-mapped memory.

    Stream plan of one training step (main = the caller's stream, side = the module's own):
      forward   main: [barrier ch0 if rows changed] -> pull-pool kernel (NVLink-bound)
                side: bucket + push entries -> barrier ch1 -> sort of my inbox -> barrier ch1
                      (the backward plan only depends on the ids: it is built in the shadow of
                      the forward pull and of whatever the model computes before backward)
      backward  main: push-all-gather of the pooled gradients -> barrier ch0 -> wait(side) ->
                      segmented reduction + fused update of MY rows
    Hazards: a peer may overwrite my inbox only after I sorted it (2nd ch1 barrier), my gradient
    buffer only after my update read it and my rows may be read only after the update (ch0
    barrier at the start of the next forward)."""

    @staticmethod
    def forward(ctx, anchor, ids, lengths, module):
        pg = module.peer_group()
        pg.raise_on_status()
        main = torch.cuda.current_stream(ids.device)
        _mark(module, "start")
        ctx.plan, ctx.plan_ready, ctx.pg = None, None, pg
        if module.peer_forward == "push":
            return _PeerPoolFn._forward_push(ctx, ids, lengths, module, pg, main)
        if module._peer_dirty:
            # rows updated since the last barrier (backward / optimizer.step / a weight load) must
            # be complete on every rank before anybody reads them
            ops.peer_barrier(pg, channel=0)
            module._peer_dirty = False
        if ctx.needs_input_grad[0]:
            side = module._side_stream(ids.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ops.peer_bucket_push(pg, ids, num_rows=module.num_embeddings, lengths=lengths,
                                     last_n=module.last_n, zero_pad=module.skip_pad, pad_id=module.pad_id,
                                     **module._own_batching(ids))
                ops.peer_barrier(pg, channel=1)
                ctx.plan = ops.peer_plan(pg, module.emb.weight.shape[0])
                ops.peer_barrier(pg, channel=1)
                ctx.plan_ready = torch.cuda.Event()
                ctx.plan_ready.record(side)
            ids.record_stream(side)
            if lengths is not None:
                lengths.record_stream(side)
            ctx.plan.buf.record_stream(main)
        out = ops.peer_pool_fwd(
            pg, ids, num_rows=module.num_embeddings, dim=module.emb_dim, dtype=module.emb.weight.dtype,
            lengths=lengths, last_n=module.last_n,
            pool_mode=N.POOL_MEAN if module.mode == "mean" else N.POOL_SUM, zero_pad=module.skip_pad,
            pad_id=module.pad_id, **module._own_batching(ids))
        _mark(module, "peer_pool")
        scale = None
        if module.mode == "mean":
            scale = 1.0 / pooled_counts(ids, lengths, module.last_n, module.skip_pad,
                                        module.pad_id).clamp_(min=1).float()
        ctx.module = module
        ctx.save_for_backward(scale)
        return out

    @staticmethod
    def _forward_push(ctx, ids, lengths, module, pg, main):
        """Owner-side partial pooling: entries to the owners, one partial row per (bag, owner)
        pair back -- (W-1) b T R bytes over NVLink at most instead of (W-1)/W n R.
          main: bucket + push entries -> barrier -> pool my inbox, STORE the partial rows into
                the requesters' parts (zero rows for pairs without an entry) -> [join side] ->
                barrier -> sum parts
          side: (after the first barrier) unpack + sort of my inbox = the backward plan
        The owners only read their own shard, so no barrier guards the table rows; inbox, parts
        and gradient buffers are protected by the three barriers of the step."""
        dt = module.emb.weight.dtype
        parts = pg.parts_view(module.emb_dim, dt)
        ops.peer_bucket_push(pg, ids, num_rows=module.num_embeddings, lengths=lengths, last_n=module.last_n,
                             zero_pad=module.skip_pad, pad_id=module.pad_id, **module._own_batching(ids))
        ops.peer_barrier(pg, channel=0)
        _mark(module, "bucket_push+barrier")
        if ctx.needs_input_grad[0]:
            side = module._side_stream(ids.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ctx.plan = ops.peer_plan(pg, module.emb.weight.shape[0])
                ctx.plan_ready = torch.cuda.Event()
                ctx.plan_ready.record(side)
            ctx.plan.buf.record_stream(main)
        ops.peer_pool_push(pg, module.emb_dim, dt)
        _mark(module, "pool_push")
        if ctx.plan_ready is not None:
            main.wait_event(ctx.plan_ready)   # peers may refill my inbox after the next barrier
        ops.peer_barrier(pg, channel=0)
        scale = None
        if module.mode == "mean":
            scale = 1.0 / pooled_counts(ids, lengths, module.last_n, module.skip_pad,
                                        module.pad_id).clamp_(min=1).float()
        out = ops.sum_partials(parts, scale)
        _mark(module, "barrier+sum")
        module._peer_dirty = False
        ctx.module = module
        ctx.save_for_backward(scale)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (scale,) = ctx.saved_tensors
        module = ctx.module
        pg = ctx.pg                      # the group (arena) of the batch shape this forward used
        g = grad_out.contiguous()
        if scale is not None:
            g = g * scale.unsqueeze(1).to(g.dtype)
        if g.dtype != module.emb.weight.dtype:
            g = g.to(module.emb.weight.dtype)
        main = torch.cuda.current_stream(g.device)
        _mark(module, "bwd_start")
        if module._peer_dirty:
            # two backward passes without a forward in between (several outstanding forwards): the
            # peers' previous update may still be reading the gradient buffer this pass overwrites
            ops.peer_barrier(pg, channel=0)
        if module._fused_push_ok(g):
            # ONE launch: its first CTAs push my gradients to every rank table by table, the others reduce and
            # update my rows as soon as the tables they touch have arrived from everybody (no barrier, the
            # NVLink transfer of table t + 1 runs under the HBM-bound update of table t)
            main.wait_event(ctx.plan_ready)
            t = module.num_tables
            module.emb.begin_update()
            module.emb._apply_fused(ctx.plan, g, 1, None, None, 0.0, peer_push=dict(
                group=pg, tables=t, bags_per_table=int(pg.layout.bags_total) // t,
                rows_per_table=module.local_rows, push_ctas=module.push_ctas))
            module.emb.end_update()
            module.emb._release_plan(ctx.plan)
            res = None
            _mark(module, "push+apply (fused)")
        else:
            ops.peer_allgather_push(pg, g, int(pg.layout.off_grads))
            _mark(module, "grads_push")
            ops.peer_barrier(pg, channel=0)
            main.wait_event(ctx.plan_ready)
            _mark(module, "barrier+plan")
            # guard = my arena's status word: an overflowed inbox (flagged by the sender) or a timed-out
            # barrier leaves the shard untouched in this step; the host raises at its next status check
            res = module.emb.consume(ctx.plan, pg.grads_view(module.emb_dim, module.emb.weight.dtype),
                                     slots_per_grad_row=1, guard=pg.status_word())
            _mark(module, "apply")
        ctx.plan = None
        pg.snapshot_status()
        module._peer_dirty = True
        return res, None, None, None


class _PeerSeqFn(torch.autograd.Function):
    """exchange="peer", sequence mode (RowWiseShardedEmbedding): every lookup is its own output row.

      forward   main: [barrier ch0 if rows changed] -> pool_kernel<PEER> with one slot per bag: each row LOADED
                      from its owner's shard over NVLink -- (W-1)/W n R bytes, bit-identical to the unsharded gather
                side: bucket by owner, entries STORED into the owners' inboxes (entry = local row, slot of the
                      owner's gradient buffer) -> barrier ch1 -> unpack + sort of my inbox -> barrier ch1
      backward  main: rows_scatter_push: gradient row i STORED into its owner's buffer, once (an all-gather would
                      move it W - 1 times) -> barrier ch0 -> wait(side) -> segmented reduction + update of MY rows
    Hazards as in _PeerPoolFn."""

    @staticmethod
    def forward(ctx, anchor, ids, module):
        pg = module.peer_group()
        pg.raise_on_status()
        main = torch.cuda.current_stream(ids.device)
        ctx.plan, ctx.plan_ready, ctx.pg, ctx.dest = None, None, pg, None
        batching = module._own_batching(ids)
        if module._peer_dirty:
            ops.peer_barrier(pg, channel=0)
            module._peer_dirty = False
        if ctx.needs_input_grad[0]:
            side = module._side_stream(ids.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ctx.dest = ops.peer_bucket_push_rows(
                    pg, ids, num_rows=module.num_embeddings, zero_pad=module.skip_pad, pad_id=module.pad_id,
                    ids_per_table=batching["bags_per_table"], num_tables=batching["num_tables"])
                ops.peer_barrier(pg, channel=1)
                ctx.plan = ops.peer_plan(pg, module.emb.weight.shape[0], buf=module.__dict__.get("_seq_plan_buf"))
                module._seq_plan_buf = ctx.plan.buf
                ops.peer_barrier(pg, channel=1)
                ctx.plan_ready = torch.cuda.Event()
                ctx.plan_ready.record(side)
            ids.record_stream(side)
            ctx.dest.record_stream(main)
        out = ops.peer_pool_fwd(pg, ids, num_rows=module.num_embeddings, dim=module.emb_dim,
                                dtype=module.emb.weight.dtype, pool_mode=N.POOL_SUM, zero_pad=module.skip_pad,
                                pad_id=module.pad_id, **batching)
        ctx.module = module
        return out

    @staticmethod
    def backward(ctx, grad_out):
        module, pg = ctx.module, ctx.pg
        g = grad_out.contiguous()
        if g.dtype != module.emb.weight.dtype:
            g = g.to(module.emb.weight.dtype)
        main = torch.cuda.current_stream(g.device)
        if module._peer_dirty:
            ops.peer_barrier(pg, channel=0)     # two backward passes without a forward in between
        main.wait_event(ctx.plan_ready)         # dest is written on the side stream
        ops.peer_rows_scatter_push(pg, g.view(-1, g.shape[-1]), ctx.dest)
        ops.peer_barrier(pg, channel=0)
        res = module.emb.consume(ctx.plan, pg.grads_view(module.emb_dim, module.emb.weight.dtype),
                                 slots_per_grad_row=1, guard=pg.status_word())
        ctx.plan, ctx.dest = None, None
        pg.snapshot_status()
        module._peer_dirty = True
        return res, None, None


CH_ENTRIES, CH_PARTS, CH_GRADS, CH_REPEAT = 0, 1, 2, 3   # barrier channels of a pipelined group's arena


class _PeerPipelinedFn(torch.autograd.Function):
    """exchange="peer", push forward, T >= 2 tables: the step is cut into table groups, each with its own
    exchange arena, and the kernels are issued by ROLE, not by group: everything bound by the local HBM
    stays on the caller's stream, everything bound by NVLink goes to a second stream, the sort of the
    inboxes to a third.  The groups are staggered by construction (the NVLink stream works on group g while
    the HBM stream buckets group g + 1 / sums group g - 1 / updates group g - 1), and no stream ever sits in
    a full barrier: a producer SIGNALS its peers right after its kernel, a consumer WAITS right before its
    own (recemb_peer_signal / recemb_peer_wait, one flag set per group and phase).

      forward   main: bucket + push entries_g, signal E_g        (all g)  ...  wait P_g, sum parts_g  (all g)
                nv:   wait E_g, pool my inbox_g + STORE partial rows to the requesters, signal P_g     (all g)
                side: (after wait E_g) unpack + sort inbox_g = the backward plan of group g            (all g)
      backward  nv:   push gradients_g to every owner, [join sort_g], signal G_g                      (all g)
                main: wait G_g, segmented reduction + update of group g's rows of my shard            (all g)

    Hazards: my inbox_g is refilled by a peer's bucket of the NEXT step, which that peer enqueues after its
    update_g, i.e. after it waited for my G_g -- which I only signal once my pooling and my sort of inbox_g
    are done.  parts_g / gradient buffer_g are refilled after the peer waited for my next E_g, which I signal
    after sum_g / update_g (stream order on main).  Owners read only their own shard; the nv stream's pooling
    of the next step starts after main's bucket of that step, i.e. after the update.
    Same kernels and per-group arithmetic as _PeerPoolFn's push path: the forward is bit-identical."""

    @staticmethod
    def forward(ctx, anchor, ids, lengths, module):
        groups = module._pipe
        dev = ids.device
        main = torch.cuda.current_stream(dev)
        nv, side = module._pipe_streams(dev)
        t, dim, dt = module.num_tables, module.emb_dim, module.emb.weight.dtype
        b = ids.shape[0] // t
        for grp in groups:
            grp.pg.raise_on_status()
        need = bool(ctx.needs_input_grad[0])
        out = torch.empty((ids.shape[0], dim), dtype=dt, device=dev)
        scale = None
        if module.mode == "mean":
            scale = 1.0 / pooled_counts(ids, lengths, module.last_n, module.skip_pad, module.pad_id).clamp_(min=1).float()
        pooled = []
        for grp in groups:
            pg, rows = grp.pg, slice(grp.t0 * b, grp.t1 * b)
            tg = grp.t1 - grp.t0
            batching = dict(bags_per_table=b, num_tables=tg) if tg > 1 else dict(bags_per_table=0, num_tables=0)
            ops.peer_bucket_push(pg, ids[rows], num_rows=module.num_embeddings,
                                 lengths=None if lengths is None else lengths[rows], last_n=module.last_n,
                                 zero_pad=module.skip_pad, pad_id=module.pad_id, **batching)
            ops.peer_signal(pg, CH_ENTRIES)
            bucketed = torch.cuda.Event()
            bucketed.record(main)
            with torch.cuda.stream(nv):
                nv.wait_event(bucketed)
                ops.peer_wait(pg, CH_ENTRIES)
                arrived = torch.cuda.Event()
                arrived.record(nv)
                ops.peer_pool_push(pg, dim, dt)
                ops.peer_signal(pg, CH_PARTS)
                ev = torch.cuda.Event()
                ev.record(nv)
                pooled.append(ev)
            grp.plan, grp.plan_ready = None, None
            if need:
                with torch.cuda.stream(side):
                    side.wait_event(arrived)
                    grp.plan = ops.peer_plan(pg, tg * module.local_rows, buf=grp.plan_buf)
                    grp.plan_buf = grp.plan.buf
                    grp.plan_ready = torch.cuda.Event()
                    grp.plan_ready.record(side)
        for grp, ev in zip(groups, pooled):
            pg, rows = grp.pg, slice(grp.t0 * b, grp.t1 * b)
            main.wait_event(ev)
            ops.peer_wait(pg, CH_PARTS)
            ops.sum_partials(pg.parts_view(dim, dt), None if scale is None else scale[rows], out=out[rows])
        module._peer_dirty = False
        ctx.module, ctx.b = module, b
        ctx.save_for_backward(scale)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (scale,) = ctx.saved_tensors
        module, b = ctx.module, ctx.b
        groups = module._pipe
        dim, dt = module.emb_dim, module.emb.weight.dtype
        g = grad_out.contiguous()
        if scale is not None:
            g = g * scale.unsqueeze(1).to(g.dtype)
        if g.dtype != dt:
            g = g.to(dt)
        dev = g.device
        main = torch.cuda.current_stream(dev)
        nv, _ = module._pipe_streams(dev)
        nv.wait_stream(main)
        g.record_stream(nv)
        pushed = []
        with torch.cuda.stream(nv):
            for grp in groups:
                pg, rows = grp.pg, slice(grp.t0 * b, grp.t1 * b)
                if grp.plan is None:
                    raise N.NativeError("pipelined peer step: backward without a recorded forward")
                if module._peer_dirty:
                    # two backward passes without a forward in between: the peers' previous update may still be
                    # reading the gradient buffer this pass overwrites
                    ops.peer_barrier(pg, channel=CH_REPEAT)
                ops.peer_allgather_push(pg, g[rows], int(pg.layout.off_grads))
                nv.wait_event(grp.plan_ready)   # my G_g also tells the peers that inbox_g may be refilled
                ops.peer_signal(pg, CH_GRADS)
                ev = torch.cuda.Event()
                ev.record(nv)
                pushed.append(ev)
        module.emb.begin_update()
        for grp, ev in zip(groups, pushed):
            pg = grp.pg
            main.wait_event(ev)
            ops.peer_wait(pg, CH_GRADS)
            module.emb._apply_fused(grp.plan, pg.grads_view(dim, dt), 1, None, None, 0.0,
                                    rows=(grp.t0 * module.local_rows, grp.t1 * module.local_rows),
                                    guard=pg.status_word())
            pg.snapshot_status()
            grp.plan = None
        module.emb.end_update()
        module._peer_dirty = True
        return None, None, None, None


class _PipeGroup:
    """Tables [t0, t1) of the stacked shard with their own exchange arena."""

    def __init__(self, t0: int, t1: int, pg: PeerGroup):
        self.t0, self.t1, self.pg = t0, t1, pg
        self.plan, self.plan_ready, self.plan_buf = None, None, None


class RowWiseShardedEmbeddingBag(nn.Module):
    """Pooled multi-hot lookup into a row-wise sharded table.

    forward(ids [b, P] int64 of THIS rank's batch, lengths [b] optional) -> [b, D]; with
    num_tables = T > 1 (T equal-shaped tables, one launch / one collective for all of them):
    ids [T, b, P], lengths [T, b] -> [T, b, D].
    `emb.weight` holds this rank's shard [T * ceil((N - rank) / W), D] (state_dict key
    `emb.weight`; `gather_full_weight()` reassembles the global table(s) for a reference-shaped
    checkpoint)."""

    def __init__(self, num_embeddings: int, emb_dim: int, mode: str = "sum", *, last_n: int = 0,
                 num_tables: int = 1, skip_pad: bool = False, pad_id: int = 0, group=None,
                 comm: Optional[Collectives] = None,
                 dtype: torch.dtype = torch.float32, device=None,
                 fused_optimizer: Optional[FusedOptimizerConfig] = None,
                 local_pool: Optional[Callable] = None, local_backward: Optional[Callable] = None,
                 reduce_partials: Optional[Callable] = None, exchange: Optional[str] = None,
                 bucket: Optional[Callable] = None, pool_entries: Optional[Callable] = None,
                 entries_backward: Optional[Callable] = None, capacity_factor: Optional[float] = None,
                 peer_forward: Optional[str] = None, pipeline_groups: Optional[int] = None):
        super().__init__()
        if mode not in ("sum", "mean"):
            raise ValueError("mode must be 'sum' or 'mean'")
        if comm is None:
            comm = Collectives(group) if dist.is_available() and dist.is_initialized() else SingleProcess()
        self.comm = comm
        self.num_embeddings, self.emb_dim = int(num_embeddings), int(emb_dim)
        self.mode, self.last_n, self.skip_pad, self.pad_id = mode, int(last_n), skip_pad, pad_id
        self.num_tables = int(num_tables)
        self.local_rows = local_rows_of(num_embeddings, comm.world, comm.rank)
        self.emb = EmbeddingTable(self.num_tables * self.local_rows, emb_dim, dtype=dtype, device=device)
        if fused_optimizer is not None:
            self.emb.enable_fused_optimizer(fused_optimizer)
        # the three compute hooks default to the CUDA kernels; tests of the host logic on
        # CPU/gloo inject oracle implementations (never done in product code)
        self.local_pool = local_pool or self._cuda_local_pool
        self.local_backward = local_backward or self._cuda_local_backward
        self.reduce_partials = reduce_partials or ops.sum_partials
        self.exchange = exchange or ("route" if comm.world > 1 else "gather")
        if self.exchange not in ("route", "gather", "peer"):
            raise ValueError("exchange must be 'route', 'gather' or 'peer'")
        # peer exchange: inbox capacity per sender = capacity_factor * (my slots / W); hashed ids
        # spread evenly (binomial), skewed in-range ids (identity hashing) need more head-room
        self.capacity_factor = capacity_factor
        # "pull": requester loads rows from the owners (one kernel, bit-identical to unsharded);
        # "push": owners pool and store partial rows (fewer NVLink bytes, bf16 partial rounding)
        # default: push as soon as rows would cross NVLink (measured 0.59 vs 0.73 ms at W = 2,
        # 1.09 vs 1.15 ms at W = 8 on cfg 5), pull on a single rank
        if peer_forward is None:
            peer_forward = "pull" if comm.world == 1 else "push"
        if peer_forward not in ("pull", "push"):
            raise ValueError("peer_forward must be 'pull' or 'push'")
        self.peer_forward = peer_forward
        # peer + push + several tables: the step is pipelined over table groups (_PeerPipelinedFn); needs the
        # fused optimizer (each group's rows are updated in place).  RECEMB_PEER_GROUPS overrides the default
        # (1 = the unpipelined step: measured on cfg 5 at W = 2 / 4 the extra launches of G = 2 / 4 groups cost as
        # much as the overlap wins, 0.54 / 0.56 / 0.67 ms and 0.633 / 0.631 / 0.717 ms for G = 1 / 2 / 4).
        if pipeline_groups is None:
            pipeline_groups = int(os.environ.get("RECEMB_PEER_GROUPS", "1"))
        self.pipeline_groups = max(1, min(int(pipeline_groups), self.num_tables))
        # the backward's gradient push fused into the update launch (_fused_push_ok); CTAs that push
        # (True: whenever rows cross NVLink, i.e. world > 1; "force": also on one rank, for tests; False: never)
        env = os.environ.get("RECEMB_PEER_FUSED_PUSH", "1")
        self.fused_push = False if env == "0" else ("force" if env == "force" else True)
        # (cfg 5 step at W = 8 with 16 / 32 / 64 pusher CTAs: 0.749 / 0.770 / 0.786 ms; W = 4: 0.586 / 0.595 / 0.598)
        self.push_ctas = int(os.environ.get("RECEMB_PEER_PUSH_CTAS", "16"))
        self._pipe = None
        self._pipe_key = None
        self._peer: Optional[PeerGroup] = None
        self._peer_key = None
        self._peer_dirty = True
        self.bucket = bucket or self._cuda_bucket
        self.pool_entries = pool_entries or ops.pool_entries
        self.entries_backward = entries_backward or self._cuda_entries_backward

    # ----------------------------------------------------------- CUDA hooks ----
    def _cuda_local_pool(self, shard, ids_all, len_all):
        return ops.pool_fwd(shard, ids_all, lengths=len_all, last_n=self.last_n, pool_mode=N.POOL_SUM,
                            zero_pad=self.skip_pad, pad_id=self.pad_id, num_rows=self.num_embeddings,
                            shard_world=self.comm.world, shard_rank=self.comm.rank, **self._batching(ids_all))

    def _cuda_local_backward(self, ids_all, len_all, g_all):
        plan = ops.BackwardPlan.build(
            ids_all, num_rows=self.num_embeddings, zero_pad=self.skip_pad, pad_id=self.pad_id,
            bag_size=ids_all.shape[1], lengths=len_all, last_n=self.last_n,
            shard_world=self.comm.world, shard_rank=self.comm.rank,
            ids_per_table=self._batching(ids_all)["bags_per_table"] * ids_all.shape[1],
            num_tables=self._batching(ids_all)["num_tables"])
        return self.emb.consume(plan, g_all, slots_per_grad_row=ids_all.shape[1])

    def _cuda_bucket(self, ids, lengths):
        b_tot = ids.shape[0]
        return ops.shard_bucket(ids, num_rows=self.num_embeddings, world=self.comm.world, rank=self.comm.rank,
                                bags_total=b_tot, lengths=lengths, last_n=self.last_n, zero_pad=self.skip_pad,
                                pad_id=self.pad_id,
                                bags_per_table=b_tot // self.num_tables if self.num_tables > 1 else 0,
                                num_tables=self.num_tables if self.num_tables > 1 else 0)

    def _cuda_entries_backward(self, recv, g_all, wait=lambda: None):
        plan = ops.plan_from_entries(recv, self.emb.weight.shape[0])   # overlaps the gradient all-gather
        wait()
        return self.emb.consume(plan, g_all, slots_per_grad_row=1)

    def _own_batching(self, ids):
        """this rank's bags are [T, b]: bag g belongs to table g // b."""
        if self.num_tables == 1:
            return dict(bags_per_table=0, num_tables=0)
        return dict(bags_per_table=ids.shape[0] // self.num_tables, num_tables=self.num_tables)

    # ----------------------------------------------------------- peer group ----
    def peer_capacity(self, n_slots: int) -> int:
        w = self.comm.world
        f = self.capacity_factor
        if f is None:   # RECEMB_PEER_CAPACITY: process-wide default (tuning runs); hashed ids spread binomially
            f = 1.0 if w == 1 else float(os.environ.get("RECEMB_PEER_CAPACITY", "1.5"))
        cap = min(n_slots, int(n_slots / w * f) + 64)
        return max(2, cap + (cap & 1))

    def peer_group(self) -> PeerGroup:
        if self._peer is None and self._pipe:
            return self._pipe[0].pg
        if self._peer is None:
            raise N.NativeError("peer exchange: call forward first (the group is built for its batch shape)")
        return self._peer

    def _ensure_peer_group(self, ids: torch.Tensor) -> None:
        """(Re)builds the peer group when the batch shape or the weight storage changed.
        Collective over the module's process group: every rank must reach it together."""
        w = self.emb.weight
        key = (w.data_ptr(), tuple(ids.shape))
        if self._peer is not None and self._peer_key == key:
            return
        cache = self.__dict__.setdefault("_peer_cache", {})
        if cache and (next(iter(cache))[0] != w.data_ptr() or len(cache) >= 8):
            self.close_peer()                      # the weight moved (or too many shapes): start over
            cache = self.__dict__.setdefault("_peer_cache", {})
        if key not in cache:
            # a new batch shape (e.g. the last, smaller batch of an epoch) needs its own arena; the
            # 51 GB shard stays mapped once
            cap = self.peer_capacity(ids.numel())
            # sequence mode: the gradient buffer holds one row per inbox entry, [W][cap][D]
            bags_total = cap if getattr(self, "sequence_mode", False) else ids.shape[0]
            if self.comm.world == 1:
                layout = arena_layout(1, cap, bags_total, self.emb_dim, w.dtype)
                arena = PeerGroup.new_arena(layout, w.device)
                cache[key] = PeerGroup.local(1, 0, [arena], [w.detach()], layout)
            else:
                first = next(iter(cache.values())) if cache else None
                cache[key] = PeerGroup.connect(w.detach(), cap=cap, bags_total=bags_total, group=self.comm.group,
                                               table_ptrs=None if first is None else first.table_ptrs())
        self._peer, self._peer_key = cache[key], key
        self._peer_dirty = True

    def _fused_push_ok(self, g: torch.Tensor) -> bool:
        """The pooled backward as one launch (gradient push + gated segmented reduction): fused optimizer with a
        scalar per-row state, rows of 256 / 512 bytes.  RECEMB_PEER_FUSED_PUSH=0 keeps push -> barrier -> update."""
        f = self.emb.fused
        if f is None or f.accumulate or f.kind not in ("rowwise_adagrad", "sgd") or self.fused_push is False:
            return False
        row_bytes = self.emb_dim * self.emb.weight.element_size()
        # one rank: nothing crosses NVLink, the plain copy + update is faster (0.456 vs 0.59 ms on cfg 5)
        return ((self.comm.world > 1 or self.fused_push == "force") and row_bytes in (256, 512)
                and g.dtype == self.emb.weight.dtype
                and self.num_tables <= 64 and not getattr(self, "sequence_mode", False))

    # ------------------------------------------------------ pipelined groups ----
    def _pipelined(self) -> bool:
        return (self.exchange == "peer" and self.peer_forward == "push" and self.pipeline_groups > 1
                and self.emb.fused is not None and not self.emb.fused.accumulate)

    def _pipe_streams(self, device):
        if getattr(self, "_pstreams", None) is None:
            # [NVLink-bound kernels, inbox sorts]; the NVLink stream gets priority so that its (short) kernels
            # are not queued behind a full-grid HBM-bound kernel of the caller's stream
            self._pstreams = [torch.cuda.Stream(device=device, priority=-1), torch.cuda.Stream(device=device)]
        return self._pstreams

    def _ensure_pipeline(self, ids: torch.Tensor) -> None:
        """(Re)builds one peer group (arena + mapped shard pointers) per table group for this batch shape.
        Collective over the module's process group."""
        w = self.emb.weight
        key = (w.data_ptr(), tuple(ids.shape))
        if self._pipe is not None and self._pipe_key == key:
            return
        self.close_peer()
        t, g = self.num_tables, self.pipeline_groups
        b, p = ids.shape[0] // t, ids.shape[1]
        bounds = [round(i * t / g) for i in range(g + 1)]
        groups = []
        for t0, t1 in zip(bounds[:-1], bounds[1:]):
            view = w.detach()[t0 * self.local_rows:t1 * self.local_rows]
            bags = (t1 - t0) * b
            cap = self.peer_capacity(bags * p)
            if self.comm.world == 1:
                layout = arena_layout(1, cap, bags, self.emb_dim, w.dtype)
                pg = PeerGroup.local(1, 0, [PeerGroup.new_arena(layout, w.device)], [view], layout)
            else:
                pg = PeerGroup.connect(view, cap=cap, bags_total=bags, group=self.comm.group)
            groups.append(_PipeGroup(t0, t1, pg))
        self._pipe, self._pipe_key = groups, key
        self._peer_dirty = True

    def _side_stream(self, device) -> "torch.cuda.Stream":
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=device)
        return self._side

    def close_peer(self) -> None:
        """Collective: unmap the peers' memory (before this rank's shard / arena may be freed)."""
        cache = self.__dict__.get("_peer_cache") or {}
        pipe = self.__dict__.get("_pipe") or []
        if self._peer is None and not cache and not pipe:
            return
        torch.cuda.synchronize(self.emb.weight.device)
        if self.comm.world > 1:
            dist.barrier(group=self.comm.group)
        for pg in cache.values():
            pg.close()
        for grp in pipe:
            grp.pg.close()
        if self.comm.world > 1:
            dist.barrier(group=self.comm.group)
        self._peer, self._peer_key, self._peer_cache = None, None, {}
        self._pipe, self._pipe_key = None, None

    def _batching(self, ids_all):
        """gathered bags are [W, T, b]: bag g belongs to table (g // b) % T."""
        if self.num_tables == 1:
            return dict(bags_per_table=0, num_tables=0)
        b = ids_all.shape[0] // (self.comm.world * self.num_tables)
        return dict(bags_per_table=b, num_tables=self.num_tables)

    # -------------------------------------------------------------- forward ----
    def forward(self, ids: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        if ids.dtype != torch.int64 or ids.dim() != (2 if self.num_tables == 1 else 3):
            raise N.NativeError("ids must be int64 [batch, bag_size] ([num_tables, batch, bag_size] when batched)")
        fn = {"route": _RoutedPoolFn, "gather": _ShardedPoolFn, "peer": _PeerPoolFn}[self.exchange]
        if self.num_tables == 1:
            if self.exchange == "peer":
                self._ensure_peer_group(ids)
            return fn.apply(self.emb.grad_anchor(), ids.contiguous(), lengths, self)
        t, b, p = ids.shape
        if t != self.num_tables:
            raise N.NativeError(f"expected ids for {self.num_tables} tables, got {t}")
        flat = ids.contiguous().view(t * b, p)
        if self._pipelined():
            self._ensure_pipeline(flat)
            fn = _PeerPipelinedFn
        elif self.exchange == "peer":
            self._ensure_peer_group(flat)
        out = fn.apply(self.emb.grad_anchor(), flat,
                       None if lengths is None else lengths.contiguous().view(t * b), self)
        return out.view(t, b, -1)

    @torch.no_grad()
    def load_full_weight(self, full: torch.Tensor) -> None:
        """Take this rank's rows (r % W == rank) out of a global [N, D] (or [T, N, D]) table."""
        full = full.view(self.num_tables, self.num_embeddings, self.emb_dim)
        shard = full[:, self.comm.rank::self.comm.world].reshape(-1, self.emb_dim)
        self.emb.weight.copy_(shard.to(self.emb.weight.device, self.emb.weight.dtype))
        self._peer_dirty = True

    @torch.no_grad()
    def _gather_rows(self, local: torch.Tensor) -> torch.Tensor:
        """[T * local_rows, ...] of every rank -> global [T, N, ...] (row r of rank s = global row r * W + s)."""
        w, t = self.comm.world, self.num_tables
        rows_max = local_rows_of(self.num_embeddings, w, 0)
        tail = tuple(local.shape[1:])
        mine = local.reshape((t, self.local_rows) + tail)
        pad = torch.zeros((t, rows_max) + tail, dtype=mine.dtype, device=mine.device)
        pad[:, :self.local_rows] = mine
        shards = self.comm.all_gather(pad)                                  # [W, T, rows_max, ...]
        full = torch.empty((t, self.num_embeddings) + tail, dtype=pad.dtype, device=pad.device)
        for s in range(w):
            full[:, s::w] = shards[s, :, :local_rows_of(self.num_embeddings, w, s)]
        return full[0] if t == 1 else full

    @torch.no_grad()
    def gather_full_weight(self) -> torch.Tensor:
        """Global [N, D] ([T, N, D]) table under the unsharded module's key (checkpoint / export)."""
        return self._gather_rows(self.emb.weight)

    @torch.no_grad()
    def gather_full_optimizer_state(self) -> dict:
        """Fused-optimizer state of the GLOBAL table(s), laid out like an unsharded EmbeddingTable's
        (`state1` [N] for row-wise Adagrad, [N, D] otherwise; `state2` for Adam) + the step count:
        what a rank-0 checkpoint stores next to gather_full_weight() (collective)."""
        if self.emb.fused is None:
            raise N.NativeError("gather_full_optimizer_state needs the fused optimizer mode")
        self.emb._ensure_state()
        out = {"kind": self.emb.fused.kind, "step": self.emb.fused_step}
        for name, key in (("opt_state1", "state1"), ("opt_state2", "state2")):
            buf = self.emb._buffers.get(name)
            out[key] = None if buf is None else self._gather_rows(buf)
        return out

    @torch.no_grad()
    def load_full_optimizer_state(self, state: dict) -> None:
        """Inverse of gather_full_optimizer_state: every rank keeps the rows it owns (resume, also
        onto a different world size)."""
        if self.emb.fused is None or state["kind"] != self.emb.fused.kind:
            raise N.NativeError("optimizer kind of the checkpoint does not match this module")
        self.emb._ensure_state()
        self.emb.fused_step = int(state["step"])
        t, w, r = self.num_tables, self.comm.world, self.comm.rank
        for name, key in (("opt_state1", "state1"), ("opt_state2", "state2")):
            buf = self.emb._buffers.get(name)
            if buf is None:
                continue
            full = state[key]
            full = full.reshape((t, self.num_embeddings) + tuple(full.shape[(1 if t == 1 else 2):]))
            buf.copy_(full[:, r::w].reshape(buf.shape).to(buf.device, buf.dtype))


class RowWiseShardedEmbedding(RowWiseShardedEmbeddingBag):
    """Sequence-mode lookup (FlatEmbedding semantics: one output row per id, `out[..., :] =
    table[floor_mod(id, N)]`, commons/layers.py:56-61) into a row-wise sharded table.

    forward(ids [...] int64 of THIS rank) -> [..., D]; with num_tables = T > 1: ids [T, ...] -> [T, ..., D].
    exchange="peer": each row is pulled from its owner over NVLink in the forward and each gradient row is
    pushed to its owner once in the backward (_PeerSeqFn).  The NCCL exchanges ("route" / "gather") treat a
    lookup as a bag of one slot (functional, but the gradients are all-gathered)."""

    sequence_mode = True

    def __init__(self, num_embeddings: int, emb_dim: int, **kw):
        kw.setdefault("pipeline_groups", 1)
        kw["peer_forward"] = "pull"
        super().__init__(num_embeddings, emb_dim, mode="sum", **kw)

    def forward(self, ids: torch.Tensor) -> torch.Tensor:
        t = self.num_tables
        if t > 1 and (ids.dim() < 2 or ids.shape[0] != t):
            raise N.NativeError(f"expected ids [T = {t}, ...], got {tuple(ids.shape)}")
        shape = tuple(ids.shape)
        flat = ids.contiguous().view(-1, 1)
        if self.exchange != "peer":
            out = super().forward(flat.view(t, -1, 1) if t > 1 else flat)
            return out.reshape(shape + (self.emb_dim,))
        self._ensure_peer_group(flat)
        out = _PeerSeqFn.apply(self.emb.grad_anchor(), flat, self)
        return out.view(shape + (self.emb_dim,))


# ------------------------------------------------------------------ table-wise partitioning ----
class _TableWiseFn(torch.autograd.Function):
    """exchange over peer memory, tables partitioned whole: table t lives on rank t % W.

      forward   main: bucket: every lookup of table t goes to rank t % W's inbox -> barrier -> the owner pools its
                      (sender, bag) runs from its tables and STORES the pooled row (exactly one owner per bag: it IS
                      the result) into the requester's parts[owner][bag] -> [join side] -> barrier -> pick
                      parts[t % W][bags of table t]
                side: (after the first barrier) unpack + sort of my inbox = the backward plan
      backward  main: ONE launch: the first CTAs push the gradients of table t to rank t % W only (each gradient row
                      crosses NVLink at most once), the others reduce + update my tables, gated per local table
                      (other update kinds: all-gather push -> barrier -> update)"""

    @staticmethod
    def forward(ctx, anchor, ids, lengths, module):
        pg = module.peer_group()
        pg.raise_on_status()
        main = torch.cuda.current_stream(ids.device)
        t, dim, dt = module.num_tables, module.emb_dim, module.emb.weight.dtype
        b = ids.shape[0] // t
        ops.peer_bucket_push(pg, ids, num_rows=module.num_embeddings, lengths=lengths, last_n=module.last_n,
                             zero_pad=module.skip_pad, pad_id=module.pad_id, bags_per_table=b, num_tables=t,
                             tablewise=True)
        ops.peer_barrier(pg, channel=0)
        ctx.plan, ctx.plan_ready, ctx.pg = None, None, pg
        if ctx.needs_input_grad[0]:
            side = module._side_stream(ids.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ctx.plan = ops.peer_plan(pg, module.emb.weight.shape[0], buf=module.__dict__.get("_tw_plan_buf"))
                module._tw_plan_buf = ctx.plan.buf
                ctx.plan_ready = torch.cuda.Event()
                ctx.plan_ready.record(side)
        ops.peer_pool_push(pg, dim, dt, tablewise_bags_per_table=b)
        if ctx.plan_ready is not None:
            main.wait_event(ctx.plan_ready)       # peers may refill my inbox after the next barrier
        ops.peer_barrier(pg, channel=0)
        parts = pg.parts_view(dim, dt).view(pg.world, t, b, dim)
        out = parts[module._owner_idx, module._table_idx].reshape(t * b, dim)   # table t comes from rank t % W
        scale = None
        if module.mode == "mean":
            scale = 1.0 / pooled_counts(ids, lengths, module.last_n, module.skip_pad, module.pad_id).clamp_(min=1).float()
            out = (out.float() * scale.unsqueeze(1)).to(dt)
        ctx.module, ctx.b = module, b
        ctx.save_for_backward(scale)
        module._peer_dirty = False
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (scale,) = ctx.saved_tensors
        module, pg, b = ctx.module, ctx.pg, ctx.b
        dt = module.emb.weight.dtype
        g = grad_out.contiguous()
        if scale is not None:
            g = g * scale.unsqueeze(1).to(g.dtype)
        if g.dtype != dt:
            g = g.to(dt)
        main = torch.cuda.current_stream(g.device)
        if module._peer_dirty:
            ops.peer_barrier(pg, channel=0)       # two backward passes without a forward in between
        f = module.emb.fused
        row_bytes = module.emb_dim * module.emb.weight.element_size()
        fused_ok = (f is not None and not f.accumulate and f.kind in ("rowwise_adagrad", "sgd")
                    and row_bytes in (256, 512) and module.fused_push is not False
                    and (pg.world > 1 or module.fused_push == "force"))
        if fused_ok:
            main.wait_event(ctx.plan_ready)
            module.emb.begin_update()
            module.emb._apply_fused(ctx.plan, g, 1, None, None, 0.0, peer_push=dict(
                group=pg, tables=module.num_tables, bags_per_table=b, rows_per_table=module.num_embeddings,
                push_ctas=module.push_ctas, tablewise=True))
            module.emb.end_update()
            res = None
        else:
            ops.peer_allgather_push(pg, g, int(pg.layout.off_grads))
            ops.peer_barrier(pg, channel=0)
            main.wait_event(ctx.plan_ready)
            res = module.emb.consume(ctx.plan, pg.grads_view(module.emb_dim, dt), slots_per_grad_row=1,
                                     guard=pg.status_word())
        ctx.plan = None
        pg.snapshot_status()
        module._peer_dirty = True
        return res, None, None, None


class TableWiseShardedEmbeddingBag(nn.Module):
    """Pooled multi-hot lookup into T equal-shaped tables partitioned TABLE-wise over the ranks of one NVLink box:
    table t lives whole on rank t % W (north_star item 4: "row-wise or table-wise sharding").  Peer-memory
    exchange only (CUDA).  forward(ids [T, b, P] int64 of THIS rank's batch, lengths [T, b]) -> [T, b, D].

    Against row-wise sharding of the same tables: every bag has ONE owner, so one pooled row per bag comes back
    (instead of W partial rows) and every gradient row is pushed to one rank (instead of all W): (W-1)/W * b T R
    bytes each way per GPU.  The price is balance: T must be a multiple of W for equal work, and one hot table
    is one hot GPU.  `emb.weight` holds this rank's ceil((T - rank) / W) tables stacked; gather_full_weight() /
    load_full_weight() translate to the unsharded [T, N, D] layout."""

    def __init__(self, num_embeddings: int, emb_dim: int, num_tables: int, mode: str = "sum", *, last_n: int = 0,
                 skip_pad: bool = False, pad_id: int = 0, group=None, comm: Optional[Collectives] = None,
                 dtype: torch.dtype = torch.float32, device=None,
                 fused_optimizer: Optional[FusedOptimizerConfig] = None):
        super().__init__()
        if mode not in ("sum", "mean"):
            raise ValueError("mode must be 'sum' or 'mean'")
        if comm is None:
            comm = Collectives(group) if dist.is_available() and dist.is_initialized() else SingleProcess()
        self.comm = comm
        w, r = comm.world, comm.rank
        if num_tables < w:
            raise ValueError(f"table-wise partitioning needs at least one table per rank ({num_tables} tables, {w} ranks)")
        self.num_embeddings, self.emb_dim, self.num_tables = int(num_embeddings), int(emb_dim), int(num_tables)
        self.mode, self.last_n, self.skip_pad, self.pad_id = mode, int(last_n), skip_pad, pad_id
        self.local_tables = (self.num_tables - r + w - 1) // w
        self.emb = EmbeddingTable(self.local_tables * self.num_embeddings, emb_dim, dtype=dtype, device=device)
        if fused_optimizer is not None:
            self.emb.enable_fused_optimizer(fused_optimizer)
        self._owner_of_table = torch.arange(self.num_tables) % w
        env = os.environ.get("RECEMB_PEER_FUSED_PUSH", "1")
        self.fused_push = False if env == "0" else ("force" if env == "force" else True)
        self.push_ctas = int(os.environ.get("RECEMB_PEER_PUSH_CTAS", "16"))
        self._peer: Optional[PeerGroup] = None
        self._peer_key = None
        self._peer_dirty = True

    def _side_stream(self, device):
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=device)
        return self._side

    def peer_group(self) -> PeerGroup:
        if self._peer is None:
            raise N.NativeError("peer exchange: call forward first (the group is built for its batch shape)")
        return self._peer

    def _ensure_peer_group(self, ids: torch.Tensor) -> None:
        w = self.emb.weight
        key = (w.data_ptr(), tuple(ids.shape))
        if self._peer is not None and self._peer_key == key:
            return
        self.close_peer()
        t, world = self.num_tables, self.comm.world
        b, p = ids.shape[0] // t, ids.shape[1]
        cap = -(-t // world) * b * p            # exact: a sender has that many slots for an owner at most
        cap = max(2, cap + (cap & 1))
        if world == 1:
            layout = arena_layout(1, cap, ids.shape[0], self.emb_dim, w.dtype)
            self._peer = PeerGroup.local(1, 0, [PeerGroup.new_arena(layout, w.device)], [w.detach()], layout)
        else:
            self._peer = PeerGroup.connect(w.detach(), cap=cap, bags_total=ids.shape[0], group=self.comm.group)
        self._peer_key = key
        self._peer_dirty = True
        # device-resident index tensors of the final pick (built here, outside any CUDA-graph capture)
        self._owner_idx = self._owner_of_table.to(w.device)
        self._table_idx = torch.arange(t, device=w.device)

    def close_peer(self) -> None:
        """Collective: unmap the peers' memory (before this rank's tables / arena may be freed)."""
        if self._peer is None:
            return
        torch.cuda.synchronize(self.emb.weight.device)
        if self.comm.world > 1:
            dist.barrier(group=self.comm.group)
        self._peer.close()
        if self.comm.world > 1:
            dist.barrier(group=self.comm.group)
        self._peer, self._peer_key = None, None

    def forward(self, ids: torch.Tensor, lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
        if ids.dtype != torch.int64 or ids.dim() != 3 or ids.shape[0] != self.num_tables:
            raise N.NativeError(f"ids must be int64 [num_tables = {self.num_tables}, batch, bag_size]")
        t, b, p = ids.shape
        flat = ids.contiguous().view(t * b, p)
        self._ensure_peer_group(flat)
        out = _TableWiseFn.apply(self.emb.grad_anchor(), flat,
                                 None if lengths is None else lengths.contiguous().view(t * b), self)
        return out.view(t, b, self.emb_dim)

    # ---------------------------------------------------------- checkpoints ----
    @torch.no_grad()
    def load_full_weight(self, full: torch.Tensor) -> None:
        """Take this rank's tables (t % W == rank) out of a global [T, N, D] tensor."""
        full = full.view(self.num_tables, self.num_embeddings, self.emb_dim)
        mine = full[self.comm.rank::self.comm.world].reshape(-1, self.emb_dim)
        self.emb.weight.copy_(mine.to(self.emb.weight.device, self.emb.weight.dtype))
        self._peer_dirty = True

    @torch.no_grad()
    def _gather_tables(self, local: torch.Tensor) -> torch.Tensor:
        """[local_tables * N, ...] of every rank -> global [T, N, ...] (local table j of rank s = table j * W + s)."""
        w, t, n = self.comm.world, self.num_tables, self.num_embeddings
        lt_max = -(-t // w)
        tail = tuple(local.shape[1:])
        pad = torch.zeros((lt_max * n,) + tail, dtype=local.dtype, device=local.device)
        pad[: self.local_tables * n] = local
        shards = self.comm.all_gather(pad).view((w, lt_max, n) + tail)
        full = torch.empty((t, n) + tail, dtype=pad.dtype, device=pad.device)
        for s in range(w):
            cnt = (t - s + w - 1) // w
            full[s::w] = shards[s, :cnt]
        return full

    @torch.no_grad()
    def gather_full_weight(self) -> torch.Tensor:
        """Global [T, N, D] under the unsharded module's layout (collective)."""
        return self._gather_tables(self.emb.weight.detach())

    @torch.no_grad()
    def gather_full_optimizer_state(self) -> dict:
        """Fused-optimizer state of the GLOBAL tables ([T, N] for row-wise Adagrad, [T, N, D] otherwise) + the step
        count, as RowWiseShardedEmbeddingBag.gather_full_optimizer_state (collective)."""
        if self.emb.fused is None:
            raise N.NativeError("gather_full_optimizer_state needs the fused optimizer mode")
        self.emb._ensure_state()
        out = {"kind": self.emb.fused.kind, "step": self.emb.fused_step}
        for name, key in (("opt_state1", "state1"), ("opt_state2", "state2")):
            buf = self.emb._buffers.get(name)
            out[key] = None if buf is None else self._gather_tables(buf)
        return out

    @torch.no_grad()
    def load_full_optimizer_state(self, state: dict) -> None:
        """Inverse of gather_full_optimizer_state: every rank keeps the tables it owns (resume on any world size)."""
        if self.emb.fused is None or state["kind"] != self.emb.fused.kind:
            raise N.NativeError("optimizer kind of the checkpoint does not match this module")
        self.emb._ensure_state()
        self.emb.fused_step = int(state["step"])
        for name, key in (("opt_state1", "state1"), ("opt_state2", "state2")):
            buf = self.emb._buffers.get(name)
            if buf is None:
                continue
            full = state[key]
            full = full.reshape((self.num_tables, self.num_embeddings) + tuple(full.shape[2:]))
            buf.copy_(full[self.comm.rank::self.comm.world].reshape(buf.shape).to(buf.device, buf.dtype))
