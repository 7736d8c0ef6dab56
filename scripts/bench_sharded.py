#!/usr/bin/env python
"""cfg 5 (BASELINE.json configs[4]): large-vocab row-wise sharded tables, weak scaling.

  torchrun --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 scripts/bench_sharded.py [--check]

Per GPU: T = 8 tables of 25 M x 128 bf16 rows (51.2 GB; W = 8 -> the named 8 x 200 M x 128),
local batch b = 8192 bags of P = 20 uniform int64 ids per table, pooled sum, fused row-wise
Adagrad.  --check first verifies sharded == unsharded at a reduced vocabulary with the real
NCCL collectives and CUDA kernels.  Prints one JSON line (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

os.environ.setdefault("RECEMB_PEER_BARRIER_TIMEOUT_S", "30")   # a benchmark must never sit in a dead barrier

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import recommendations_b200 as R  # noqa: E402
from recommendations_b200 import _native as N  # noqa: E402
from recommendations_b200.sharded import RowWiseShardedEmbeddingBag  # noqa: E402

T, B_LOCAL, P, DIM = 8, 8192, 20, 128
NVLINK_GBS = 770.0


def ids_for(rank, t_tables, b, p, seed=5000):
    g = torch.Generator().manual_seed(seed + 8 * rank)
    return torch.randint(-2 ** 63, 2 ** 63 - 1, (t_tables, b, p), generator=g, dtype=torch.int64)


PEER_FORWARD = None   # None: the module's default (pull on one rank, push otherwise)


def check(world, rank, dev, exchange=None):
    """sharded == unsharded against the CPU oracle: lives under tests/ (the only place besides smoke()
    and bench.py's CPU arm that may touch oracle/); run here under torchrun with the real peer mappings."""
    sys.path.insert(0, str(ROOT / "tests"))
    import check_sharded
    check_sharded.check(world, rank, dev, exchange, PEER_FORWARD, ids_for)


def run_cfg5(world, rank, dev, steps, warmup, rows_per_gpu=25_000_000, exchange=None, graph=False,
             full_check=False, comm=None, partition="row"):
    """Times the cfg 5 step on an already initialised process group; returns the result dict
    (every rank computes it, rank 0 prints it)."""
    class A:
        pass
    args = A()
    args.steps, args.warmup, args.rows_per_gpu = steps, warmup, rows_per_gpu
    n_rows = args.rows_per_gpu * world
    # comm: sharded.SingleProcess() runs the W = 1 anchor on one rank of a larger job (world = 1 here)
    if partition == "table":
        # the same 8 tables of 25 M x W rows, partitioned whole: table t on rank t % W (needs T % W == 0 for balance)
        from recommendations_b200.sharded import TableWiseShardedEmbeddingBag
        mod = TableWiseShardedEmbeddingBag(n_rows, DIM, T, dtype=torch.bfloat16, device=dev, comm=comm,
                                           fused_optimizer=R.FusedOptimizerConfig(kind="rowwise_adagrad", lr=0.05))
        mod.exchange, mod.peer_forward = "peer", "push"
        mod._pipelined = lambda: False
        mod._fused_push_ok = lambda g: world > 1
    else:
        mod = RowWiseShardedEmbeddingBag(n_rows, DIM, num_tables=T, dtype=torch.bfloat16, device=dev, exchange=exchange,
                                         peer_forward=PEER_FORWARD, comm=comm,
                                         fused_optimizer=R.FusedOptimizerConfig(kind="rowwise_adagrad", lr=0.05))
    ids_host = ids_for(rank, T, B_LOCAL, P).pin_memory()
    ids = ids_host.to(dev)
    grad = torch.randn(T, B_LOCAL, DIM, device=dev, dtype=torch.bfloat16)

    def step():
        out = mod(ids)
        out.backward(grad)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if full_check and mod.exchange == "peer" and partition == "row":
        # full size (51.2 GB per GPU), no unsharded copy possible: self-consistency instead.
        # (1) the forward is deterministic; (2) the pull forward (rows loaded from their owners, pooled
        # in slot order: the unsharded arithmetic) and the push forward (owner-side partial pools
        # rounded to bf16, summed in owner order) agree within the bf16 tolerance of north_star
        with torch.no_grad():
            want_mode = mod.peer_forward
            outs = {}
            for pf in ("pull", "push"):
                mod.peer_forward = pf
                a, b2 = mod(ids), mod(ids)
                assert torch.equal(a, b2), f"{pf} forward is not deterministic"
                outs[pf] = a.float()
            mod.peer_forward = want_mode
        torch.testing.assert_close(outs["push"], outs["pull"], rtol=1e-2, atol=8e-2)
        checksum = float(outs["pull"].double().sum())
        if rank == 0:
            print(f"[check] full size: pull == push within bf16 tolerance on {world} rank(s); "
                  f"rank-0 output checksum {checksum:.6e}", flush=True)
        del outs
    for _ in range(args.warmup):
        step()
    barrier()
    launches_per_step = None
    if graph:
        # the peer step has fixed shapes and no host synchronisation: capture fwd + bwd + update once
        if mod.exchange != "peer":
            raise SystemExit("--graph needs --exchange peer (the NCCL exchanges synchronise with the host)")
        l0 = N.launch_count()
        g = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread may query events while this thread captures
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            step()
        launches_per_step = N.launch_count() - l0
        eager_step, step = step, g.replay
        for _ in range(2):
            step()
        barrier()
    launches0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        tt = torch.tensor([ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    lookups = world * T * B_LOCAL * P
    row_bytes = DIM * 2
    nv_in = (world - 1) * T * B_LOCAL * (P * 8 + 2 * row_bytes)   # ids + partials (fwd) + grads (bwd), per GPU
    if mod.exchange == "peer":
        # rows pulled from remote owners (fwd) + entries pushed into my inbox + gathered gradients (bwd)
        remote = (world - 1) / world
        nv_in = int(remote * T * B_LOCAL * P * (row_bytes + 8) + (world - 1) * T * B_LOCAL * row_bytes)
        if mod.peer_forward == "push":
            # entries in + one partial row per (bag, remote owner) pair (zero rows included) + gathered gradients
            nv_in = int(remote * T * B_LOCAL * P * 8 + 2 * (world - 1) * T * B_LOCAL * row_bytes)
        if partition == "table":
            # entries in + one pooled row per bag of a remote owner back + its gradient row once
            nv_in = int(remote * T * B_LOCAL * (P * 8 + 2 * row_bytes))
    t_step = ms / args.steps * 1e-3
    mod_exchange = mod.exchange
    mod_peer_forward = mod.peer_forward
    mod_groups = mod.pipeline_groups if mod._pipelined() else 1
    mod_fused_push = bool(mod._fused_push_ok(grad.view(-1, DIM))) if mod_exchange == "peer" else False
    phases = None
    if graph:
        step = eager_step
    if mod.exchange == "peer":
        mod.peer_group().raise_on_status(synchronize=True)
    if os.environ.get("RECEMB_PHASES"):
        mod.phase_events = []
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        ev = mod.phase_events
        acc = {}
        for (n0, e0_), (n1, e1_) in zip(ev[:-1], ev[1:]):
            if n1 in ("start",):
                continue
            acc.setdefault(n1, []).append(e0_.elapsed_time(e1_))
        phases = {k: round(sum(v) / len(v), 4) for k, v in acc.items()}
        mod.phase_events = None
    mod.close_peer()
    del mod, grad, ids
    if graph:
        del g, step, eager_step
    torch.cuda.empty_cache()
    return {
            "metric": "embedding_lookups_per_sec_fwd_bwd", "value": lookups / t_step, "unit": "lookups/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"cfg5: {T} tables x {n_rows} x {DIM} bf16 {partition}-wise sharded over {world} GPU(s), "
                                   f"b={B_LOCAL}/GPU, P={P}, pooled sum, fused row-wise Adagrad",
                       "table_bytes_per_gpu": T * args.rows_per_gpu * row_bytes, "exchange": mod_exchange,
                       "peer_forward": mod_peer_forward if mod_exchange == "peer" else None,
                       "pipeline_groups": mod_groups, "fused_gradient_push": mod_fused_push},
            "nvlink": {"bytes_in_per_gpu_per_step": nv_in, "achieved_gbs": nv_in / t_step / 1e9,
                       "peak_gbs": NVLINK_GBS, "frac": nv_in / t_step / 1e9 / NVLINK_GBS},
            "gpu_launches": (launches_per_step * args.steps if graph else N.launch_count() - launches0),
            "cuda_graph": bool(graph), "phases_ms": phases}


def phase_bench(world, rank, dev, rows_per_gpu=25_000_000, reps=10):
    """Peer exchange, one phase at a time (all ranks run the same phase together, `reps` launches
    back to back between two events): where the step time goes at this W."""
    from recommendations_b200 import ops
    n_rows = rows_per_gpu * world
    mod = RowWiseShardedEmbeddingBag(n_rows, DIM, num_tables=T, dtype=torch.bfloat16, device=dev, exchange="peer",
                                     pipeline_groups=1,
                                     fused_optimizer=R.FusedOptimizerConfig(kind="rowwise_adagrad", lr=0.05))
    ids = ids_for(rank, T, B_LOCAL, P).to(dev)
    grad = torch.randn(T, B_LOCAL, DIM, device=dev, dtype=torch.bfloat16)
    out = mod(ids)
    out.backward(grad)          # builds the group, fills inbox / gradient buffer once
    pg = mod.peer_group()
    flat = ids.view(T * B_LOCAL, P)
    batching = dict(bags_per_table=B_LOCAL, num_tables=T)
    total_rows = mod.emb.weight.shape[0]
    res = {}

    def timed(name, fn):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = round(float(t.item()), 4)

    timed("pull_pool_fwd", lambda: ops.peer_pool_fwd(pg, flat, num_rows=n_rows, dim=DIM, dtype=torch.bfloat16,
                                                     **batching))
    timed("bucket_push", lambda: ops.peer_bucket_push(pg, flat, num_rows=n_rows, **batching))
    timed("grads_push", lambda: ops.peer_allgather_push(pg, grad, int(pg.layout.off_grads)))
    parts = pg.parts_view(DIM, torch.bfloat16)
    ops.peer_bucket_push(pg, flat, num_rows=n_rows, **batching)
    timed("pool_push(owner pools inbox, stores partial rows)", lambda: ops.peer_pool_push(pg, DIM, torch.bfloat16))
    timed("sum_partials", lambda: ops.sum_partials(parts))
    timed("barrier", lambda: ops.peer_barrier(pg, 0))
    plan = ops.peer_plan(pg, total_rows)
    timed("plan(unpack+sort)", lambda: ops.peer_plan(pg, total_rows))
    gv = pg.grads_view(DIM, torch.bfloat16)
    timed("apply(seg+adagrad)", lambda: mod.emb.consume(plan, gv, slots_per_grad_row=1))
    if mod.emb.fused is not None:
        hp = ops.make_optim_params(lr=0.05, eps=1e-10)
        s1 = mod.emb._buffers.get("opt_state1")
        for ctas in (8, 32, 128):
            timed(f"fused push+apply (push_ctas={ctas})", lambda: ops.peer_bwd_apply_fused(
                plan, grad.view(-1, DIM), group=pg, table=mod.emb.weight.data, update=N.UPD_ROWWISE_ADAGRAD, state1=s1,
                hp=hp, tables=T, bags_per_table=B_LOCAL, rows_per_table=mod.local_rows, push_ctas=ctas))
    pg.raise_on_status(synchronize=True)
    mod.close_peer()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--rows-per-gpu", type=int, default=25_000_000)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--exchange", default=None, choices=["route", "gather", "peer"])
    ap.add_argument("--phase-bench", action="store_true", help="time the phases of the peer exchange one by one")
    ap.add_argument("--peer-forward", default=None, choices=["pull", "push"])
    ap.add_argument("--graph", action="store_true", help="replay the step from one CUDA graph (peer exchange only)")
    ap.add_argument("--partition", default="row", choices=["row", "table"])
    args = ap.parse_args()
    global PEER_FORWARD
    PEER_FORWARD = args.peer_forward
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.check:
        check(world, rank, dev, args.exchange)
    if args.phase_bench:
        res = {"phase_ms": phase_bench(world, rank, dev, args.rows_per_gpu), "n_gpus": world}
    else:
        res = run_cfg5(world, rank, dev, args.steps, args.warmup, args.rows_per_gpu, args.exchange, args.graph,
                       full_check=args.check, partition=args.partition)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
