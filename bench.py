#!/usr/bin/env python
"""bench.py -- embedding lookups/sec fwd+bwd on BASELINE.json configs[1]:
"LTHM embedding fwd+bwd on 1xB200, batch 8192, history len 200, 10 tables of 1Mx64 fp32".

A step = one pass of the hot path over one synthetic batch: for each of the 10 tables a
sequence gather of [8192, 200] ids (FlatEmbedding semantics, commons/layers.py:56-61) and its
backward with the reference's optimizer fused (element-wise Adagrad lr 0.5,
embedding_module_gen.py:137): sort-based dedup plan + segmented reduction + update.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N > 1 (torchrun, one rank per GPU): tables of this size are replicated by the reference
(pure data parallel) -- every rank runs the same per-GPU workload on its own
batch (weak scaling, no data-path collective); the row-wise sharded large-vocab path is
measured by --workload cfg5.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

T_TABLES, BATCH, HIST, ROWS, DIM = 10, 8192, 200, 1_000_000, 64
LR, EPS = 0.5, 1e-10
METRIC = "embedding_lookups_per_sec_fwd_bwd"
UNIT = "lookups/s"
WORKLOAD = ("cfg2: LTHM embedding fwd+bwd, batch 8192 x history 200, 10 tables 1Mx64 fp32, sequence gather + "
            "element-wise Adagrad(lr=0.5)")


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, from the
    last `ncu --set full` capture committed under profiles/ (null if none)."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(p.read_text())["seg_kernel_level0_dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def host_ids(t: int, rank: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(1000 + t + 100 * rank)
    return torch.randint(-2 ** 63, 2 ** 63 - 1, (BATCH * HIST,), generator=g, dtype=torch.int64)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                 "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (ts, r) in self.rows if t0 <= ts <= t1 + 0.2] or [r for (_, r) in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------ CPU baseline ----
def cpu_baseline_run(steps: int, warmup: int, ids_tables, weights, grad: torch.Tensor):
    """The reference's own CPU path for the WHOLE config -- every one of the T tables per step:
    remainder -> F.embedding -> autograd -> torch.optim.Adagrad(lr=0.5) with dense gradients, one
    module + optimizer per table as the reference holds them (oracle port; all host threads).
    Returns the per-step wall times."""
    from oracle import embedding_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    mods = [O.FlatTableCPU(w) for w in weights]
    opts = [torch.optim.Adagrad(m.parameters(), lr=LR) for m in mods]
    g3 = grad.view(BATCH, HIST, DIM)

    def step():
        for m, o, ids in zip(mods, opts, ids_tables):
            O.cpu_train_step(m, o, ids.view(BATCH, HIST), g3)
    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port, since the
    reference is pure Python over torch and cannot travel), rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ids = [host_ids(t) for t in range(T_TABLES)]
    weights = []
    for t in range(T_TABLES):
        torch.manual_seed(1234 + t)
        weights.append(torch.randn(ROWS, DIM))
    grad = torch.randn(BATCH * HIST, DIM, generator=torch.Generator().manual_seed(4321))
    times = cpu_baseline_run(args.steps, args.warmup, ids, weights, grad)
    total = sum(times)
    value = T_TABLES * BATCH * HIST * args.steps / total
    sample = (f"the whole config per step: all {T_TABLES} tables (FlatEmbedding {ROWS}x{DIM} fp32, ids [{BATCH},{HIST}] "
              f"each), fwd + dense bwd + torch.optim.Adagrad, {args.warmup} warm-up + {args.steps} timed")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "lookups_per_step_per_gpu": T_TABLES * BATCH * HIST},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------- B200 arm ----
def bind_to_gpu_numa(local_rank: int):
    """Multi-GPU runs: pin this rank to the CPUs of the NUMA node its GPU hangs off BEFORE the pinned id buffers
    are allocated (first touch puts them on that node), so that 8 ranks x 131 MB of ids per step do not all
    cross one socket's memory controller / the inter-socket link on their way to PCIe.  The node comes from
    sysfs (/sys/bus/pci/devices/<bdf>/numa_node), the CPU set from NVML's affinity mask as a fallback.  Best
    effort: returns a short description of what it bound to, or None (no NUMA information, mask as wide as the
    machine, affinity not settable)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            ent = vis.split(",")[local_rank].strip()
            if not ent.isdigit():
                return None
            idx = int(ent)
        handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        allowed = os.sched_getaffinity(0)
        cpus, how = set(), None
        try:
            bus = pynvml.nvmlDeviceGetPciInfo(handle).busId
            bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
            if len(bus.split(":")[0]) == 8:         # NVML prints an 8-digit domain, sysfs a 4-digit one
                bus = bus[4:]
            node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
            if node >= 0:
                for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                    lo, _, hi = part.partition("-")
                    cpus.update(range(int(lo), int(hi or lo) + 1))
                how = f"numa node {node}"
        except Exception:
            cpus = set()
        if not cpus:
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
            cpus = {i for i in range(ncpu) if (words[i // 64] >> (i % 64)) & 1}
            how = "nvml affinity"
        cpus &= allowed
        if len(cpus) < 2 or cpus == allowed:
            return None
        os.sched_setaffinity(0, cpus)
        return f"{how}: {len(cpus)} of {len(allowed)} cpus"
    except Exception:
        return None


T_CFG5 = 8   # tables of the sharded config (scripts/bench_sharded.py)


def run_b200(args):
    import torch.distributed as dist

    import recommendations_b200  # noqa: F401
    from recommendations_b200 import _native as N
    from recommendations_b200 import ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = bind_to_gpu_numa(local) if world > 1 and os.environ.get("RECEMB_BENCH_NUMA", "1") != "0" else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    lib = N.load()
    n = BATCH * HIST

    # ---- resident state: the 10 tables stacked as one [T*ROWS, DIM] allocation (table-batched
    # mode: one launch per phase covers all tables), Adagrad state, ids, upstream grads, outputs
    gen = torch.Generator(device=dev)
    table_all = torch.empty(T_TABLES * ROWS, DIM, device=dev)
    for t in range(T_TABLES):
        gen.manual_seed(1234 + t)
        table_all[t * ROWS:(t + 1) * ROWS].normal_(generator=gen)
    state_all = torch.zeros(T_TABLES * ROWS, DIM, device=dev)
    ids_host = torch.cat([host_ids(t, rank) for t in range(T_TABLES)]).pin_memory()
    ids_dev = ids_host.to(dev)
    gen.manual_seed(4321)
    grads = torch.randn(T_TABLES * n, DIM, device=dev, generator=gen)
    outs = torch.empty(T_TABLES * n, DIM, device=dev)
    n_all, rows_all = T_TABLES * n, T_TABLES * ROWS
    plan_bytes = int(lib.recemb_bwd_plan_bytes(n_all, rows_all))
    ws_bytes = int(lib.recemb_bwd_apply_workspace_bytes(n_all, DIM))
    plan_buf = torch.empty(plan_bytes, dtype=torch.uint8, device=dev)
    ws_buf = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    hp = ops.make_optim_params(lr=LR, eps=EPS)

    ev_apply, ev_gather = [], []  # (start, stop) around the two dominant launches

    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(device=dev)

    def step(timed: bool):
        # the backward plan (hash + radix sort of the slots) only depends on the ids: it is built on
        # a second stream while the forward gather streams rows on the first
        if args.overlap_plan:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                plan = ops.BackwardPlan.build(ids_dev, num_rows=ROWS, ids_per_table=n, buf=plan_buf)
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        ops.gather_fwd(table_all, ids_dev, out=outs, ids_per_table=n)
        if timed:
            b.record()
            ev_gather.append((a, b))
        if args.overlap_plan:
            main.wait_stream(side)
        else:
            plan = ops.BackwardPlan.build(ids_dev, num_rows=ROWS, ids_per_table=n, buf=plan_buf)
        if timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()  # creates the underlying cudaEvent_t; the library re-records both
            b.record()  # around its level-0 launch
            lib.recemb_time_next_apply(a.cuda_event, b.cuda_event)
            ev_apply.append((a, b))
        ops.bwd_apply(plan, grads, table=table_all, update=N.UPD_ADAGRAD, state1=state_all, hp=hp,
                      workspace=ws_buf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step(True)
    e1.record()
    barrier()
    t_wall1 = time.perf_counter()
    launches = N.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    if world > 1:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    value = world * n_all * args.steps / (ms_total * 1e-3)

    # unique rows over the stacked table (for the algorithmic byte count of the dominant kernel)
    plan = ops.BackwardPlan.build(ids_dev, num_rows=ROWS, ids_per_table=n, buf=plan_buf)
    n_valid, uniq = (int(v) for v in plan.counters.cpu())
    assert n_valid == n_all
    row_bytes = DIM * 4
    apply_ms = [a.elapsed_time(b) for a, b in ev_apply]
    gather_ms = [a.elapsed_time(b) for a, b in ev_gather]
    apply_bytes = n_all * (8 + row_bytes) + uniq * (2 * row_bytes + 2 * DIM * 4)
    gather_bytes = n_all * (8 + 2 * row_bytes)
    peak, peak_kind = peaks()
    apply_gbs = apply_bytes / (statistics.mean(apply_ms) * 1e-3) / 1e9
    gather_gbs = gather_bytes / (statistics.mean(gather_ms) * 1e-3) / 1e9
    step_bytes = gather_bytes + apply_bytes
    roofline = {
        "bound": "hbm", "kernel": "seg_kernel<16,1,float,float,L0> (segmented reduce + fused Adagrad), "
                                  "one launch per step over all 10 tables",
        "achieved": apply_gbs, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
        "frac": apply_gbs / peak, "traffic": ncu_traffic(),
        "algorithmic_bytes_per_launch": apply_bytes, "avg_launch_ms": statistics.mean(apply_ms),
        "gather_kernel": {"achieved": gather_gbs, "frac": gather_gbs / peak,
                          "algorithmic_bytes_per_launch": gather_bytes,
                          "avg_launch_ms": statistics.mean(gather_ms)},
        "whole_step": {"algorithmic_bytes": step_bytes,
                       "achieved": step_bytes / (ms_total / args.steps * 1e-3) / 1e9,
                       "frac": step_bytes / (ms_total / args.steps * 1e-3) / 1e9 / peak},
    }

    # ---- the same step through the drop-in nn.Module (EmbeddingCollection: 10 reference-style tables,
    # per-table state_dict keys, one launch per phase) with autograd: module(ids); out.backward(g)
    module_path = run_module_path(args, N, dev, world, ids_host, ids_dev, grads, barrier)

    # ---- end to end through the C-ABI host entry point (HOST ids in, counters out) ----
    e2e_cabi = run_e2e(args, lib, N, ops, dev, world, table_all, state_all, ids_host, grads, outs, hp, barrier)
    e2e = dict(module_path.pop("e2e"))
    e2e["c_abi_host_entry"] = e2e_cabi

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # NB: the tables were already updated by the timed steps -- any weights do for a timing
        w_cpu = [table_all[t * ROWS:(t + 1) * ROWS].cpu() for t in range(T_TABLES)]
        ids_cpu = [ids_host[t * n:(t + 1) * n].clone() for t in range(T_TABLES)]
        times = cpu_baseline_run(3, 1, ids_cpu, w_cpu, grads[:n].cpu())
        best = min(times)
        cpu = {"value": n_all / best, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"the whole config: all {T_TABLES} tables per step (same ids / weights as the GPU run), fwd + "
                         f"dense bwd + torch.optim.Adagrad on CPU, 1 warm-up + 3 timed, best; "
                         f"os.cpu_count()={os.cpu_count()}"}
        del w_cpu, ids_cpu

    del table_all, state_all, grads, outs, plan_buf, ws_buf
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs (kernel-level, inputs resident): cfg 3 ranker pooled + tcgen05
    # interaction, cfg 4 long-history Zipf + row-wise Adagrad, the k-shift series, the plan alone
    configs = None
    if world == 1 and not args.no_configs:
        try:
            sys.path.insert(0, str(ROOT / "scripts"))
            import bench_configs
            configs = bench_configs.run(["cfg2plan", "kshift", "cfg3", "cfg4", "frontend"], quiet=True)
        except Exception as exc:
            configs = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- the path that really exchanges data over NVLink (cfg 5: row-wise sharded large-vocab tables),
    # same process group; at world 1 it is the W = 1 anchor of the scaling series
    sharded = None
    if not args.no_sharded:
        try:
            sys.path.insert(0, str(ROOT / "scripts"))
            import bench_sharded
            # peer exchange: rows pulled / entries and gradients pushed through NVLink peer memory
            # inside the lookup kernels, whole step replayed from one CUDA graph
            sharded = bench_sharded.run_cfg5(world, rank, dev, args.steps, args.warmup, exchange="peer",
                                             graph=True)
            if world > 1:
                # the W = 1 anchor of the weak-scaling series in the SAME run: rank 0 alone times the step on
                # its 1/W slice of the job (same per-GPU table bytes and batch), the others wait
                if rank == 0:
                    from recommendations_b200.sharded import SingleProcess
                    anchor = bench_sharded.run_cfg5(1, 0, dev, args.steps, args.warmup, exchange="peer",
                                                    graph=True, comm=SingleProcess())
                    sharded["w1_anchor_ms_per_step"] = anchor["ms_per_step"]
                    sharded["efficiency_vs_w1"] = anchor["ms_per_step"] / sharded["ms_per_step"]
                dist.barrier()
                if T_CFG5 % world == 0:
                    # the same tables partitioned TABLE-wise (table t whole on rank t % W): one pooled row per bag
                    # back, every gradient row pushed to one rank -- north_star item 4 names both partitionings
                    tw = bench_sharded.run_cfg5(world, rank, dev, args.steps, args.warmup, exchange="peer",
                                                graph=True, partition="table")
                    if rank == 0:
                        tw["efficiency_vs_w1"] = sharded["w1_anchor_ms_per_step"] / tw["ms_per_step"]
                    sharded["tablewise"] = {k: tw[k] for k in ("value", "ms_per_step", "nvlink", "config", "gpu_launches")
                                            } | ({"efficiency_vs_w1": tw.get("efficiency_vs_w1")} if rank == 0 else {})
        except Exception as exc:  # the headline line must survive a failure of the extra run
            sharded = {"error": f"{type(exc).__name__}: {exc}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "lookups_per_step_per_gpu": n_all, "unique_rows_per_step": uniq,
                       "layout": "tables stacked [10*1M, 64]; table-batched launches (ids_per_table)",
                       "l2": "inputs larger than L2: 419 MB out + 419 MB grad + 256 MB table per table vs 126 MB",
                       "multi_gpu": "replicas (tables replicated as in the reference), weak scaling"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "module_path": module_path, "configs": configs, "sharded_cfg5": sharded,
            "clocks": clocks,
            "host": {"cpu_count": os.cpu_count(),
                     "rank0_bound_to": numa_cpus},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_module_path(args, N, dev, world, ids_host, ids_dev, grads, barrier):
    """cfg 2 through recommendations_b200.EmbeddingCollection -- what INTEGRATION.md tells a maintainer to
    use in place of ten FlatEmbedding modules -- (a) device-resident ids, CUDA events; (b) end to end:
    the step's ids start in PINNED HOST memory and are copied inside the timed region (double-buffered on
    a copy stream so step s+1's copy runs under step s), and the host reads a slice of the step's output
    back every step."""
    import torch.distributed as dist

    import recommendations_b200 as R
    n = BATCH * HIST
    n_all = T_TABLES * n
    coll = R.EmbeddingCollection(T_TABLES, ROWS, DIM, device=dev,
                                 fused_optimizer=R.FusedOptimizerConfig(kind="adagrad", lr=LR, eps=EPS))
    g4 = grads.view(T_TABLES, BATCH, HIST, DIM)
    ids3 = ids_dev.view(T_TABLES, BATCH, HIST)

    def mstep():
        coll(ids3).backward(g4)

    def reduce_max(ms):
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        return ms

    for _ in range(args.warmup):
        mstep()
    barrier()
    c0 = N.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        mstep()
    e1.record()
    barrier()
    ms_dev = reduce_max(e0.elapsed_time(e1))
    launches = N.launch_count() - c0

    main = torch.cuda.current_stream(dev)
    copy_stream = torch.cuda.Stream(device=dev)
    bufs = [torch.empty_like(ids_dev) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    probe = torch.zeros(2, 16).pin_memory()

    def run(steps):
        seen = 0.0
        for s in range(steps):
            k = s % 2
            if s >= 2:
                copy_stream.wait_event(done[k])       # step s-2 has finished reading bufs[k]
            with torch.cuda.stream(copy_stream):
                bufs[k].copy_(ids_host, non_blocking=True)
                copied[k].record(copy_stream)
            main.wait_event(copied[k])
            out = coll(bufs[k].view(T_TABLES, BATCH, HIST))
            out.backward(g4)
            probe[k].copy_(out.view(-1)[:16], non_blocking=True)
            done[k].record(main)
            if s > 0:
                done[1 - k].synchronize()              # the host consumes step s-1's result
                seen += float(probe[1 - k][0])
        done[(steps - 1) % 2].synchronize()
        return seen + float(probe[(steps - 1) % 2][0])

    run(max(2, args.warmup))
    barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(args.steps)
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ms_e2e = reduce_max(e0.elapsed_time(e1))
    del coll
    torch.cuda.empty_cache()
    return {"ms_per_step": ms_dev / args.steps, "value": world * n_all * args.steps / (ms_dev * 1e-3), "unit": UNIT,
            "gpu_launches": launches,
            "api": "recommendations_b200.EmbeddingCollection(10, 1M, 64, fused adagrad): out = module(ids); "
                   "out.backward(grad) -- autograd.Function over the C ABI, plan built on a side stream during forward",
            "e2e": {"value": world * n_all * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": n_all * 8, "d2h_bytes_per_step": 64,
                    "ms_per_step": ms_e2e / args.steps, "wall_ms_per_step": wall * 1e3 / args.steps,
                    "api": "EmbeddingCollection.forward / backward on ids copied from pinned host memory every step "
                           "(copy stream, double-buffered); 64 bytes of the step's output read back by the host every "
                           "step (a training step's outputs and updated tables stay on the device)"}}


def run_e2e(args, lib, N, ops, dev, world, table_all, state_all, ids_host, grads, outs, hp, barrier):
    """Same step through recemb_flat_step_host (one table-batched call per step): the step's ids
    start in PINNED HOST memory and are copied inside the timed region; the step's result
    (valid / unique counters) is read back by the host every step.  Two streams alternate
    between steps so the H2D copy of step s+1 runs while step s computes
    (wait_event_after_copy orders the kernels); `serial` is the same without that overlap."""
    import ctypes as C
    import torch.distributed as dist
    n = BATCH * HIST
    n_all, rows_all = T_TABLES * n, T_TABLES * ROWS
    plan_bytes = int(lib.recemb_bwd_plan_bytes(n_all, rows_all))
    ws_bytes = int(lib.recemb_bwd_apply_workspace_bytes(n_all, DIM))
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    plan_streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    scratch = [torch.empty(n_all, dtype=torch.int64, device=dev) for _ in range(2)]
    plans = [torch.empty(plan_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    wss = [torch.empty(ws_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    counters = torch.zeros(2, 2, dtype=torch.int64).pin_memory()
    done = [torch.cuda.Event() for _ in range(2)]
    for i in range(2):
        done[i].record(streams[i])  # materialise the cudaEvent_t handles

    def enqueue(s, overlap):
        k = s % 2
        N.check(lib.recemb_flat_step_host(
            ids_host.data_ptr(), n_all, n, scratch[k].data_ptr(), table_all.data_ptr(), ROWS, DIM, N.F32,
            outs.data_ptr(), grads.data_ptr(), N.UPD_ADAGRAD, state_all.data_ptr(), None, C.byref(hp),
            plans[k].data_ptr(), plan_bytes, wss[k].data_ptr(), ws_bytes, counters[k].data_ptr(),
            done[1 - k].cuda_event if overlap else None,
            plan_streams[k].cuda_stream if args.overlap_plan else None, dev.index, streams[k].cuda_stream),
            "recemb_flat_step_host")
        done[k].record(streams[k])

    def run(steps, overlap):
        checked = 0
        for s in range(steps):
            if not overlap and s > 0:
                streams[(s - 1) % 2].synchronize()
            enqueue(s, overlap)
            if overlap and s > 0:
                streams[(s - 1) % 2].synchronize()  # host reads step s-1's result
            if s > 0:
                checked += int(counters[(s - 1) % 2, 0])
        streams[(steps - 1) % 2].synchronize()
        checked += int(counters[(steps - 1) % 2, 0])
        return checked

    out = {}
    for name, overlap in (("serial", False), ("pipelined", True)):
        run(max(2, args.warmup), overlap)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        checked = run(args.steps, overlap)
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        assert checked == n_all * args.steps, "e2e steps did not process every lookup"
        out[name] = (ms, wall)
    ms, wall = out["pipelined"]
    return {"value": world * n_all * args.steps / (ms * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": n_all * 8, "d2h_bytes_per_step": 16,
            "ms_per_step": ms / args.steps, "wall_ms_per_step": wall * 1e3 / args.steps,
            "serial_ms_per_step": out["serial"][0] / args.steps,
            "api": "recemb_flat_step_host (C ABI; pinned host ids in, counters out, every step); H2D of "
                   "step s+1 overlaps the kernels of step s on a second stream; the plan's sort runs on a "
                   "third stream next to the gather"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg 3 / cfg 4 / k-shift sub-measurements")
    ap.add_argument("--no-overlap-plan", dest="overlap_plan", action="store_false",
                    help="build the backward plan after the gather on the same stream (default: concurrently)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        print(f"[bench] note: warmup {args.warmup} < 3", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
