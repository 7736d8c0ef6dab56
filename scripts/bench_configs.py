#!/usr/bin/env python
"""Kernel-level measurements of the BASELINE.json configs that are not the bench.py headline
(cfg 3 ranker pooled + dot interaction, cfg 4 long-history Zipf + row-wise Adagrad, and the
k-shift secondary series of cfg 2).  One JSON line per measurement; inputs resident in HBM,
CUDA events, 3 warm-ups, inputs larger than L2.  Algorithmic bytes follow SURVEY.md section 8(d)."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from recommendations_b200 import _native as N  # noqa: E402
from recommendations_b200 import ops  # noqa: E402

DEV = torch.device("cuda:0")
PEAK = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


RESULTS = []   # every report() of this process, in order (bench.py collects them into its JSON line)
QUIET = False


def report(name, ms, alg_bytes, lookups=None, **extra):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    line = {"name": name, "ms": round(ms, 4), "algorithmic_GB": round(alg_bytes / 1e9, 4),
            "achieved_GBs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / PEAK, 3)}
    if lookups:
        line["G_lookups_per_s"] = round(lookups / (ms * 1e-3) / 1e9, 3)
    line.update(extra)
    RESULTS.append(line)
    if not QUIET:
        print(json.dumps(line), flush=True)


def uniform_ids(n, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randint(-2 ** 63, 2 ** 63 - 1, (n,), generator=g, dtype=torch.int64, device=DEV)


def kshift_series():
    n_rows, dim, k, n = 1_000_000, 64, 8, 8192 * 200
    w = torch.randn(n_rows, dim, device=DEV)
    state = torch.zeros_like(w)
    ids = uniform_ids(n, 1000)
    grad = torch.randn(n, dim, device=DEV)
    r = dim * 4
    ms = timeit(lambda: ops.kshift_fwd(w, ids, k, N.EPI_RSQRT_K))
    report("cfg2-kshift(k=8) fwd: fused bag, 1M x 64 fp32, ids [8192,200]", ms, n * (8 + k * r + r), n * k)
    plan_buf = torch.empty(int(N.load().recemb_bwd_plan_bytes(n * k, n_rows)), dtype=torch.uint8, device=DEV)
    hp = ops.make_optim_params(lr=0.5, eps=1e-10)

    def bwd():
        # the 1/sqrt(k) epilogue backward is folded into the segmented reduction (grad_div)
        plan = ops.BackwardPlan.build(ids, num_rows=n_rows, hash_mode=N.HASH_ROTL_FLOORMOD, slots_per_id=k,
                                      buf=plan_buf)
        ops.bwd_apply(plan, grad, table=w, update=N.UPD_ADAGRAD, state1=state, slots_per_grad_row=k, hp=hp,
                      grad_div=k ** 0.5)
    ms = timeit(bwd)
    plan = ops.BackwardPlan.build(ids, num_rows=n_rows, hash_mode=N.HASH_ROTL_FLOORMOD, slots_per_id=k, buf=plan_buf)
    uniq = int(plan.counters.cpu()[1])
    report("cfg2-kshift(k=8) bwd: plan + segmented reduce (1/sqrt(k) folded in) + Adagrad (50% of shift>=1 lookups collapse)",
           ms, n * 8 + n * k * r + uniq * 4 * r, n * k, unique_rows=uniq)


def cfg3():
    b, f, dim, p, n_rows = 16384, 26, 128, 20, 1_000_000
    w = torch.randn(f * n_rows, dim, device=DEV, dtype=torch.bfloat16)
    state = torch.zeros(f * n_rows, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(3000)
    ids = torch.randint(0, n_rows, (f * b, p), generator=g, device=DEV, dtype=torch.int64)
    lengths = torch.randint(1, p + 1, (f * b,), generator=g, device=DEV, dtype=torch.int32)
    valid = int(lengths.sum())
    dense = torch.randn(b, dim, device=DEV, dtype=torch.bfloat16)
    r = dim * 2
    pool = lambda: ops.pool_fwd(w, ids, lengths=lengths, num_rows=n_rows, bags_per_table=b, num_tables=f,
                                hash_mode=N.HASH_IDENTITY)
    ms = timeit(pool)
    report("cfg3 pooled fwd: 26 x [1M,128] bf16 table-batched, B 16384, P 20 masked sum", ms,
           valid * (8 + r) + f * b * (r + 4), valid)
    pooled = pool().view(f, b, dim)
    feats = torch.cat([dense.unsqueeze(1), pooled.permute(1, 0, 2)], dim=1).contiguous()   # [B, 27, D]
    fp = f + 1
    ms = timeit(lambda: ops.dot_interaction_fwd(feats))
    report("cfg3 dot-interaction fwd (tcgen05): [16384,27,128] bf16 -> [16384,351]", ms,
           b * fp * r + b * fp * (fp - 1) // 2 * 2, flops_G=round(2 * b * fp * (fp - 1) // 2 * dim / 1e9, 3),
           tensor_GFLOPs_issued=round(2 * (b / 4) * 128 * 128 * dim / 1e9, 2))
    # the forward as one pipeline: pooled rows written feature-interleaved straight into the interaction's
    # input (recemb_layout.out_features), dense row copied into slot 0 -- vs lookup, permute + cat, interaction
    feats_buf = torch.empty(b, fp, dim, device=DEV, dtype=torch.bfloat16)

    def fwd_pipeline():
        feats_buf[:, 0].copy_(dense)
        ops.pool_fwd(w, ids, lengths=lengths, num_rows=n_rows, bags_per_table=b, num_tables=f,
                     hash_mode=N.HASH_IDENTITY, out=feats_buf, out_features=fp, out_feature_offset=1)
        return ops.dot_interaction_fwd(feats_buf)

    def fwd_two_step():
        pooled_ = pool().view(f, b, dim)
        return ops.dot_interaction_fwd(torch.cat([dense.unsqueeze(1), pooled_.permute(1, 0, 2)], dim=1))
    ms = timeit(fwd_pipeline)
    ms2 = timeit(fwd_two_step)
    report("cfg3 fwd pipeline: pooled lookup -> [B,27,128] in place -> tcgen05 interaction (no cat / permute)", ms,
           valid * (8 + r) + f * b * 4 + 2 * b * fp * r + b * fp * (fp - 1) // 2 * 2, valid,
           ms_lookup_cat_interaction=round(ms2, 4))
    go = torch.randn(b, fp * (fp - 1) // 2, device=DEV, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.dot_interaction_bwd(feats, go))
    report("cfg3 dot-interaction bwd (tcgen05)", ms, b * fp * (fp - 1) // 2 * 2 + 2 * b * fp * r)
    gpool = torch.randn(f * b, dim, device=DEV, dtype=torch.bfloat16)
    plan_buf = torch.empty(int(N.load().recemb_bwd_plan_bytes(f * b * p, f * n_rows)), dtype=torch.uint8, device=DEV)
    hp = ops.make_optim_params(lr=0.05, eps=1e-10)

    def bwd():
        plan = ops.BackwardPlan.build(ids, num_rows=n_rows, hash_mode=N.HASH_IDENTITY, bag_size=p, lengths=lengths,
                                      ids_per_table=b * p, num_tables=f, buf=plan_buf)
        ops.bwd_apply(plan, gpool, table=w, update=N.UPD_ROWWISE_ADAGRAD, state1=state, slots_per_grad_row=p, hp=hp)
    ms = timeit(bwd)
    plan = ops.BackwardPlan.build(ids, num_rows=n_rows, hash_mode=N.HASH_IDENTITY, bag_size=p, lengths=lengths,
                                  ids_per_table=b * p, num_tables=f, buf=plan_buf)
    uniq = int(plan.counters.cpu()[1])
    report("cfg3 pooled bwd: plan + segmented reduce + row-wise Adagrad", ms,
           f * b * p * 8 + f * b * r + uniq * (2 * r + 8), valid, unique_rows=uniq)
    ms = timeit(lambda: ops.BackwardPlan.build(ids, num_rows=n_rows, hash_mode=N.HASH_IDENTITY, bag_size=p,
                                               lengths=lengths, ids_per_table=b * p, num_tables=f, buf=plan_buf))
    report("cfg3 plan alone (key kernel + library radix sort)", ms, f * b * p * 8 * 2, valid)


def zipf_rows(n, n_rows, alpha, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    ranks = torch.arange(1, n_rows + 1, device=DEV, dtype=torch.float64)
    cdf = torch.cumsum(ranks.pow(-alpha), 0)
    cdf /= cdf[-1].clone()
    u = torch.rand(n, generator=g, device=DEV, dtype=torch.float64)
    r = torch.searchsorted(cdf, u).clamp_(max=n_rows - 1)
    perm = torch.randperm(n_rows, generator=g, device=DEV)
    return perm[r]


def cfg4():
    t, b, l, n_rows, dim = 10, 4096, 1024, 1_000_000, 64
    n = b * l
    w = torch.randn(t * n_rows, dim, device=DEV)
    state = torch.zeros(t * n_rows, device=DEV)
    ids = torch.cat([zipf_rows(n, n_rows, 1.05, 2000 + i) for i in range(t)])
    out = torch.empty(t * n, dim, device=DEV)
    grad = torch.randn(t * n, dim, device=DEV)
    r = dim * 4
    ms_f = timeit(lambda: ops.gather_fwd(w, ids, out=out, ids_per_table=n, hash_mode=N.HASH_IDENTITY), iters=5)
    report("cfg4 fwd: 10 x [1M,64] fp32, B 4096 x L 1024, Zipf(1.05) rows, table-batched gather", ms_f,
           t * n * (8 + 2 * r), t * n)
    plan_buf = torch.empty(int(N.load().recemb_bwd_plan_bytes(t * n, t * n_rows)), dtype=torch.uint8, device=DEV)
    hp = ops.make_optim_params(lr=0.05, eps=1e-10)

    def bwd():
        plan = ops.BackwardPlan.build(ids, num_rows=n_rows, hash_mode=N.HASH_IDENTITY, ids_per_table=n, buf=plan_buf)
        ops.bwd_apply(plan, grad, table=w, update=N.UPD_ROWWISE_ADAGRAD, state1=state, hp=hp)
    ms_b = timeit(bwd, iters=5)
    plan = ops.BackwardPlan.build(ids, num_rows=n_rows, hash_mode=N.HASH_IDENTITY, ids_per_table=n, buf=plan_buf)
    uniq = int(plan.counters.cpu()[1])
    top = int(torch.bincount(ids[:n]).max())
    report("cfg4 bwd: plan + segmented reduce + row-wise Adagrad (hot rows -> multi-level records)", ms_b,
           t * n * (8 + r) + uniq * (2 * r + 8), t * n, unique_rows=uniq, top1_row_share=round(top / n, 4))
    report("cfg4 fwd+bwd", ms_f + ms_b, t * n * (8 + 2 * r) + t * n * (8 + r) + uniq * (2 * r + 8), t * n)
    ms = timeit(lambda: ops.BackwardPlan.build(ids, num_rows=n_rows, hash_mode=N.HASH_IDENTITY, ids_per_table=n,
                                               buf=plan_buf), iters=5)
    report("cfg4 plan alone (key kernel + library radix sort, Zipf keys)", ms, t * n * 8 * 2, t * n)


def cfg2_plan():
    """The plan of the headline config alone: 16.4 M slots, 24-bit keys (10 stacked 1M-row tables)."""
    t, n, n_rows = 10, 8192 * 200, 1_000_000
    ids = torch.cat([uniform_ids(n, 1000 + i) for i in range(t)])
    plan_buf = torch.empty(int(N.load().recemb_bwd_plan_bytes(t * n, t * n_rows)), dtype=torch.uint8, device=DEV)
    ms = timeit(lambda: ops.BackwardPlan.build(ids, num_rows=n_rows, ids_per_table=n, buf=plan_buf))
    report("cfg2 plan alone (key kernel + library radix sort (cub onesweep); RECEMB_PLAN_SORT=own: hand-written 2 x 12-bit LSD)", ms, t * n * 8 * 2, t * n)


def front_end():
    """The sequence front end at the cfg 2 shape (B 8192 x L 200, D 64 fp32): QueryTower's trim decided on the
    device + windowed lookup, and QueryTower's input sum as one kernel vs the chain of torch ops it replaces."""
    import recommendations_b200 as R
    b, l, dim, n_rows = 8192, 200, 64, 1_000_000
    g = torch.Generator(device=DEV).manual_seed(77)
    ids = torch.randint(1, 2 ** 62, (b, l), generator=g, device=DEV, dtype=torch.int64)
    valid = torch.randint(1, 161, (b,), generator=g, device=DEV)          # histories of 1..160 events, right-padded
    valid[0] = 160
    ids[torch.arange(l, device=DEV).unsqueeze(0) >= valid.unsqueeze(1)] = 0
    w = torch.randn(n_rows, dim, device=DEV)
    r = dim * 4
    ms_win = timeit(lambda: R.SequenceWindow.from_ids(ids, export_span=8))
    report("front end: sequence window (trim rule of query_tower.py:73-79 on ids == 0), [8192, 200] ids", ms_win, b * l * 8)
    win = R.SequenceWindow.from_ids(ids, export_span=8)
    keep = win.keep
    out = torch.empty(b * l, dim, device=DEV)
    ms_full = timeit(lambda: ops.gather_fwd(w, ids, out=out, zero_pad=True, pad_id=0, flip_len=l))
    ms_w = timeit(lambda: ops.gather_fwd(w, ids, out=out, zero_pad=True, pad_id=0, flip_len=l, window=win))
    report(f"front end: windowed gather + flip + pad mask, keep {keep} of {l} columns (full lookup: {ms_full:.4f} ms)",
           ms_w, b * l * 8 + b * keep * 2 * r, b * keep, ms_full_lookup=round(ms_full, 4), keep=keep)
    mods = (R.FlatEmbedding(4, dim, device=DEV), R.PatternFromTimelocal(3600, 24, dim, device=DEV),
            R.PatternFromTimelocal(3600, 168, dim, device=DEV), R.PatternFromTimelocal(86400, 7, dim, device=DEV))
    base = torch.randn(b, l, dim, device=DEV)
    labels = torch.randint(0, 4, (b, l), generator=g, device=DEV)
    ts = torch.randint(1_600_000_000, 1_700_000_000, (b, l), generator=g, device=DEV)
    mask = ids == 0
    pad = torch.randn(1, 1, dim, device=DEV)
    with torch.no_grad():
        fused = lambda: R.fused_lookup_sum(base, [(mods[0], labels), (mods[1], ts), (mods[2], ts), (mods[3], ts)],
                                           mask=mask, masked_row=pad)
        chain = lambda: torch.where(mask.unsqueeze(-1), pad.expand(b, l, -1),
                                    base + mods[0](labels) + mods[1](ts) + mods[2](ts) + mods[3](ts))
        assert torch.equal(fused(), chain())
        ms_f, ms_c = timeit(fused), timeit(chain)
    unmasked = int((~mask).sum())
    # one write per position, one read of the dense term per UNMASKED position, labels + timestamps + mask
    report("front end: QueryTower input sum (4 tiny-table lookups + dense term + pad select) as one kernel", ms_f,
           b * l * (r + 16 + 1) + unmasked * r, unmasked * 4, ms_chain_of_lookups_and_adds=round(ms_c, 4),
           unmasked_share=round(unmasked / (b * l), 3))


RUNNERS = {"kshift": kshift_series, "cfg3": cfg3, "cfg4": cfg4, "cfg2plan": cfg2_plan, "frontend": front_end}


def run(which, quiet=True):
    """Runs the named measurements and returns their result dicts (used by bench.py)."""
    global QUIET
    QUIET = quiet
    del RESULTS[:]
    for name in which:
        RUNNERS[name]()
        torch.cuda.empty_cache()
    return list(RESULTS)


if __name__ == "__main__":
    run(sys.argv[1:] or ["cfg2plan", "kshift", "cfg3", "cfg4"], quiet=False)
