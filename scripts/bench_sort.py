#!/usr/bin/env python
"""Times the backward plan (key kernel + hand-written radix sort, csrc/sort.cu) alone at the BASELINE shapes
and checks it against torch.sort(stable=True).  python scripts/bench_sort.py [cfg2 cfg4 cfg3 kshift]"""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "scripts"))
from recommendations_b200 import _native as N  # noqa: E402
from recommendations_b200 import ops  # noqa: E402
import bench_configs as BC  # noqa: E402

DEV = BC.DEV


def check(plan, rows):
    want_rows, order = torch.sort(rows, stable=True)
    assert torch.equal(plan.sorted_rows, want_rows), "keys not sorted"
    assert torch.equal(plan.sorted_slots, order), "not stable in slot"


def run(name):
    if name == "cfg2":
        t, n, n_rows = 10, 8192 * 200, 1_000_000
        ids = torch.cat([BC.uniform_ids(n, 1000 + i) for i in range(t)])
        kw = dict(num_rows=n_rows, ids_per_table=n)
        rows = torch.remainder(ids, n_rows) + (torch.arange(t * n, device=DEV) // n) * n_rows
    elif name == "cfg4":
        t, n, n_rows = 10, 4096 * 1024, 1_000_000
        ids = torch.cat([BC.zipf_rows(n, n_rows, 1.05, 2000 + i) for i in range(t)])
        kw = dict(num_rows=n_rows, ids_per_table=n, hash_mode=N.HASH_IDENTITY)
        rows = ids + (torch.arange(t * n, device=DEV) // n) * n_rows
    elif name == "cfg3":
        f, b, p, n_rows = 26, 16384, 20, 1_000_000
        g = torch.Generator(device=DEV).manual_seed(3000)
        ids = torch.randint(0, n_rows, (f * b, p), generator=g, device=DEV, dtype=torch.int64)
        kw = dict(num_rows=n_rows, ids_per_table=b * p, num_tables=f, hash_mode=N.HASH_IDENTITY, bag_size=p)
        rows = ids.view(-1) + (torch.arange(f * b * p, device=DEV) // (b * p)) * n_rows
    else:  # k-shift: 13.1 M slots, 44 % of them on 127 collapse rows
        n, n_rows, k = 8192 * 200, 1_000_000, 8
        ids = BC.uniform_ids(n, 1000)
        kw = dict(num_rows=n_rows, hash_mode=N.HASH_ROTL_FLOORMOD, slots_per_id=k)
        rows = torch.stack([ops.row_index(ids, N.HASH_ROTL_FLOORMOD, n_rows, c) for c in range(k)], dim=1).view(-1)
    n_slots = rows.numel()
    buf = torch.empty(int(N.load().recemb_bwd_plan_bytes(n_slots, int(rows.max()) + 1 if name != "cfg2" else 10_000_000)) + (64 << 20),
                      dtype=torch.uint8, device=DEV)
    plan = ops.BackwardPlan.build(ids, buf=buf, **kw)
    check(plan, rows)
    c0 = N.launch_count()
    ms = BC.timeit(lambda: ops.BackwardPlan.build(ids, buf=buf, **kw), iters=10)
    print(json.dumps({"name": name, "slots": n_slots, "plan_ms": round(ms, 4), "G_slots_per_s": round(n_slots / ms / 1e6, 2),
                      "launches_per_plan": (N.launch_count() - c0) // 13, "stable_sorted": True}), flush=True)


if __name__ == "__main__":
    for nm in sys.argv[1:] or ["cfg2", "cfg4", "cfg3", "kshift"]:
        run(nm)
        torch.cuda.empty_cache()
