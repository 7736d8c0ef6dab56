"""Row-wise sharded pooled lookup on ONE GPU with the W owners emulated in sequence: the owner-side
kernels (shard-filtered pool / plan / update) and the requester-side reduction reproduce the
unsharded result."""
import pytest
import torch

from recommendations_b200 import _native as N
from recommendations_b200 import ops
from oracle import embedding_oracle as O
from conftest import seeded_ids

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_sharded_pool_equals_unsharded(world, dtype):
    n_rows, dim, m, p = 100003, 128, 4001, 20
    torch.manual_seed(world)
    full = torch.randn(n_rows, dim).to(dtype)
    ids = seeded_ids(m * p, 70, (m, p))
    lengths = torch.randint(0, p + 1, (m,), generator=torch.Generator().manual_seed(1))
    parts = []
    for r in range(world):
        shard = full[r::world].contiguous().to(DEV)
        parts.append(ops.pool_fwd(shard, ids.to(DEV), lengths=lengths.to(DEV), num_rows=n_rows,
                                  shard_world=world, shard_rank=r))
    out = ops.sum_partials(torch.stack(parts))
    want = O.pooled_bag(full, ids, lengths=lengths)
    bf = dtype == torch.bfloat16
    # bf16: every owner rounds its partial to bf16 before the exchange (half the NVLink bytes of
    # fp32 partials), so the sum carries up to W half-ulps of values of magnitude ~8 (ulp 0.06)
    torch.testing.assert_close(out.float().cpu(), want.float(), rtol=1e-2 if bf else 1e-5,
                               atol=8e-2 if bf else 1e-5)
    # every slot is pooled by exactly one owner
    ones = torch.ones(n_rows, 16)
    cnt = sum(ops.pool_fwd(ones[r::world].contiguous().to(DEV), ids.to(DEV), lengths=lengths.to(DEV),
                           num_rows=n_rows, shard_world=world, shard_rank=r)[:, 0] for r in range(world))
    assert torch.equal(cnt.cpu().long(), lengths)


@pytest.mark.parametrize("world", [2, 8])
def test_sharded_update_equals_unsharded(world):
    n_rows, dim, m, p = 50021, 64, 3000, 20
    torch.manual_seed(3)
    full = torch.randn(n_rows, dim)
    ids = seeded_ids(m * p, 71, (m, p))
    go = torch.randn(m, dim, generator=torch.Generator().manual_seed(2))
    hp = ops.make_optim_params(lr=0.5, eps=1e-10)
    # unsharded
    w_ref = full.clone().to(DEV)
    s_ref = torch.zeros(n_rows, device=DEV)
    plan = ops.BackwardPlan.build(ids.to(DEV), num_rows=n_rows, bag_size=p)
    ops.bwd_apply(plan, go.to(DEV), table=w_ref, update=N.UPD_ROWWISE_ADAGRAD, state1=s_ref,
                  slots_per_grad_row=p, hp=hp)
    n_valid_total = 0
    for r in range(world):
        shard = full[r::world].contiguous().to(DEV)
        st = torch.zeros(shard.shape[0], device=DEV)
        plan_r = ops.BackwardPlan.build(ids.to(DEV), num_rows=n_rows, bag_size=p, shard_world=world,
                                        shard_rank=r)
        assert plan_r.num_rows == shard.shape[0]
        n_valid_total += int(plan_r.counters.cpu()[0])
        rows = plan_r.sorted_rows.cpu()
        glob = O.row_index(ids, n_rows, 0).reshape(-1)
        want_rows = torch.sort(glob[glob % world == r] // world).values
        assert torch.equal(rows[rows < shard.shape[0]], want_rows)  # dedup input bit-exact
        ops.bwd_apply(plan_r, go.to(DEV), table=shard, update=N.UPD_ROWWISE_ADAGRAD, state1=st,
                      slots_per_grad_row=p, hp=hp)
        torch.testing.assert_close(shard, w_ref[r::world], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(st, s_ref[r::world], rtol=1e-5, atol=1e-7)
    assert n_valid_total == m * p


@pytest.mark.parametrize("world,tables", [(1, 1), (2, 1), (8, 1), (4, 3)])
def test_routed_exchange_kernels(world, tables):
    """bucket -> (emulated exchange) -> owner run-pool / plan: every sender rank bucketed on one GPU,
    the buckets re-assembled per owner exactly as the all-to-all would."""
    n_rows, dim, b, p = 30011, 64, 301, 20
    m = tables * b
    torch.manual_seed(world)
    full = torch.randn(tables, n_rows, dim)
    per_rank = []
    for r in range(world):
        ids = seeded_ids(m * p, 80 + r, (m, p))
        lengths = torch.randint(0, p + 1, (m,), generator=torch.Generator().manual_seed(r))
        ent, cnt = ops.shard_bucket(ids.to(DEV), num_rows=n_rows, world=world, rank=r, bags_total=m,
                                    lengths=lengths.to(DEV), bags_per_table=b if tables > 1 else 0,
                                    num_tables=tables if tables > 1 else 0)
        cnt = cnt.cpu()
        assert int(cnt.sum()) == int(lengths.sum())
        # oracle bucketing: stable by owner, entry = (local row + table offset) << 32 | global bag
        rows = O.row_index(ids, n_rows, 0)
        owner = rows % world
        t_of_bag = (torch.arange(m) // b).unsqueeze(1).expand(m, p) if tables > 1 else torch.zeros(m, p, dtype=torch.long)
        use = torch.arange(p).unsqueeze(0) < lengths.unsqueeze(1)
        bag = torch.arange(m).unsqueeze(1).expand(m, p)
        want = []
        for o in range(world):
            lr_o = (n_rows - o + world - 1) // world
            sel = use & (owner == o)
            want.append((((rows[sel] // world) + t_of_bag[sel] * lr_o) << 32) | (r * m + bag[sel]))
            assert int(cnt[o]) == int(sel.sum())
        assert torch.equal(ent.cpu()[:int(cnt.sum())], torch.cat(want))  # bit-exact, stable
        per_rank.append((ids, lengths, ent, cnt))
    for o in range(world):
        lr_o = (n_rows - o + world - 1) // world
        shard = full[:, o::world].reshape(-1, dim).contiguous().to(DEV)
        recv = torch.cat([ent[int(cnt[:o].sum()):int(cnt[:o + 1].sum())] for (_, _, ent, cnt) in per_rank])
        part = ops.pool_entries(shard, recv, world * m).cpu()
        for r, (ids, lengths, _, _) in enumerate(per_rank):
            rows = O.row_index(ids, n_rows, 0)
            use = (torch.arange(p).unsqueeze(0) < lengths.unsqueeze(1)) & (rows % world == o)
            tt = (torch.arange(m) // b) if tables > 1 else torch.zeros(m, dtype=torch.long)
            want = torch.zeros(m, dim)
            for j in range(p):
                contrib = full[tt, rows[:, j]]
                want = torch.where(use[:, j:j + 1], want + contrib, want)
            assert torch.equal(part[r * m:(r + 1) * m], want)  # same fp32 order as the unsharded sum
        plan = ops.plan_from_entries(recv, tables * lr_o)
        keys = (recv >> 32).cpu()
        order = torch.argsort(keys, stable=True)
        assert torch.equal(plan.sorted_rows.cpu(), keys[order])
        assert torch.equal(plan.sorted_slots.cpu(), (recv & 0xFFFFFFFF).cpu()[order])


def test_routed_single_process_module_matches_pooled_bag():
    from recommendations_b200.sharded import RowWiseShardedEmbeddingBag
    import recommendations_b200 as R
    n_rows, dim, m, p = 9973, 64, 777, 20
    ids = seeded_ids(m * p, 73, (m, p)).to(DEV)
    lengths = torch.randint(0, p + 1, (m,), generator=torch.Generator().manual_seed(5)).to(DEV)
    a = RowWiseShardedEmbeddingBag(n_rows, dim, mode="sum", exchange="route", device=DEV)
    b = R.PooledEmbeddingBag(n_rows, dim, mode="sum", device=DEV)
    b.load_state_dict(a.state_dict())
    go = torch.randn(m, dim, device=DEV)
    oa, ob = a(ids, lengths), b(ids, lengths)
    assert torch.equal(oa, ob)
    oa.backward(go)
    ob.backward(go)
    torch.testing.assert_close(a.emb.weight.grad, b.emb.weight.grad, rtol=1e-5, atol=1e-6)


def test_single_process_module_matches_pooled_bag():
    from recommendations_b200.sharded import RowWiseShardedEmbeddingBag
    import recommendations_b200 as R
    n_rows, dim, m, p = 9973, 64, 777, 20
    ids = seeded_ids(m * p, 72, (m, p)).to(DEV)
    lengths = torch.randint(0, p + 1, (m,), generator=torch.Generator().manual_seed(4)).to(DEV)
    a = RowWiseShardedEmbeddingBag(n_rows, dim, mode="mean", device=DEV)
    b = R.PooledEmbeddingBag(n_rows, dim, mode="mean", device=DEV)
    b.load_state_dict(a.state_dict())
    go = torch.randn(m, dim, device=DEV)
    oa, ob = a(ids, lengths), b(ids, lengths)
    torch.testing.assert_close(oa, ob, rtol=1e-6, atol=1e-6)
    oa.backward(go)
    ob.backward(go)
    torch.testing.assert_close(a.emb.weight.grad, b.emb.weight.grad, rtol=1e-5, atol=1e-6)
