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
