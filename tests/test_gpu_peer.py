"""Peer exchange of the row-wise sharded pooled lookup, W ranks emulated on ONE GPU: every "peer"
pointer is another allocation on the same device (the kernels cannot tell; only the cross-rank
barrier is left out because the virtual ranks run one after the other).  Forward must be
bit-identical to the unsharded pooled bag, the inboxes bit-identical to the routed buckets, the
owner-side update equal to the unsharded update."""
import pytest
import torch

from recommendations_b200 import _native as N
from recommendations_b200 import ops
from recommendations_b200.peer import PeerGroup, arena_layout
from oracle import embedding_oracle as O
from conftest import seeded_ids

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_groups(world, full, cap, bags_total, dtype):
    """full [T, N, D] -> per-rank stacked shards, arenas and PeerGroups on one device."""
    t, n_rows, dim = full.shape
    shards = [full[:, r::world].reshape(-1, dim).contiguous().to(dtype).to(DEV) for r in range(world)]
    layout = arena_layout(world, cap, bags_total, dim, dtype)
    arenas = [PeerGroup.new_arena(layout, DEV) for _ in range(world)]
    return shards, arenas, [PeerGroup.local(world, r, arenas, shards, layout) for r in range(world)]


@pytest.mark.parametrize("world,tables", [(1, 1), (2, 1), (3, 2), (8, 1), (8, 4)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_peer_pool_is_bit_identical_to_unsharded(world, tables, dtype):
    n_rows, dim, b, p = 100003, 128, 1001, 20
    m = tables * b
    torch.manual_seed(world)
    full = torch.randn(tables, n_rows, dim).to(dtype)
    shards, _, groups = make_groups(world, full.float(), 64, m, dtype)
    stacked = full.reshape(-1, dim).contiguous().to(DEV)
    batching = dict(bags_per_table=b if tables > 1 else 0, num_tables=tables if tables > 1 else 0)
    for r in range(world):
        ids = seeded_ids(m * p, 90 + r, (m, p)).to(DEV)
        lengths = torch.randint(0, p + 1, (m,), generator=torch.Generator().manual_seed(r)).to(DEV)
        for mode in (N.POOL_SUM, N.POOL_MEAN):
            got = ops.peer_pool_fwd(groups[r], ids, num_rows=n_rows, dim=dim, dtype=dtype, lengths=lengths,
                                    pool_mode=mode, **batching)
            want = ops.pool_fwd(stacked, ids, lengths=lengths, num_rows=n_rows, pool_mode=mode, **batching)
            assert torch.equal(got, want)
        # last-N window + pad skipping follow the same code path
        got = ops.peer_pool_fwd(groups[r], ids, num_rows=n_rows, dim=dim, dtype=dtype, lengths=lengths, last_n=5,
                                zero_pad=True, pad_id=int(ids[0, 0]), **batching)
        want = ops.pool_fwd(stacked, ids, lengths=lengths, num_rows=n_rows, last_n=5, zero_pad=True,
                            pad_id=int(ids[0, 0]), **batching)
        assert torch.equal(got, want)
    # against the CPU oracle too (one rank, table 0)
    ids = seeded_ids(b * p, 90, (b, p))
    if tables == 1:
        lengths = torch.randint(0, p + 1, (b,), generator=torch.Generator().manual_seed(0))
        got = ops.peer_pool_fwd(groups[0], ids.to(DEV), num_rows=n_rows, dim=dim, dtype=dtype,
                                lengths=lengths.to(DEV))
        want = O.pooled_bag(full[0], ids, lengths=lengths)
        if dtype == torch.float32:
            assert torch.equal(got.cpu(), want)
        else:
            torch.testing.assert_close(got.float().cpu(), want.float(), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("world,tables", [(1, 1), (2, 1), (8, 1), (4, 3)])
def test_peer_backward_exchange(world, tables):
    """bucket_push / allgather_push of every virtual sender, then every owner's plan + update."""
    n_rows, dim, b, p = 30011, 64, 301, 20
    m = tables * b
    torch.manual_seed(world + 10)
    full = torch.randn(tables, n_rows, dim)
    cap = m * p  # worst case: no overflow
    shards, arenas, groups = make_groups(world, full, cap, m, torch.float32)
    batching = dict(bags_per_table=b if tables > 1 else 0, num_tables=tables if tables > 1 else 0)
    senders = []
    for r in range(world):
        ids = seeded_ids(m * p, 80 + r, (m, p))
        lengths = torch.randint(0, p + 1, (m,), generator=torch.Generator().manual_seed(r))
        go = torch.randn(m, dim, generator=torch.Generator().manual_seed(100 + r))
        ops.peer_bucket_push(groups[r], ids.to(DEV), num_rows=n_rows, lengths=lengths.to(DEV), **batching)
        ops.peer_allgather_push(groups[r], go.to(DEV), int(groups[r].layout.off_grads))
        # the routed bucketing is the (already oracle-checked) reference for the inbox contents
        ent, cnt = ops.shard_bucket(ids.to(DEV), num_rows=n_rows, world=world, rank=r, bags_total=m,
                                    lengths=lengths.to(DEV), **batching)
        senders.append((ids, lengths, go, ent.cpu(), cnt.cpu()))
    torch.cuda.synchronize()
    hp = ops.make_optim_params(lr=0.5, eps=1e-10)
    # unsharded reference: dense gradient of the global batch per table
    gw = torch.zeros(tables, n_rows, dim)
    for (ids, lengths, go, _, _) in senders:
        rows = O.row_index(ids, n_rows, 0)
        use = torch.arange(p).unsqueeze(0) < lengths.unsqueeze(1)
        tt = (torch.arange(m) // b).unsqueeze(1).expand(m, p)
        flat = (tt * n_rows + rows)[use]
        gw.view(-1, dim).index_add_(0, flat, go.unsqueeze(1).expand(-1, p, -1)[use])
    for o in range(world):
        g = groups[o]
        counts = g.counts_view().cpu()
        inbox = g.inbox_view().cpu()
        for r, (_, _, _, ent, cnt) in enumerate(senders):
            assert int(counts[r]) == int(cnt[o])
            lo = int(cnt[:o].sum())
            assert torch.equal(inbox[r, :int(cnt[o])], ent[lo:lo + int(cnt[o])])  # bit-exact, stable
        # gathered gradients: slice r = sender r's rows
        gv = g.grads_view(dim, torch.float32).cpu()
        for r, (_, _, go, _, _) in enumerate(senders):
            assert torch.equal(gv[r * m:(r + 1) * m], go)
        assert int(g.status_word().cpu()[0]) == 0
        # plan: valid pairs sorted by row (stable), sentinel keys last
        total_rows = shards[o].shape[0]
        plan = ops.peer_plan(g, total_rows)
        recv = torch.cat([inbox[r, :int(counts[r])] for r in range(world)])
        keys = recv >> 32
        order = torch.argsort(keys, stable=True)
        n_valid = recv.numel()
        assert torch.equal(plan.sorted_rows.cpu()[:n_valid], keys[order])
        assert torch.equal(plan.sorted_slots.cpu()[:n_valid], (recv & 0xFFFFFFFF)[order])
        assert bool((plan.sorted_rows.cpu()[n_valid:] == total_rows).all())
        # dense gradient of this owner's rows == the unsharded gradient restricted to them
        dense = torch.zeros_like(shards[o])
        ops.bwd_apply(plan, g.grads_view(dim, torch.float32), table=dense, update=N.UPD_DENSE_GRAD,
                      slots_per_grad_row=1)
        want = gw[:, o::world].reshape(-1, dim)
        torch.testing.assert_close(dense.cpu(), want, rtol=1e-5, atol=1e-5)
        # fused row-wise Adagrad on the shard == on the unsharded table
        st = torch.zeros(total_rows, device=DEV)
        w_o = shards[o].clone()
        ops.bwd_apply(plan, g.grads_view(dim, torch.float32), table=w_o, update=N.UPD_ROWWISE_ADAGRAD, state1=st,
                      slots_per_grad_row=1, hp=hp)
        gsq = (want * want).mean(dim=1)
        touched = gsq > 0
        w_want = shards[o].cpu().clone()
        w_want[touched] -= 0.5 * want[touched] / (gsq[touched].sqrt().unsqueeze(1) + 1e-10)
        torch.testing.assert_close(w_o.cpu(), w_want, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("world,tables", [(1, 1), (2, 1), (8, 1), (4, 3)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_peer_push_forward(world, tables, dtype):
    """Push forward: every virtual sender buckets into the owners' inboxes, every owner pools its
    inbox and stores the partial rows into the senders' parts regions, every sender sums them."""
    n_rows, dim, b, p = 30011, 64, 301, 20
    m = tables * b
    torch.manual_seed(world + 20)
    full = torch.randn(tables, n_rows, dim).to(dtype)
    shards, arenas, groups = make_groups(world, full.float(), m * p, m, dtype)
    batching = dict(bags_per_table=b if tables > 1 else 0, num_tables=tables if tables > 1 else 0)
    inputs = []
    for r in range(world):
        ids = seeded_ids(m * p, 60 + r, (m, p))
        lengths = torch.randint(0, p + 1, (m,), generator=torch.Generator().manual_seed(40 + r))
        if r == 1:
            lengths[: m // 2] = 0                # a long gap of bags without entries
        if r == 2:
            lengths[:] = 0                       # a sender that sends nothing at all
        # the owners write EVERY row of parts (zero rows for (bag, owner) pairs without an entry):
        # nothing of this fill may survive
        groups[r].parts_view(dim, dtype).fill_(float("nan"))
        ops.peer_bucket_push(groups[r], ids.to(DEV), num_rows=n_rows, lengths=lengths.to(DEV), **batching)
        inputs.append((ids, lengths))
    for o in range(world):                      # (barrier) owners pool and push
        ops.peer_pool_push(groups[o], dim, dtype)
    stacked = full.reshape(-1, dim).contiguous().to(DEV)
    for r, (ids, lengths) in enumerate(inputs):  # (barrier) requesters sum in owner order
        parts = groups[r].parts_view(dim, dtype)
        got = ops.sum_partials(parts)
        want = ops.pool_fwd(stacked, ids.to(DEV), lengths=lengths.to(DEV), num_rows=n_rows, **batching)
        if world == 1:
            assert torch.equal(got, want)        # one owner: same fp32 order as the unsharded sum
        elif dtype == torch.float32:
            torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
        else:                                    # partials are rounded to bf16 before they travel
            torch.testing.assert_close(got.float(), want.float(), rtol=1e-2, atol=8e-2)
        # each partial equals the oracle's pool restricted to that owner's rows (fp32: exactly)
        if dtype == torch.float32 and tables == 1:
            rows = O.row_index(ids, n_rows, 0)
            use = torch.arange(p).unsqueeze(0) < lengths.unsqueeze(1)
            for o in range(world):
                sel = use & (rows % world == o)
                ref = torch.zeros(m, dim)
                for j in range(p):
                    ref = torch.where(sel[:, j:j + 1], ref + full[0].float()[rows[:, j]], ref)
                assert torch.equal(parts[o].cpu(), ref)


def test_peer_inbox_overflow_sets_status():
    world, n_rows, dim, m, p = 4, 1009, 32, 64, 20
    full = torch.randn(1, n_rows, dim)
    shards, arenas, groups = make_groups(world, full, 16, m, torch.float32)   # far too small
    ids = seeded_ids(m * p, 5, (m, p)).to(DEV)
    ops.peer_bucket_push(groups[1], ids, num_rows=n_rows)
    torch.cuda.synchronize()
    assert int(groups[1].status_word().cpu()[0]) & 1
    for o in range(world):
        assert int(groups[o].counts_view().cpu()[1]) == 16   # clamped to the capacity
    groups[1].snapshot_status()
    torch.cuda.synchronize()
    with pytest.raises(N.NativeError, match="overflow"):
        groups[1].raise_on_status()


@pytest.mark.parametrize("mode", ["sum", "mean"])
@pytest.mark.parametrize("tables", [1, 3])
def test_peer_module_single_rank_matches_other_exchanges(mode, tables):
    """exchange="peer" with W = 1 (own memory as the only peer) through the nn.Module + autograd."""
    from recommendations_b200.sharded import RowWiseShardedEmbeddingBag
    n_rows, dim, b, p = 9973, 64, 257, 20
    shape = (b, p) if tables == 1 else (tables, b, p)
    ids = seeded_ids(tables * b * p, 73, shape).to(DEV)
    lengths = torch.randint(0, p + 1, shape[:-1], generator=torch.Generator().manual_seed(5)).to(DEV)
    a = RowWiseShardedEmbeddingBag(n_rows, dim, mode=mode, exchange="peer", num_tables=tables, device=DEV)
    c = RowWiseShardedEmbeddingBag(n_rows, dim, mode=mode, exchange="gather", num_tables=tables, device=DEV)
    c.load_state_dict(a.state_dict())
    go = torch.randn(*shape[:-1], dim, device=DEV)
    oa, oc = a(ids, lengths), c(ids, lengths)
    torch.testing.assert_close(oa, oc, rtol=1e-6, atol=1e-6)
    oa.backward(go)
    oc.backward(go)
    torch.testing.assert_close(a.emb.weight.grad, c.emb.weight.grad, rtol=1e-5, atol=1e-6)
    # second step reuses the group (same shapes) and the barrier bookkeeping
    a.emb.weight.grad = None
    oa2 = a(ids, lengths)
    assert torch.equal(oa2, oa)
    oa2.backward(go)
    torch.testing.assert_close(a.emb.weight.grad, c.emb.weight.grad, rtol=1e-5, atol=1e-6)
    a.peer_group().raise_on_status(synchronize=True)


@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_peer_module_push_forward_single_rank(mode):
    from recommendations_b200.sharded import RowWiseShardedEmbeddingBag
    n_rows, dim, b, p, tables = 9973, 64, 257, 20, 3
    ids = seeded_ids(tables * b * p, 74, (tables, b, p)).to(DEV)
    lengths = torch.randint(0, p + 1, (tables, b), generator=torch.Generator().manual_seed(6)).to(DEV)
    a = RowWiseShardedEmbeddingBag(n_rows, dim, mode=mode, exchange="peer", peer_forward="push", num_tables=tables,
                                   device=DEV)
    c = RowWiseShardedEmbeddingBag(n_rows, dim, mode=mode, exchange="gather", num_tables=tables, device=DEV)
    c.load_state_dict(a.state_dict())
    go = torch.randn(tables, b, dim, device=DEV)
    for _ in range(2):
        a.emb.weight.grad = None
        c.emb.weight.grad = None
        oa, oc = a(ids, lengths), c(ids, lengths)
        torch.testing.assert_close(oa, oc, rtol=1e-6, atol=1e-6)
        oa.backward(go)
        oc.backward(go)
        torch.testing.assert_close(a.emb.weight.grad, c.emb.weight.grad, rtol=1e-5, atol=1e-6)
    a.peer_group().raise_on_status(synchronize=True)


def test_peer_module_changing_batch_shapes():
    """A second batch shape gets its own arena (cached), the first one keeps working afterwards."""
    from recommendations_b200.sharded import RowWiseShardedEmbeddingBag
    n_rows, dim, p = 9973, 64, 20
    a = RowWiseShardedEmbeddingBag(n_rows, dim, exchange="peer", device=DEV)
    c = RowWiseShardedEmbeddingBag(n_rows, dim, exchange="gather", device=DEV)
    c.load_state_dict(a.state_dict())
    groups = []
    for b in (257, 31, 257, 64, 31):
        ids = seeded_ids(b * p, 200 + b, (b, p)).to(DEV)
        go = torch.randn(b, dim, device=DEV)
        a.emb.weight.grad = None
        c.emb.weight.grad = None
        oa, oc = a(ids), c(ids)
        assert torch.equal(oa, oc)
        oa.backward(go)
        oc.backward(go)
        torch.testing.assert_close(a.emb.weight.grad, c.emb.weight.grad, rtol=1e-5, atol=1e-6)
        groups.append(id(a.peer_group()))
    assert groups[0] == groups[2] and groups[1] == groups[4] and len(set(groups)) == 3
    a.close_peer()
    assert a._peer is None and not a._peer_cache


def test_peer_module_fused_step_in_cuda_graph():
    """The peer step has fixed shapes and no host synchronisation: forward + backward + fused
    row-wise Adagrad replay from one CUDA graph and match the eager steps."""
    from recommendations_b200.sharded import RowWiseShardedEmbeddingBag
    import recommendations_b200 as R
    n_rows, dim, b, p, tables = 50021, 128, 512, 20, 2
    ids = seeded_ids(tables * b * p, 11, (tables, b, p)).to(DEV)
    go = torch.randn(tables, b, dim, device=DEV, dtype=torch.bfloat16)
    cfg = dict(kind="rowwise_adagrad", lr=0.05)
    eager = RowWiseShardedEmbeddingBag(n_rows, dim, exchange="peer", num_tables=tables, dtype=torch.bfloat16,
                                       device=DEV, fused_optimizer=R.FusedOptimizerConfig(**cfg))
    graphed = RowWiseShardedEmbeddingBag(n_rows, dim, exchange="peer", num_tables=tables, dtype=torch.bfloat16,
                                         device=DEV, fused_optimizer=R.FusedOptimizerConfig(**cfg))
    graphed.load_state_dict(eager.state_dict())

    def step(mod):
        out = mod(ids)
        out.backward(go)
        return out

    # warm-up on a side stream (allocations, peer group), then capture one step
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(graphed)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out_static = step(graphed)   # captured, not executed
    for _ in range(3):
        graph.replay()               # graphed: warm-up + 3 replays = 4 steps
    for _ in range(4):
        out_eager = step(eager)      # out_* = the forward of step 4
    torch.cuda.synchronize()
    assert torch.equal(out_static, out_eager)
    assert torch.equal(graphed.emb.weight, eager.emb.weight)
    graphed.peer_group().raise_on_status(synchronize=True)


@pytest.mark.parametrize("tables,groups", [(5, 2), (8, 4), (3, 3)])
@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_peer_pipelined_groups_equal_the_unpipelined_step(tables, groups, mode):
    """Table groups with one arena each, kernels issued by role on three streams (_PeerPipelinedFn): same
    kernels, so the forward on equal tables is bit-identical to the one-arena push step; the fused update
    agrees to fp32 re-association (the segmented reduction cuts its chunks per group), and so do the
    forwards of the following steps."""
    from recommendations_b200.sharded import RowWiseShardedEmbeddingBag
    import recommendations_b200 as R
    n_rows, dim, b, p = 9973, 64, 257, 20
    ids = seeded_ids(tables * b * p, 81, (tables, b, p)).to(DEV)
    lengths = torch.randint(0, p + 1, (tables, b), generator=torch.Generator().manual_seed(8)).to(DEV)
    go = torch.randn(tables, b, dim, device=DEV)
    mods = []
    for g in (groups, 1):
        m = RowWiseShardedEmbeddingBag(n_rows, dim, mode=mode, exchange="peer", peer_forward="push", num_tables=tables,
                                       device=DEV, pipeline_groups=g,
                                       fused_optimizer=R.FusedOptimizerConfig(kind="adagrad", lr=0.1,
                                                                              initial_accumulator_value=0.1))
        if mods:
            m.load_state_dict(mods[0].state_dict())
        mods.append(m)
    assert mods[0]._pipelined() and not mods[1]._pipelined()
    for it in range(3):
        outs = [m(ids, lengths) for m in mods]
        if it == 0:
            assert torch.equal(outs[0], outs[1])
        else:
            torch.testing.assert_close(outs[0], outs[1], rtol=1e-5, atol=1e-5)
        for o in outs:
            o.backward(go)
        torch.testing.assert_close(mods[0].emb.weight, mods[1].emb.weight, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(mods[0].emb.opt_state1, mods[1].emb.opt_state1, rtol=1e-5, atol=1e-5)
    assert mods[0].emb.fused_step == mods[1].emb.fused_step == 3
    assert len(mods[0]._pipe) == groups
    for m in mods:
        m.peer_group().raise_on_status(synchronize=True)
        m.close_peer()


def test_peer_pipelined_step_in_cuda_graph():
    """The pipelined step forks onto two streams and joins again: capturable, replays match eager steps."""
    from recommendations_b200.sharded import RowWiseShardedEmbeddingBag
    import recommendations_b200 as R
    n_rows, dim, b, p, tables = 50021, 128, 512, 20, 4
    ids = seeded_ids(tables * b * p, 12, (tables, b, p)).to(DEV)
    go = torch.randn(tables, b, dim, device=DEV, dtype=torch.bfloat16)
    mk = lambda: RowWiseShardedEmbeddingBag(n_rows, dim, exchange="peer", peer_forward="push", num_tables=tables,  # noqa: E731
                                            dtype=torch.bfloat16, device=DEV, pipeline_groups=2,
                                            fused_optimizer=R.FusedOptimizerConfig(kind="rowwise_adagrad", lr=0.05))
    eager, graphed = mk(), mk()
    graphed.load_state_dict(eager.state_dict())

    def step(mod):
        out = mod(ids)
        out.backward(go)
        return out

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(graphed)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out_static = step(graphed)
    for _ in range(3):
        graph.replay()
    for _ in range(4):
        out_eager = step(eager)
    torch.cuda.synchronize()
    assert torch.equal(out_static, out_eager)
    assert torch.equal(graphed.emb.weight, eager.emb.weight)
    graphed.peer_group().raise_on_status(synchronize=True)


def test_guarded_update_leaves_the_table_untouched():
    """recemb_bwd_apply_guarded: a non-zero guard word (the peer arena's status: overflowed inbox, timed-out
    barrier) skips the whole update -- an incomplete gradient is never applied."""
    w = torch.randn(500, 64, device=DEV)
    state = torch.zeros(500, device=DEV)
    ids = seeded_ids(4000, 13).to(DEV)
    grad = torch.randn(4000, 64, device=DEV)
    plan = ops.BackwardPlan.build(ids, num_rows=500)
    hp = ops.make_optim_params(lr=0.1, eps=1e-10)
    for flag, changed in ((1, False), (0, True)):
        w0, s0 = w.clone(), state.clone()
        guard = torch.full((1,), flag, dtype=torch.int32, device=DEV)
        ops.bwd_apply(plan, grad, table=w, update=N.UPD_ROWWISE_ADAGRAD, state1=state, hp=hp, guard=guard)
        assert torch.equal(w, w0) != changed and torch.equal(state, s0) != changed


# ---------------------------------------------------------------- sequence mode ----
@pytest.mark.parametrize("world,tables", [(1, 1), (2, 1), (4, 3), (8, 2)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_peer_sequence_mode_emulated_ranks(world, tables, dtype):
    """Sequence mode (one output / gradient row per lookup): the forward pulls every row from its owner
    (bit-identical to the unsharded gather); in the backward every sender buckets its lookups into the owners'
    inboxes and stores each gradient row ONCE into the slot its entry names; every owner's plan + dense
    gradient then equals the unsharded scatter-add restricted to the rows it owns."""
    n_rows, dim, b, length = 20011, 64, 37, 50
    n = tables * b * length
    torch.manual_seed(7 + world)
    full = torch.randn(tables, n_rows, dim).to(dtype)
    cap = n                                            # room for everything: no overflow in this test
    shards, arenas, groups = make_groups(world, full.float(), cap, cap, dtype)
    stacked = full.reshape(-1, dim).contiguous().to(DEV)
    batching = dict(ids_per_table=b * length if tables > 1 else 0, num_tables=tables if tables > 1 else 0)
    pool_batching = dict(bags_per_table=b * length if tables > 1 else 0, num_tables=tables if tables > 1 else 0)
    dense_want = [torch.zeros(tables * n_rows, dim) for _ in range(1)][0]
    for r in range(world):
        ids = seeded_ids(n, 300 + r, (n, 1))
        ids[::7] = 0                                   # pad ids: dropped with zero_pad
        got = ops.peer_pool_fwd(groups[r], ids.to(DEV), num_rows=n_rows, dim=dim, dtype=dtype, zero_pad=True,
                                pad_id=0, **pool_batching)
        want, _ = ops.gather_fwd(stacked, ids.view(-1).to(DEV), zero_pad=True, pad_id=0,
                                 ids_per_table=batching["ids_per_table"])
        assert torch.equal(got, want.view(n, dim))
        # backward, sender side
        dest = ops.peer_bucket_push_rows(groups[r], ids.to(DEV), num_rows=n_rows, zero_pad=True, pad_id=0, **batching)
        grad = torch.randn(n, dim, generator=torch.Generator().manual_seed(r)).to(dtype)
        ops.peer_rows_scatter_push(groups[r], grad.to(DEV), dest)
        d = dest.cpu()
        assert bool(((d < 0) == (ids.view(-1) == 0)).all())
        rows = O.row_index(ids.view(-1), n_rows, 0) + (torch.arange(n) // (b * length)) * n_rows
        keep = ids.view(-1) != 0
        assert bool(((d[keep] >> 32) == (rows[keep] % n_rows) % world).all())      # owner = row mod W
        dense_want.index_add_(0, rows[keep], grad.float()[keep])
    for o in range(world):                             # (barrier) owner side
        plan = ops.peer_plan(groups[o], shards[o].shape[0])
        dense = torch.zeros(shards[o].shape, dtype=torch.float32, device=DEV)
        ops.bwd_apply(plan, groups[o].grads_view(dim, dtype).float(), table=dense, update=N.UPD_DENSE_GRAD,
                      slots_per_grad_row=1)
        want_o = dense_want.view(tables, n_rows, dim)[:, o::world].reshape(-1, dim)
        torch.testing.assert_close(dense.cpu(), want_o, rtol=1e-5, atol=1e-4 if dtype == torch.float32 else 5e-2)
        groups[o].snapshot_status()
    torch.cuda.synchronize()
    for g in groups:
        g.raise_on_status()


@pytest.mark.parametrize("tables", [1, 3])
def test_peer_sequence_module_single_rank(tables):
    """RowWiseShardedEmbedding (exchange="peer", W = 1) == FlatEmbedding on the same table, forward bit-exact,
    dense gradient and fused update equal."""
    import recommendations_b200 as R
    from recommendations_b200.sharded import RowWiseShardedEmbedding
    n_rows, dim, b, length = 9973, 64, 33, 40
    shape = (b, length) if tables == 1 else (tables, b, length)
    ids = seeded_ids(tables * b * length, 91, shape).to(DEV)
    a = RowWiseShardedEmbedding(n_rows, dim, exchange="peer", num_tables=tables, device=DEV)
    out = a(ids)
    assert out.shape == shape + (dim,)
    w = a.emb.weight.detach().view(tables, n_rows, dim)
    for t in range(tables):
        ref = R.FlatEmbedding(n_rows, dim, device=DEV)
        ref._emb_table.weight.data.copy_(w[t])
        ids_t = ids if tables == 1 else ids[t]
        out_t = out if tables == 1 else out[t]
        want = ref(ids_t)
        assert torch.equal(out_t, want)
    go = torch.randn_like(out)
    out.backward(go)
    rows = (ids.view(tables, -1) % n_rows) + torch.arange(tables, device=DEV).unsqueeze(1) * n_rows
    want_g = torch.zeros_like(a.emb.weight).index_add_(0, rows.view(-1), go.view(-1, dim))
    torch.testing.assert_close(a.emb.weight.grad, want_g, rtol=1e-5, atol=1e-5)
    a.peer_group().raise_on_status(synchronize=True)
    a.close_peer()


@pytest.mark.parametrize("kind", ["rowwise_adagrad", "sgd"])
@pytest.mark.parametrize("forward", ["pull", "push"])
def test_fused_push_update_equals_push_barrier_update(kind, forward):
    """recemb_peer_bwd_apply_fused (the first CTAs of the level-0 launch push the gradients table by table, the
    reduction groups are gated on per-table arrival counts) == allgather push -> barrier -> guarded update:
    same kernel arithmetic and chunking, so weights and optimizer state are bit-identical over several steps."""
    from recommendations_b200.sharded import RowWiseShardedEmbeddingBag
    import recommendations_b200 as R
    n_rows, dim, b, p, tables = 50021, 128, 300, 20, 5
    ids = seeded_ids(tables * b * p, 31, (tables, b, p)).to(DEV)
    lengths = torch.randint(0, p + 1, (tables, b), generator=torch.Generator().manual_seed(4)).to(DEV)
    go = torch.randn(tables, b, dim, device=DEV, dtype=torch.bfloat16)
    mods = []
    for fused in ("force", False):
        m = RowWiseShardedEmbeddingBag(n_rows, dim, exchange="peer", peer_forward=forward, num_tables=tables,
                                       dtype=torch.bfloat16, device=DEV, pipeline_groups=1,
                                       fused_optimizer=R.FusedOptimizerConfig(kind=kind, lr=0.05))
        m.fused_push = fused
        if mods:
            m.load_state_dict(mods[0].state_dict())
        mods.append(m)
    for _ in range(3):
        outs = [m(ids, lengths) for m in mods]
        assert torch.equal(outs[0], outs[1])
        for o in outs:
            o.backward(go)
    assert mods[0]._fused_push_ok(go.view(-1, dim)) and not mods[1]._fused_push_ok(go.view(-1, dim))
    assert torch.equal(mods[0].emb.weight, mods[1].emb.weight)
    if kind == "rowwise_adagrad":
        assert torch.equal(mods[0].emb.opt_state1, mods[1].emb.opt_state1)
    assert mods[0].emb.fused_step == mods[1].emb.fused_step == 3
    for m in mods:
        m.peer_group().raise_on_status(synchronize=True)
        m.close_peer()


# ------------------------------------------------------------ table-wise partitioning ----
def make_tablewise_groups(world, full, cap, bags_total, dtype):
    """full [T, N, D]: rank r holds tables r, r + W, ... stacked."""
    t, n_rows, dim = full.shape
    shards = [full[r::world].reshape(-1, dim).contiguous().to(dtype).to(DEV) for r in range(world)]
    layout = arena_layout(world, cap, bags_total, dim, dtype)
    arenas = [PeerGroup.new_arena(layout, DEV) for _ in range(world)]
    return shards, arenas, [PeerGroup.local(world, r, arenas, shards, layout) for r in range(world)]


@pytest.mark.parametrize("world,tables", [(1, 3), (2, 5), (4, 8), (8, 8)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_tablewise_exchange_emulated_ranks(world, tables, dtype):
    """Table t lives whole on rank t % W: every sender routes the lookups of table t to that rank, the owner
    pools them from its local table t / W and stores the pooled row (zero rows for its empty bags only) into
    the sender's parts[owner]; the sender's result is parts[t % W][bags of t] -- bit-identical to the unsharded
    bag in fp32.  Backward: each owner's plan over its inbox + all-gathered gradients == the unsharded
    scatter-add restricted to its tables."""
    n_rows, dim, b, p = 20011, 64, 41, 20
    m = tables * b
    torch.manual_seed(3 * world + tables)
    full = torch.randn(tables, n_rows, dim).to(dtype)
    cap = -(-tables // world) * b * p
    shards, arenas, groups = make_tablewise_groups(world, full.float(), cap, m, dtype)
    stacked = full.reshape(-1, dim).contiguous().to(DEV)
    inputs = []
    for r in range(world):
        ids = seeded_ids(m * p, 500 + r, (m, p))
        lengths = torch.randint(0, p + 1, (m,), generator=torch.Generator().manual_seed(60 + r))
        if r == 0:
            lengths[: b] = 0                              # a whole table of empty bags
        groups[r].parts_view(dim, dtype).fill_(float("nan"))
        ops.peer_bucket_push(groups[r], ids.to(DEV), num_rows=n_rows, lengths=lengths.to(DEV), bags_per_table=b,
                             num_tables=tables, tablewise=True)
        inputs.append((ids, lengths))
    for o in range(world):
        ops.peer_pool_push(groups[o], dim, dtype, tablewise_bags_per_table=b)
    owner = torch.arange(tables) % world
    dense_want = torch.zeros(tables * n_rows, dim)
    for r, (ids, lengths) in enumerate(inputs):
        parts = groups[r].parts_view(dim, dtype).view(world, tables, b, dim)
        got = parts[owner.to(DEV), torch.arange(tables, device=DEV)].reshape(m, dim)
        want = ops.pool_fwd(stacked, ids.to(DEV), lengths=lengths.to(DEV), num_rows=n_rows, bags_per_table=b,
                            num_tables=tables)
        assert torch.equal(got, want)                     # one owner per bag: the unsharded arithmetic
        grad = torch.randn(m, dim, generator=torch.Generator().manual_seed(r)).to(dtype)
        ops.peer_allgather_push(groups[r], grad.to(DEV), int(groups[r].layout.off_grads))
        rows = O.row_index(ids, n_rows, 0) + (torch.arange(m) // b).unsqueeze(1) * n_rows
        use = torch.arange(p).unsqueeze(0) < lengths.unsqueeze(1)
        dense_want.index_add_(0, rows[use], grad.float().unsqueeze(1).expand(-1, p, -1)[use])
    for o in range(world):
        plan = ops.peer_plan(groups[o], shards[o].shape[0])
        dense = torch.zeros(shards[o].shape, dtype=torch.float32, device=DEV)
        ops.bwd_apply(plan, groups[o].grads_view(dim, dtype).float(), table=dense, update=N.UPD_DENSE_GRAD,
                      slots_per_grad_row=1)
        want_o = dense_want.view(tables, n_rows, dim)[o::world].reshape(-1, dim)
        torch.testing.assert_close(dense.cpu(), want_o, rtol=1e-5, atol=1e-4 if dtype == torch.float32 else 5e-2)


@pytest.mark.parametrize("mode", ["sum", "mean"])
@pytest.mark.parametrize("fused", [None, "sgd", "rowwise_adagrad"])
def test_tablewise_module_single_rank(mode, fused):
    """TableWiseShardedEmbeddingBag on one rank (owns every table) == the unsharded pooled collection; the fused
    push + update launch (forced on) == push -> barrier -> update."""
    import recommendations_b200 as R
    from recommendations_b200.sharded import TableWiseShardedEmbeddingBag
    n_rows, dim, b, p, tables = 9973, 128, 65, 20, 3
    ids = seeded_ids(tables * b * p, 77, (tables, b, p)).to(DEV)
    lengths = torch.randint(0, p + 1, (tables, b), generator=torch.Generator().manual_seed(9)).to(DEV)
    opt = None if fused is None else R.FusedOptimizerConfig(kind=fused, lr=0.05)
    tw = TableWiseShardedEmbeddingBag(n_rows, dim, tables, mode=mode, dtype=torch.bfloat16, device=DEV, fused_optimizer=opt)
    ref = R.EmbeddingCollection(tables, n_rows, dim, kind="pooled", mode=mode, dtype=torch.bfloat16, device=DEV,
                                fused_optimizer=opt)
    tw.load_full_weight(ref.table.weight.detach().view(tables, n_rows, dim))
    if fused is not None:
        tw.fused_push = "force"
    go = torch.randn(tables, b, dim, device=DEV, dtype=torch.bfloat16)
    for _ in range(2):
        out, want = tw(ids, lengths), ref(ids, lengths)
        torch.testing.assert_close(out.float(), want.float(), rtol=1e-2, atol=1e-2)
        if mode == "sum":
            assert torch.equal(out, want)
        out.backward(go)
        want.backward(go)
        if fused is None:
            gw = torch.cat([m.weight.grad for m in ref.members()])
            torch.testing.assert_close(tw.emb.weight.grad.float(), gw.float(), rtol=1e-2, atol=1e-2)
            tw.emb.weight.grad = None
            for m in ref.members():
                m.weight.grad = None
        else:
            torch.testing.assert_close(tw.emb.weight.float(), ref.table.weight.float(), rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(tw.gather_full_weight().view(-1, dim).float(), tw.emb.weight.float())
    tw.peer_group().raise_on_status(synchronize=True)
    tw.close_peer()
