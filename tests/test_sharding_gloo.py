"""Host logic of the row-wise sharded bag on CPU with gloo, world_size 2: collectives, owner
mapping, partial-sum reduction, gradient routing.  The three compute hooks are replaced by
oracle implementations HERE ONLY (the product hooks are the CUDA kernels); the test proves
sharded == unsharded."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import embedding_oracle as O

N_ROWS, DIM, B, P = 1001, 16, 37, 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _window(ids, lengths, last_n):
    m, p = ids.shape
    pos = torch.arange(p).unsqueeze(0)
    hi = torch.full((m, 1), p) if lengths is None else lengths.long().clamp(0, p).unsqueeze(1)
    lo = (hi - last_n).clamp(min=0) if last_n > 0 else torch.zeros_like(hi)
    return (pos >= lo) & (pos < hi)


def _make_hooks(world, rank, last_n):
    def local_pool(shard, ids_all, len_all):
        rows = O.row_index(ids_all, N_ROWS, 0)
        use = _window(ids_all, len_all, last_n) & (rows % world == rank)
        loc = torch.div(rows, world, rounding_mode="floor")
        out = torch.zeros(ids_all.shape[0], shard.shape[1])
        for j in range(ids_all.shape[1]):
            out = torch.where(use[:, j:j + 1], out + shard[loc[:, j].clamp(max=shard.shape[0] - 1)], out)
        return out

    def local_backward(ids_all, len_all, g_all):
        rows = O.row_index(ids_all, N_ROWS, 0)
        use = _window(ids_all, len_all, last_n) & (rows % world == rank)
        loc = torch.div(rows, world, rounding_mode="floor")
        n_local = (N_ROWS - rank + world - 1) // world
        gw = torch.zeros(n_local, g_all.shape[1])
        g = g_all.unsqueeze(1).expand(-1, ids_all.shape[1], -1)
        gw.index_add_(0, loc[use], g[use])
        return gw

    def reduce(recv, scale):
        out = recv.sum(0)
        return out if scale is None else out * scale.unsqueeze(1)

    # routed exchange hooks (oracle restatements of csrc/route.cu)
    def bucket(ids, lengths):
        rows = O.row_index(ids, N_ROWS, 0)
        owner, loc = rows % world, torch.div(rows, world, rounding_mode="floor")
        use = _window(ids, lengths, last_n)
        bag = torch.arange(ids.shape[0]).unsqueeze(1).expand_as(ids)
        ent, cnt = [], []
        for o in range(world):
            m = use & (owner == o)  # boolean indexing keeps slot order: stable
            ent.append((loc[m] << 32) | (rank * ids.shape[0] + bag[m]))
            cnt.append(int(m.sum()))
        return torch.cat(ent), torch.tensor(cnt, dtype=torch.int64)

    def pool_entries(shard, recv, out_rows):
        out = torch.zeros(out_rows, shard.shape[1])
        out.index_add_(0, recv & 0xFFFFFFFF, shard[recv >> 32])
        return out

    def entries_backward(recv, g_all, wait=lambda: None):
        wait()
        n_local = (N_ROWS - rank + world - 1) // world
        gw = torch.zeros(n_local, g_all.shape[1])
        gw.index_add_(0, recv >> 32, g_all[recv & 0xFFFFFFFF])
        return gw

    return dict(local_pool=local_pool, local_backward=local_backward, reduce_partials=reduce, bucket=bucket,
                pool_entries=pool_entries, entries_backward=entries_backward)


def _worker(rank, world, port, mode, last_n, exchange, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from recommendations_b200.sharded import RowWiseShardedEmbeddingBag
        torch.manual_seed(0)
        full = torch.randn(N_ROWS, DIM)
        g = torch.Generator().manual_seed(100 + rank)
        ids = torch.randint(-2 ** 63, 2 ** 63 - 1, (B, P), generator=g, dtype=torch.int64)
        lengths = torch.randint(0, P + 1, (B,), generator=g)
        go = torch.randn(B, DIM, generator=g)
        mod = RowWiseShardedEmbeddingBag(N_ROWS, DIM, mode=mode, last_n=last_n, exchange=exchange,
                                         **_make_hooks(world, rank, last_n))
        assert mod.emb.weight.shape[0] == (N_ROWS - rank + world - 1) // world
        mod.load_full_weight(full)
        out = mod(ids, lengths)
        want = O.pooled_bag(full, ids, lengths=lengths, last_n=last_n, mode=mode)
        torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)
        out.backward(go)
        # unsharded gradient of the GLOBAL batch (all ranks' bags), then this rank's rows
        all_ids = [torch.empty_like(ids) for _ in range(world)]
        all_len = [torch.empty_like(lengths) for _ in range(world)]
        all_go = [torch.empty_like(go) for _ in range(world)]
        dist.all_gather(all_ids, ids)
        dist.all_gather(all_len, lengths)
        dist.all_gather(all_go, go)
        wr = full.clone().requires_grad_(True)
        for i, l, gg in zip(all_ids, all_len, all_go):
            rows = O.row_index(i, N_ROWS, 0)
            use = _window(i, l, last_n).float().unsqueeze(-1)
            pooled = (wr[rows] * use).sum(1)
            if mode == "mean":
                pooled = pooled / use.sum(1).clamp(min=1)
            pooled.backward(gg)
        torch.testing.assert_close(mod.emb.weight.grad, wr.grad[rank::world], rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(mod.gather_full_weight(), full)
        # checkpoint of the fused-optimizer state: global layout out, owned rows back in
        from recommendations_b200.table import FusedOptimizerConfig
        for kind, shape in (("rowwise_adagrad", (N_ROWS,)), ("adam", (N_ROWS, DIM))):
            fm = RowWiseShardedEmbeddingBag(N_ROWS, DIM, exchange=exchange, fused_optimizer=FusedOptimizerConfig(kind=kind),
                                            **_make_hooks(world, rank, last_n))
            glob = torch.arange(N_ROWS, dtype=torch.float32)
            glob = glob if len(shape) == 1 else glob.unsqueeze(1).expand(*shape).contiguous()
            state = {"kind": kind, "step": 5, "state1": glob, "state2": glob * 2 if kind == "adam" else None}
            fm.load_full_optimizer_state(state)
            assert fm.emb.fused_step == 5
            assert torch.equal(fm.emb._buffers["opt_state1"].reshape(-1)[:3],
                               glob[rank::world].reshape(-1)[:3])
            back = fm.gather_full_optimizer_state()
            assert back["step"] == 5 and torch.equal(back["state1"], glob)
            if kind == "adam":
                assert torch.equal(back["state2"], glob * 2)
        result[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,last_n,exchange", [("sum", 0, "route"), ("mean", 0, "route"), ("sum", 2, "route"),
                                                  ("sum", 0, "gather"), ("mean", 2, "gather")])
def test_sharded_equals_unsharded_gloo_world2(mode, last_n, exchange):
    world = 2
    result = mp.Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), mode, last_n, exchange, result), nprocs=world, join=True)
    assert dict(result) == {0: 1, 1: 1}


def test_local_rows_partition_is_exact():
    from recommendations_b200.sharded import local_rows_of
    for n in (1, 7, 8, 9, 1000, 200_000_000):
        for w in (1, 2, 4, 8):
            assert sum(local_rows_of(n, w, r) for r in range(w)) == n


def _tablewise_worker(rank, world, port, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from recommendations_b200.sharded import TableWiseShardedEmbeddingBag
        t, n, d = 5, 37, 8
        full = torch.arange(t * n * d, dtype=torch.float32).view(t, n, d)
        mod = TableWiseShardedEmbeddingBag(n, d, t)
        assert mod.local_tables == (t - rank + world - 1) // world
        assert mod.emb.weight.shape == (mod.local_tables * n, d)
        mod.load_full_weight(full)
        # rank r holds tables r, r + W, ... whole, stacked in that order
        assert torch.equal(mod.emb.weight.detach().view(mod.local_tables, n, d), full[rank::world])
        assert torch.equal(mod.gather_full_weight(), full)
        assert mod._owner_of_table.tolist() == [i % world for i in range(t)]
        # fused-optimizer state: global layout out, owned tables back in
        from recommendations_b200.table import FusedOptimizerConfig
        for kind, shape in (("rowwise_adagrad", (t, n)), ("adam", (t, n, d))):
            fm = TableWiseShardedEmbeddingBag(n, d, t, fused_optimizer=FusedOptimizerConfig(kind=kind))
            glob = torch.arange(t * n, dtype=torch.float32).view(t, n)
            glob = glob if len(shape) == 2 else glob.unsqueeze(2).expand(*shape).contiguous()
            fm.load_full_optimizer_state({"kind": kind, "step": 7, "state1": glob,
                                          "state2": glob * 2 if kind == "adam" else None})
            assert fm.emb.fused_step == 7
            assert torch.equal(fm.emb._buffers["opt_state1"].reshape((fm.local_tables,) + shape[1:]), glob[rank::world])
            back = fm.gather_full_optimizer_state()
            assert back["step"] == 7 and torch.equal(back["state1"], glob)
            if kind == "adam":
                assert torch.equal(back["state2"], glob * 2)
        result[rank] = 1
    finally:
        dist.destroy_process_group()


def test_tablewise_partition_host_logic_gloo_world2():
    """Table-wise partitioning, host side on CPU / gloo: which tables a rank holds, the checkpoint round trip
    global [T, N, D] -> owned tables -> global (the exchange itself is peer memory: tests/test_gpu_peer*.py)."""
    world = 2
    result = mp.Manager().dict()
    mp.spawn(_tablewise_worker, args=(world, _free_port(), result), nprocs=world, join=True)
    assert dict(result) == {0: 1, 1: 1}
    with pytest.raises(ValueError):
        from recommendations_b200.sharded import TableWiseShardedEmbeddingBag, SingleProcess

        class _Four(SingleProcess):
            def __init__(self):
                super().__init__()
                self.world = 4
        TableWiseShardedEmbeddingBag(10, 8, 3, comm=_Four())
