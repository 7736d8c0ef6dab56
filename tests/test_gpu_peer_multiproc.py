"""The REAL multi-process peer exchange on one GPU: two ranks (two processes, both on cuda:0) export
their shard and arena with CUDA IPC, map each other's memory, and run the pull / push forward, the
entry / gradient pushes and the device-side barrier exactly as on two GPUs (the kernels only see
mapped pointers).  gloo carries the 64-byte handles; NCCL is not involved.  The two contexts
time-slice the GPU, so a barrier costs a context switch instead of microseconds -- fine for parity."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

N_ROWS, DIM, T, B, P = 40009, 64, 3, 129, 20


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _ids(rank):
    g = torch.Generator().manual_seed(900 + rank)
    ids = torch.randint(-2 ** 63, 2 ** 63 - 1, (T, B, P), generator=g, dtype=torch.int64)
    lengths = torch.randint(0, P + 1, (T, B), generator=g)
    go = torch.randn(T, B, DIM, generator=g)
    return ids, lengths, go


def _worker(rank, world, port, peer_forward, result):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import embedding_oracle as O
        from recommendations_b200.sharded import RowWiseShardedEmbeddingBag
        dev = torch.device("cuda:0")
        torch.manual_seed(5)
        full = torch.randn(T, N_ROWS, DIM)
        mod = RowWiseShardedEmbeddingBag(N_ROWS, DIM, num_tables=T, device=dev, exchange="peer",
                                         peer_forward=peer_forward)
        mod.load_full_weight(full)
        ids, lengths, go = _ids(rank)
        gw = torch.zeros(T, N_ROWS, DIM)          # unsharded gradient of the GLOBAL batch
        for r in range(world):
            ids_r, len_r, go_r = _ids(r)
            for t in range(T):
                rows = O.row_index(ids_r[t], N_ROWS, 0)
                use = torch.arange(P).unsqueeze(0) < len_r[t].unsqueeze(1)
                gw[t].index_add_(0, rows[use], go_r[t].unsqueeze(1).expand(-1, P, -1)[use])
        mine = gw[:, rank::world].reshape(-1, DIM)
        for step in range(3):                     # same group, inbox / gradient buffer reused
            mod.emb.weight.grad = None
            out = mod(ids.to(dev), lengths.to(dev))
            for t in range(T):
                want = O.pooled_bag(full[t], ids[t], lengths=lengths[t])
                if peer_forward == "pull":
                    assert torch.equal(out[t].cpu(), want), "pull forward must be bit-identical to unsharded"
                else:
                    torch.testing.assert_close(out[t].cpu(), want, rtol=1e-5, atol=1e-5)
            out.backward(go.to(dev))
            torch.testing.assert_close(mod.emb.weight.grad.cpu(), mine, rtol=1e-4, atol=1e-5)
        mod.peer_group().raise_on_status(synchronize=True)
        # a second batch shape: its own arena, the shard stays mapped once (shared IPC handles)
        first = mod.peer_group()
        out_s = mod(ids[:, :40].contiguous().to(dev), lengths[:, :40].contiguous().to(dev))
        assert mod.peer_group() is not first and mod.peer_group().table_ptrs() == first.table_ptrs()
        torch.testing.assert_close(out_s.cpu(), out[:, :40].cpu(), rtol=1e-5, atol=1e-5)
        out_s.sum().backward()
        mod.peer_group().raise_on_status(synchronize=True)
        mod.close_peer()
        from recommendations_b200 import peer
        assert not peer._OPENED, "every mapped IPC handle must be closed again"
        if peer_forward == "push":
            # the pipelined step (table groups on two streams, one arena per group, fused update guarded by the
            # arena status): SGD with lr 1 leaves w0 - (gradient of the GLOBAL batch) on the rows this rank owns
            import recommendations_b200 as R
            pm = RowWiseShardedEmbeddingBag(N_ROWS, DIM, num_tables=T, device=dev, exchange="peer", peer_forward="push",
                                            pipeline_groups=2, fused_optimizer=R.FusedOptimizerConfig(kind="sgd", lr=1.0))
            pm.load_full_weight(full)
            assert pm._pipelined()
            w0 = pm.emb.weight.detach().cpu().clone()
            out_p = pm(ids.to(dev), lengths.to(dev))
            torch.testing.assert_close(out_p.cpu(), out.cpu(), rtol=1e-6, atol=1e-6)
            out_p.backward(go.to(dev))
            torch.cuda.synchronize()
            torch.testing.assert_close(pm.emb.weight.cpu(), w0 - mine, rtol=1e-4, atol=1e-5)
            pm.peer_group().raise_on_status(synchronize=True)
            pm.close_peer()
            assert not peer._OPENED
        if peer_forward == "pull":
            # sequence mode (one row per lookup): rows pulled from their owners, each gradient row pushed to its
            # owner once; SGD with lr 1 leaves w0 - (scatter-add of the GLOBAL batch) on the rows this rank owns
            import recommendations_b200 as R
            from recommendations_b200.sharded import RowWiseShardedEmbedding
            sm = RowWiseShardedEmbedding(N_ROWS, DIM, num_tables=T, device=dev, exchange="peer",
                                         fused_optimizer=R.FusedOptimizerConfig(kind="sgd", lr=1.0))
            sm.load_full_weight(full)
            w0 = sm.emb.weight.detach().cpu().clone()
            gseq = torch.zeros(T, N_ROWS, DIM)
            for r in range(world):
                ids_r, _, _ = _ids(r)
                go_r = torch.randn(T, B, P, DIM, generator=torch.Generator().manual_seed(70 + r))
                for t in range(T):
                    gseq[t].index_add_(0, O.row_index(ids_r[t].reshape(-1), N_ROWS, 0), go_r[t].reshape(-1, DIM))
            go_seq = torch.randn(T, B, P, DIM, generator=torch.Generator().manual_seed(70 + rank))
            out_q = sm(ids.to(dev))
            for t in range(T):
                assert torch.equal(out_q[t].cpu(), full[t][O.row_index(ids[t], N_ROWS, 0)])
            out_q.backward(go_seq.to(dev))
            torch.cuda.synchronize()
            torch.testing.assert_close(sm.emb.weight.cpu(), w0 - gseq[:, rank::world].reshape(-1, DIM),
                                       rtol=1e-4, atol=1e-5)
            sm.peer_group().raise_on_status(synchronize=True)
            sm.close_peer()
            assert not peer._OPENED
        if peer_forward == "push":
            # table-wise partitioning (table t whole on rank t % W), fused push + update: SGD with lr 1 leaves
            # w0 - (gradient of the GLOBAL batch) on the tables this rank owns
            import recommendations_b200 as R
            from recommendations_b200.sharded import TableWiseShardedEmbeddingBag
            full_bf = torch.randn(T, N_ROWS, 128, generator=torch.Generator().manual_seed(11)).bfloat16()
            tw = TableWiseShardedEmbeddingBag(N_ROWS, 128, T, device=dev, dtype=torch.bfloat16,
                                              fused_optimizer=R.FusedOptimizerConfig(kind="sgd", lr=1.0))
            tw.load_full_weight(full_bf)
            w0 = tw.emb.weight.detach().float().cpu().clone()
            go_tw = torch.randn(T, B, 128, generator=torch.Generator().manual_seed(40 + rank)).bfloat16()
            out_t = tw(ids.to(dev), lengths.to(dev))
            for t in range(T):
                want = O.pooled_bag(full_bf[t], ids[t], lengths=lengths[t])
                assert torch.equal(out_t[t].cpu(), want), "table-wise forward must be bit-identical to unsharded"
            out_t.backward(go_tw.to(dev))
            torch.cuda.synchronize()
            gtw = torch.zeros(T, N_ROWS, 128)
            for r in range(world):
                ids_r, len_r, _ = _ids(r)
                go_r = torch.randn(T, B, 128, generator=torch.Generator().manual_seed(40 + r)).bfloat16().float()
                for t in range(T):
                    rows = O.row_index(ids_r[t], N_ROWS, 0)
                    use = torch.arange(P).unsqueeze(0) < len_r[t].unsqueeze(1)
                    gtw[t].index_add_(0, rows[use], go_r[t].unsqueeze(1).expand(-1, P, -1)[use])
            torch.testing.assert_close(tw.emb.weight.float().cpu(), w0 - gtw[rank::world].reshape(-1, 128),
                                       rtol=2e-2, atol=2e-2)
            full_back = tw.gather_full_weight()
            assert full_back.shape == (T, N_ROWS, 128)
            tw.peer_group().raise_on_status(synchronize=True)
            tw.close_peer()
            assert not peer._OPENED
        result[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("peer_forward", ["pull", "push"])
def test_two_processes_one_gpu_peer_exchange(peer_forward):
    world = 2
    # the manager's server process must be SPAWNED: a fork of this process inherits the CUDA tensors / events
    # of the tests that ran before, and the first garbage collection over there dies in cudaEventDestroy
    # ("CUDA error: initialization error" in a forked child)
    result = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), peer_forward, result), nprocs=world, join=True)
    assert dict(result) == {0: 1, 1: 1}
