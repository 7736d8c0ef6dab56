"""Module-level semantics around the kernels (GPU): gradient modes agree with each other, a table
looked up more than once per optimizer step gets ONE update with the summed gradient, inference
forwards build no backward plan, identity ids are range-checked."""
import pytest
import torch

import recommendations_b200 as R
from recommendations_b200 import _native as N
from recommendations_b200 import ops
from oracle import embedding_oracle as O
from conftest import seeded_ids

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("flip", [False, True])
def test_sparse_coo_with_fused_pad_mask_equals_dense_grad(flip):
    """sparse=True + fused_pad_mask=True: the COO gradient must skip id == 0 positions exactly like
    the plan-based dense mode (the forward never read the table there)."""
    torch.manual_seed(3)
    w = torch.randn(100, 32)
    ids = seeded_ids(6 * 40, 35, (6, 40))
    ids[:, 25:] = 0
    ids[2, 3] = 0
    go = torch.randn(6, 40, 32)
    grads = []
    for sparse in (False, True):
        m = R.FlatEmbedding(100, 32, fused_pad_mask=True, sparse=sparse, flip_sequences=flip, device=DEV)
        m.load_state_dict({"_emb_table.weight": w})
        m(ids.to(DEV)).backward(go.to(DEV))
        g = m._emb_table.weight.grad
        assert g.is_sparse == sparse
        grads.append((g.to_dense() if sparse else g).cpu())
    torch.testing.assert_close(grads[1], grads[0], rtol=1e-6, atol=1e-6)
    # and both equal the oracle's dense gradient over the non-pad positions
    go_eff = go.flip(1) if flip else go
    keep = (ids != 0).reshape(-1)
    want = O.dense_grad(O.row_index(ids, 100, 0).reshape(-1)[keep], go_eff.reshape(-1, 32)[keep], 100)
    torch.testing.assert_close(grads[0], want, rtol=1e-5, atol=1e-5)
    assert grads[0][0].abs().sum() == want[0].abs().sum()  # row floor_mod(0, N) gets no pad gradient


@pytest.mark.parametrize("kind", ["adagrad", "adam"])
def test_shared_table_two_lookups_one_update(kind):
    """accumulate=True: history + target lookups of one shared table inside one step == torch.optim
    on the summed gradient (one update, one step count), not two sequential updates."""
    n_rows, dim = 211, 16
    torch.manual_seed(5)
    w0 = torch.randn(n_rows, dim)
    ids_a, ids_b = seeded_ids(300, 61), seeded_ids(120, 62)
    ids_b[:40] = ids_a[:40]  # overlapping rows: (g1 + g2)^2 != g1^2 + g2^2
    cfg = R.FusedOptimizerConfig(kind=kind, lr=0.05, eps=1e-8, accumulate=True,
                                 initial_accumulator_value=0.1 if kind == "adagrad" else 0.0)
    m = R.FlatEmbedding(n_rows, dim, device=DEV, fused_optimizer=cfg)
    m.load_state_dict({"_emb_table.weight": w0})
    opt = R.FusedEmbeddingOptimizer([m._emb_table])
    ref = torch.nn.Embedding.from_pretrained(w0.clone(), freeze=False)
    ropt = torch.optim.Adagrad(ref.parameters(), lr=0.05, eps=1e-8, initial_accumulator_value=0.1) \
        if kind == "adagrad" else torch.optim.Adam(ref.parameters(), lr=0.05, eps=1e-8)
    for step in range(3):
        ga = torch.randn(300, dim, generator=torch.Generator().manual_seed(10 + step))
        gb = torch.randn(120, dim, generator=torch.Generator().manual_seed(20 + step))
        opt.zero_grad()
        (m(ids_a.to(DEV)) * ga.to(DEV)).sum().backward()
        (m(ids_b.to(DEV)) * gb.to(DEV)).sum().backward()
        opt.step()
        ropt.zero_grad()
        ((ref(torch.remainder(ids_a, n_rows)) * ga).sum() + (ref(torch.remainder(ids_b, n_rows)) * gb).sum()).backward()
        ropt.step()
    assert m._emb_table.fused_step == 3
    rows = torch.cat([torch.remainder(ids_a, n_rows), torch.remainder(ids_b, n_rows)]).unique()
    got = m._emb_table.weight.cpu()
    # lazy Adam only moves touched rows (SparseAdam semantics); every row here that was touched in all
    # three steps matches dense Adam, untouched rows stay put
    torch.testing.assert_close(got[rows], ref.weight.detach()[rows], rtol=1e-5, atol=1e-5)


def test_second_backward_in_one_step_is_refused_without_accumulate():
    m = R.FlatEmbedding(50, 8, device=DEV, fused_optimizer=R.FusedOptimizerConfig(kind="adagrad", lr=0.1))
    opt = R.FusedEmbeddingOptimizer([m._emb_table])
    ids = seeded_ids(20, 1).to(DEV)
    m(ids).sum().backward()
    with pytest.raises(RuntimeError, match="accumulate=True"):
        m(ids).sum().backward()
    opt.step()
    assert m._emb_table.fused_step == 1
    m(ids).sum().backward()  # next step: fine again
    opt.step()
    assert m._emb_table.fused_step == 2
    # without a facade every backward is its own optimizer step
    m2 = R.FlatEmbedding(50, 8, device=DEV, fused_optimizer=R.FusedOptimizerConfig(kind="adagrad", lr=0.1))
    m2(ids).sum().backward()
    m2(ids).sum().backward()
    assert m2._emb_table.fused_step == 2


def test_inference_forward_builds_no_plan():
    """Under torch.no_grad() a lookup is exactly one kernel: no side-stream hash / sort, no plan buffer."""
    ids = seeded_ids(4096, 2, (64, 64)).to(DEV)
    mods = [R.FlatEmbedding(1000, 32, device=DEV, normalize_output=True),
            R.KShiftEmbedding(1000, 32, num_shifts=4, device=DEV),
            R.PooledEmbeddingBag(1000, 32, device=DEV,
                                 fused_optimizer=R.FusedOptimizerConfig(kind="sgd", lr=0.1))]
    for m in mods:
        m(ids)
        torch.cuda.synchronize()
        c0 = N.launch_count()
        with torch.no_grad():
            m(ids)
        assert N.launch_count() - c0 == 1, type(m).__name__
        c0 = N.launch_count()
        m(ids)  # training forward: gather + plan (keys + sort)
        assert N.launch_count() - c0 > 1, type(m).__name__


def test_identity_ids_out_of_range():
    """hash_ids=False: the module raises like nn.EmbeddingBag; the raw kernels drop the slot in
    forward AND in the plan (no out-of-bounds read, no update of an aliased row)."""
    w = torch.randn(40, 16)
    ids = torch.tensor([[0, 39, 5], [40, 7, -1], [2 ** 40, 1, 1]], dtype=torch.int64)
    m = R.PooledEmbeddingBag(40, 16, hash_ids=False, device=DEV)
    m.load_state_dict({"emb.weight": w})
    with pytest.raises(IndexError):
        m(ids.to(DEV))
    got = ops.pool_fwd(w.to(DEV), ids.to(DEV), hash_mode=N.HASH_IDENTITY).cpu()
    ok = (ids >= 0) & (ids < 40)
    want = torch.stack([sum((w[i] for i, k in zip(r.tolist(), o.tolist()) if k), torch.zeros(16))
                        for r, o in zip(ids, ok)])
    torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
    plan = ops.BackwardPlan.build(ids.to(DEV), num_rows=40, hash_mode=N.HASH_IDENTITY, bag_size=3)
    n_valid, n_unique = plan.counters.cpu().tolist()
    assert n_valid == int(ok.sum()) and n_unique == ids[ok].unique().numel()
    out, _ = ops.gather_fwd(w.to(DEV), ids.to(DEV), hash_mode=N.HASH_IDENTITY)
    want_g = torch.where(ok.unsqueeze(-1), w[ids.clamp(0, 39)], torch.zeros(()))
    assert torch.equal(out.cpu(), want_g)
