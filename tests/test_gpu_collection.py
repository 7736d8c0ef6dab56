"""EmbeddingCollection (T tables, one launch per phase) against T independent reference-style
FlatEmbedding / PooledEmbeddingBag modules and the oracle (commons/layers.py:44-61; callers
models/lthm/sequence/query_tower.py:24, :53)."""
import pytest
import torch

import recommendations_b200 as R
from recommendations_b200 import _native as N
from oracle import embedding_oracle as O
from conftest import seeded_ids

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _weights(t, n, d, dtype=torch.float32):
    torch.manual_seed(21)
    return [torch.randn(n, d).to(dtype) for _ in range(t)]


@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("flip,pad", [(False, False), (True, True)])
def test_flat_collection_equals_independent_tables(fused, flip, pad):
    t, n, d, b, l = 5, 997, 64, 33, 40
    ws = _weights(t, n, d)
    ids = seeded_ids(t * b * l, 77, (t, b, l))
    ids[:, :, 30:] = 0
    go = torch.randn(t, b, l, d, generator=torch.Generator().manual_seed(5))
    cfg = R.FusedOptimizerConfig(kind="adagrad", lr=0.5, initial_accumulator_value=0.1) if fused else None
    names = [f"f{i}" for i in range(t)]
    coll = R.EmbeddingCollection(names, n, d, device=DEV, fused_pad_mask=pad, flip_sequences=flip, fused_optimizer=cfg)
    coll.load_state_dict({f"tables.{nm}._emb_table.weight": w for nm, w in zip(names, ws)})
    c0 = N.launch_count()
    out = coll(ids.to(DEV))
    fwd_launches = N.launch_count() - c0
    out.backward(go.to(DEV))
    assert out.shape == (t, b, l, d)
    singles = []
    for i in range(t):
        m = R.FlatEmbedding(n, d, device=DEV, fused_pad_mask=pad, flip_sequences=flip,
                            fused_optimizer=None if cfg is None else R.FusedOptimizerConfig(**vars(cfg)))
        m.load_state_dict({"_emb_table.weight": ws[i]})
        o = m(ids[i].to(DEV))
        o.backward(go[i].to(DEV))
        assert torch.equal(out[i], o), i                                  # rows: bit-exact
        want = O.flat_embedding(ws[i], ids[i])
        if pad:
            want = want.masked_fill((ids[i] == 0).unsqueeze(-1), 0.0)
        assert torch.equal(o.cpu(), want.flip(1) if flip else want)
        singles.append(m)
    assert fwd_launches <= 2 + 4 * 3  # one gather + the plan (key kernel + <= 3 sort passes of 4 kernels), whatever T is
    for i, (nm, m) in enumerate(zip(names, singles)):
        tab = coll.members()[i]
        if fused:
            # same terms, other chunk boundaries (stacked vs single plan): fp32 re-association only
            torch.testing.assert_close(tab.weight, m._emb_table.weight, rtol=1e-5, atol=1e-5)
            torch.testing.assert_close(tab.opt_state1, m._emb_table.opt_state1, rtol=1e-5, atol=1e-5)
        else:
            torch.testing.assert_close(tab.weight.grad, m._emb_table.weight.grad, rtol=1e-5, atol=1e-5)
            assert torch.equal(tab.weight, m._emb_table.weight)
    if not fused:  # the reference's optimizer flow over the T Parameters keeps working
        opt = torch.optim.Adagrad(coll.parameters(), lr=0.5)
        opt.step()
        assert not torch.equal(coll.table.weight[:n].cpu(), ws[0])


@pytest.mark.parametrize("mode", ["sum", "mean"])
def test_pooled_collection_equals_independent_bags(mode):
    t, n, d, b, p = 4, 503, 128, 65, 20
    ws = _weights(t, n, d, torch.bfloat16)
    ids = seeded_ids(t * b * p, 78, (t, b, p))
    lengths = torch.randint(0, p + 1, (t, b), generator=torch.Generator().manual_seed(6), dtype=torch.int32)
    go = torch.randn(t, b, d, generator=torch.Generator().manual_seed(7)).bfloat16()
    cfg = R.FusedOptimizerConfig(kind="rowwise_adagrad", lr=0.1, initial_accumulator_value=0.1)
    coll = R.EmbeddingCollection(t, n, d, kind="pooled", mode=mode, dtype=torch.bfloat16, device=DEV, fused_optimizer=cfg)
    coll.load_state_dict({f"tables.table_{i}.emb.weight": w for i, w in enumerate(ws)})
    out = coll(ids.to(DEV), lengths.to(DEV))
    out.backward(go.to(DEV))
    for i in range(t):
        m = R.PooledEmbeddingBag(n, d, mode=mode, dtype=torch.bfloat16, device=DEV,
                                 fused_optimizer=R.FusedOptimizerConfig(**vars(cfg)))
        m.load_state_dict({"emb.weight": ws[i]})
        o = m(ids[i].to(DEV), lengths[i].to(DEV))
        o.backward(go[i].to(DEV))
        assert torch.equal(out[i], o), i
        assert torch.equal(o.cpu(), O.pooled_bag(ws[i], ids[i], lengths=lengths[i], mode=mode))
        torch.testing.assert_close(coll.members()[i].weight.float(), m.emb.weight.float(), rtol=1e-2, atol=1e-2)
        torch.testing.assert_close(coll.members()[i].opt_state1, m.emb.opt_state1, rtol=1e-5, atol=1e-6)


def test_collection_checkpoint_roundtrip_and_device_move():
    c = R.EmbeddingCollection(3, 100, 32, fused_optimizer=R.FusedOptimizerConfig(kind="adagrad", lr=0.5))
    c = c.to(DEV)
    ids = seeded_ids(3 * 50, 79, (3, 50)).to(DEV)
    opt = R.FusedEmbeddingOptimizer([c.table])
    c(ids).sum().backward()
    opt.step()
    sd, osd = c.state_dict(), opt.state_dict()
    c2 = R.EmbeddingCollection(3, 100, 32, device=DEV, fused_optimizer=R.FusedOptimizerConfig(kind="adagrad", lr=0.5))
    c2.load_state_dict(sd)
    opt2 = R.FusedEmbeddingOptimizer([c2.table])
    opt2.load_state_dict(osd)
    assert torch.equal(c2.table.weight, c.table.weight) and torch.equal(c2.table.opt_state1, c.table.opt_state1)
    c(ids).sum().backward()
    c2(ids).sum().backward()
    assert torch.equal(c2.table.weight, c.table.weight)
