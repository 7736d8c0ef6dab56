"""Host-side mirror of the reference module surface: constructors,
state_dict keys, parameter visibility in the two gradient modes, loud failure on CPU."""
import pytest
import torch

import recommendations_b200 as R
from recommendations_b200 import _native as N


def test_state_dict_keys_match_reference():
    assert list(R.FlatEmbedding(10, 4).state_dict()) == ["_emb_table.weight"]
    assert list(R.KShiftEmbedding(10, 4).state_dict()) == ["emb.weight"]
    assert list(R.QREmbedding(100, 4, False).state_dict()) == ["emb_q.weight", "emb_r.weight"]
    assert sorted(R.CosineVectorEmbedding(8, 4, n_proj=4, num_bins=6).state_dict()) == \
        ["emb.weight", "grid", "pos_offset", "projection_mat"]


def test_reference_constructor_signatures():
    m = R.FlatEmbedding(10, 4, 0, True, True)
    assert m.padding_idx == 0 and m._normalize_output and float(m._emb_table.weight.abs().sum()) == 0.0
    k = R.KShiftEmbedding(10, 4, 16, True, True)
    assert k._num_shifts == 16 and k._normalize_output and k.emb.sparse and k._num_bits == 64
    q = R.QREmbedding(10007, 4, True)
    assert q._div == 100 and q.num_embeddings == 10000 and q.emb_q.weight.shape == (100, 4)
    c = R.CosineVectorEmbedding(32, 512, n_proj=32, num_bins=12)
    assert c.emb.weight.shape == (13 * 32, 512) and c.pos_offset.tolist()[:3] == [0, 13, 26]


def test_init_matches_nn_embedding_under_same_seed():
    torch.manual_seed(7)
    a = R.FlatEmbedding(50, 8, padding_idx=3)
    torch.manual_seed(7)
    b = torch.nn.Embedding(50, 8, padding_idx=3)
    assert torch.equal(a._emb_table.weight, b.weight)
    assert a._emb_table.weight[3].abs().sum() == 0


def test_load_state_dict_roundtrip_from_reference_shaped_dict():
    src = {"emb.weight": torch.randn(10, 4)}
    k = R.KShiftEmbedding(10, 4)
    k.load_state_dict(src)
    assert torch.equal(k.emb.weight, src["emb.weight"])
    k.emb.enable_fused_optimizer(kind="adagrad", lr=0.5)
    k.load_state_dict(src)  # same key in fused (buffer) mode
    assert list(k.state_dict()) == ["emb.weight"]


def test_fused_mode_hides_table_from_parameters():
    k = R.KShiftEmbedding(10, 4, fused_optimizer=R.FusedOptimizerConfig(kind="rowwise_adagrad"))
    assert list(k.parameters()) == []
    f = R.FlatEmbedding(10, 4)
    assert [n for n, _ in f.named_parameters()] == ["_emb_table.weight"]
    opt = R.FusedEmbeddingOptimizer([f._emb_table], kind="adagrad", lr=0.25)
    assert list(f.parameters()) == [] and f._emb_table.fused.lr == 0.25
    opt.param_groups[0]["lr"] = 0.125  # a scheduler would do this
    opt.step()
    opt.zero_grad()
    assert f._emb_table.fused.lr == 0.125


def test_cpu_tensors_raise_instead_of_falling_back():
    with pytest.raises(N.NativeError, match="no CPU fallback"):
        R.FlatEmbedding(10, 4)(torch.tensor([[1, 2]]))
    with pytest.raises(N.NativeError):
        R.KShiftEmbedding(10, 4).get_row_idx(torch.tensor([1, 2]), 1)


def test_bad_config_is_rejected():
    with pytest.raises(ValueError):
        R.FusedOptimizerConfig(kind="lion")
    with pytest.raises(ValueError):
        R.PooledEmbeddingBag(10, 4, mode="max")


def test_fused_optimizer_checkpoint_round_trip():
    """Optimizer-state save / resume through the standard torch API (CPU tensors: no kernel runs)."""
    import io
    import recommendations_b200 as R
    from recommendations_b200.table import EmbeddingTable, FusedEmbeddingOptimizer

    tabs = [EmbeddingTable(50, 8), EmbeddingTable(30, 8)]
    opt = FusedEmbeddingOptimizer(tabs, kind="adagrad", lr=0.25)
    for i, t in enumerate(tabs):
        t._ensure_state()
        t._buffers["opt_state1"].fill_(i + 1.5)
        t.fused_step = 7 + i
    buf = io.BytesIO()
    torch.save(opt.state_dict(), buf)
    buf.seek(0)
    tabs2 = [EmbeddingTable(50, 8), EmbeddingTable(30, 8)]
    opt2 = FusedEmbeddingOptimizer(tabs2, kind="adagrad", lr=0.5)
    opt2.load_state_dict(torch.load(buf))
    for i, t in enumerate(tabs2):
        assert t.fused_step == 7 + i and t.fused.lr == 0.25
        assert torch.equal(t._buffers["opt_state1"], torch.full((([50, 30][i]), 8), i + 1.5))
    other = FusedEmbeddingOptimizer([EmbeddingTable(50, 8), EmbeddingTable(30, 8)], kind="rowwise_adagrad", lr=0.5)
    with pytest.raises(ValueError, match="optimizer"):
        other.load_state_dict(opt.state_dict())
    # the tables' own state_dict keeps the reference keys only
    assert set(tabs[0].state_dict().keys()) == {"weight"}


def test_embedding_collection_keeps_per_table_checkpoint_keys_and_shared_storage():
    """T tables stacked in one allocation, yet state_dict() shows T independent reference tables."""
    import recommendations_b200 as R
    torch.manual_seed(0)
    c = R.EmbeddingCollection(["user", "item", "ctx"], 50, 8)
    assert sorted(c.state_dict()) == ["tables.ctx._emb_table.weight", "tables.item._emb_table.weight",
                                      "tables.user._emb_table.weight"]
    assert len(list(c.parameters())) == 3
    base = c.table.weight.data_ptr()
    for i, t in enumerate(c.members()):
        assert t.weight.data_ptr() == base + i * 50 * 8 * 4          # views of the stacked storage
    ref = {k: torch.randn(50, 8) for k in c.state_dict()}
    c.load_state_dict(ref)
    assert torch.equal(c.table.weight[50:100], ref["tables.item._emb_table.weight"])
    c = c.double()                                                    # .to() must not un-stack the tables
    base = c.table.weight.data_ptr()
    for i, t in enumerate(c.members()):
        assert t.weight.dtype == torch.float64 and t.weight.data_ptr() == base + i * 50 * 8 * 8
    assert torch.equal(c.table.weight[50:100].float(), ref["tables.item._emb_table.weight"])
    # fused mode: no Parameters, weights stay in state_dict under the same keys, state is per table
    p = R.EmbeddingCollection(2, 30, 4, kind="pooled", fused_optimizer=R.FusedOptimizerConfig(kind="rowwise_adagrad"))
    assert list(p.parameters()) == [] and sorted(p.state_dict()) == ["tables.table_0.emb.weight", "tables.table_1.emb.weight"]
    assert p.members()[1].opt_state1.shape == (30,) and p.table.opt_state1.shape == (60,)
    assert p.members()[1].opt_state1.data_ptr() == p.table.opt_state1.data_ptr() + 30 * 4
    opt = R.FusedEmbeddingOptimizer([p.table])
    assert "fused" in opt.state_dict()
