"""TorchScript export through the TORCH_LIBRARY operators (recommendations_b200/export.py,
csrc_torch/torch_ops.cpp): embedding_module_gen.py:186-196 scripts ModelWrapper(model, mask_model) and
saves it; models/lthm/sequence/encoder.py:25-29 loads it back."""
import io

import pytest
import torch

import recommendations_b200 as R
import recommendations_b200.export as X
from oracle import embedding_oracle as O
from conftest import seeded_ids
from tolerances import mask_mlp


def test_operator_library_loads_and_modules_script_without_a_gpu():
    X.load_ops()
    for name in ("kshift_fwd", "gather_fwd", "pool_fwd"):
        assert hasattr(torch.ops.recemb_b200, name)
    m = X.ScriptableKShiftEmbedding(torch.randn(100, 32), 8, True)
    sm = torch.jit.script(m)
    buf = io.BytesIO()
    torch.jit.save(sm, buf)
    buf.seek(0)
    lm = torch.jit.load(buf)
    assert list(lm.state_dict()) == ["emb.weight"]                      # the reference's checkpoint key
    assert "recemb_b200.kshift_fwd" in lm.code
    with pytest.raises(Exception, match="CUDA|no CPU fallback|recemb_b200"):  # no CPU fallback
        lm(torch.zeros(3, dtype=torch.int64))
    f = torch.jit.script(X.ScriptableFlatEmbedding(torch.randn(10, 8), False))
    assert list(f.state_dict()) == ["_emb_table.weight"]


@pytest.mark.gpu
def test_scripted_model_wrapper_roundtrip_equals_eager_and_oracle():
    dev = "cuda:0"
    torch.manual_seed(3)
    ids = seeded_ids(4 * 33, 8, (4, 33)).to(dev)
    model = R.KShiftEmbedding(1009, 32, num_shifts=16, normalize_output=True, device=dev)
    mask = torch.nn.Sequential(R.KShiftEmbedding(1009, 4, num_shifts=16, normalize_output=False, device=dev),
                               mask_mlp(4).to(dev))
    wrapper = X.ModelWrapper(model, mask)                                # embedding_module_gen.py:186-190
    scripted = torch.jit.script(wrapper)
    buf = io.BytesIO()
    torch.jit.save(scripted, buf)
    buf.seek(0)
    loaded = torch.jit.load(buf, map_location=dev)                       # encoder.py:25-29
    with torch.no_grad():
        eager = model(ids) * mask(ids).sigmoid()
        got = loaded(ids)
    assert torch.equal(got, eager)
    want = O.kshift_embedding(model.emb.weight.detach().cpu(), ids.cpu(), 16, True) * \
        mask[1](O.kshift_embedding(mask[0].emb.weight.detach().cpu(), ids.cpu(), 16, False).to(dev)).sigmoid().cpu()
    torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-6)
    assert sorted(loaded.state_dict())[:2] == ["mask_model.0.emb.weight", "mask_model.1.model.0.bias"]
    # the scriptable twins share the trained tensors: an update of the training module shows up in the export
    model.emb.weight.data.mul_(2.0)
    assert torch.equal(wrapper.model.emb.weight, model.emb.weight)


@pytest.mark.gpu
def test_scripted_flat_and_pooled_ops():
    dev = "cuda:0"
    ids = seeded_ids(6 * 20, 9, (6, 20)).to(dev)
    ids[:, 15:] = 0
    fe = R.FlatEmbedding(500, 64, normalize_output=True, fused_pad_mask=True, flip_sequences=True, device=dev)
    sf = torch.jit.script(X.scriptable(fe))
    with torch.no_grad():
        assert torch.equal(sf(ids), fe(ids))
    pb = R.PooledEmbeddingBag(500, 128, mode="mean", dtype=torch.bfloat16, device=dev)
    sp = torch.jit.script(X.scriptable(pb))
    with torch.no_grad():
        assert torch.equal(sp(ids), pb(ids))
