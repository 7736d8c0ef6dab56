"""Streaming logQ kernels (csrc/logq.cu) and the DIV_FLOORMOD time-pattern gather vs the
reference-generated fixture and the CPU oracle (commons/layers.py:13-41, :189-237)."""
import numpy as np
import pytest
import torch

import recommendations_b200 as R
from recommendations_b200 import _native as N
from recommendations_b200 import ops
from oracle import embedding_oracle as O
from conftest import seeded_ids

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_streaming_logq_reference_fixture(golden):
    g = golden("streaming_logq")
    ids = torch.from_numpy(g["ids"])
    m = R.CascadedStreamingLogQCorrectionModule(int(g["num_buckets"]), g["offsets"].tolist(), alpha=float(g["alpha"]),
                                                p_init=float(g["p_init"]), device=DEV)
    torch.testing.assert_close(m(ids.to(DEV)).cpu(), torch.from_numpy(g["fwd0"]), rtol=1e-6, atol=1e-6)
    for step in range(4):
        sub = ids[torch.randperm(ids.numel(), generator=torch.Generator().manual_seed(step))[:300]]
        m.train_step(sub.to(DEV), step)
        torch.testing.assert_close(m(ids.to(DEV)).cpu(), torch.from_numpy(g["fwd_steps"][step]), rtol=1e-6, atol=1e-6)
    # the update is two rounded products and one rounded sum per bucket: bit-exact
    assert torch.equal(torch.stack([s.b for s in m.models]).cpu(), torch.from_numpy(g["b"]))
    assert torch.equal(torch.stack([s.a for s in m.models]).cpu(), torch.from_numpy(g["a"]))
    assert sorted(m.state_dict()) == sorted(f"models.{i}.{n}" for i in range(3) for n in "ab")


@pytest.mark.parametrize("num_buckets", [1, 97, 1 << 16, (1 << 24)])
def test_streaming_logq_random_vs_oracle(num_buckets):
    offsets = [0, 34144, 7465477, 64363466, 4234551, 245435435, 143244556]  # hydra-configs/model/lthm.yaml:8
    n = 20000
    ids = seeded_ids(n, 71)
    ids[:50] = ids[50:100]                     # duplicates inside one update
    ids[100] = 2 ** 63 - 1                     # id + offset wraps
    ids[101] = -2 ** 63
    m = R.CascadedStreamingLogQCorrectionModule(num_buckets, offsets, alpha=0.05, p_init=0.001, device=DEV)
    b = [torch.full((num_buckets,), 1000.0) for _ in offsets]
    a = [torch.zeros(num_buckets) for _ in offsets]
    for step in range(3):
        mask = torch.rand(n, generator=torch.Generator().manual_seed(step)) < 0.3
        m.train_step(ids.to(DEV), step, skip_mask=mask.to(DEV))
        O.logq_train_step(b, a, offsets, ids[~mask], 0.05, step)
    for i, s in enumerate(m.models):
        assert torch.equal(s.b.cpu(), b[i]) and torch.equal(s.a.cpu(), a[i]), i
        assert torch.equal(s.hash_fn(ids[:200].to(DEV)).cpu(), O.logq_hash(ids[:200], offsets[i], num_buckets))
    torch.testing.assert_close(m(ids.view(100, 200).to(DEV)).cpu(), O.logq_forward(b, offsets, ids.view(100, 200)),
                               rtol=1e-6, atol=1e-6)
    single = R.StreamingLogQCorrectionModule(16, 3, device=DEV)
    torch.testing.assert_close(single(ids[:64].to(DEV)).cpu(), torch.full((64,), -4.6052), rtol=0, atol=1e-4)


@pytest.mark.parametrize("div,mod", [(3600, 24), (3600, 168), (86400, 7), (1, 5), (7, 1)])
def test_pattern_from_timelocal(div, mod):
    ts = torch.cat([torch.randint(1_500_000_000, 1_800_000_000, (4, 61), generator=torch.Generator().manual_seed(3)),
                    torch.tensor([[0, 1, -1, -3600, -3601, 3599, 3600, 2 ** 62, -2 ** 62, -2 ** 63, 2 ** 63 - 1]
                                  + [0] * 50])])
    m = R.PatternFromTimelocal(div, mod, 16, device=DEV)
    assert list(m.state_dict()) == ["emb.weight"]
    w = m.emb.weight.detach().cpu()
    idx = O.pattern_index(ts, div, mod)
    assert torch.equal(m.index(ts.to(DEV)).cpu(), idx)
    out = m(ts.to(DEV))
    assert torch.equal(out.cpu(), w[idx])
    go = torch.randn(out.shape)
    out.backward(go.to(DEV))
    torch.testing.assert_close(m.emb.weight.grad.cpu(), O.dense_grad(idx, go, mod), rtol=1e-5, atol=1e-5)
