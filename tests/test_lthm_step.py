"""BASELINE configs[0]: the repaired-harness LTHM training step (tests/harness_lthm.py) on CPU.

  * the plain-torch ("oracle") flavour reproduces the fixture the REFERENCE classes produced
    (tests/golden/lthm_step.npz, written by oracle/make_golden.py) -- runs anywhere;
  * with /root/reference mounted, the reference flavour (imported QueryTower, TransformerBlock,
    KShiftEmbedding, CosineVectorEmbedding, CascadedStreamingLogQCorrectionModule) is re-run live
    and must still equal the fixture bit for bit, and so must the oracle flavour equal it.
The CUDA flavour is compared with the same fixture in tests/test_gpu_lthm_step.py."""
import numpy as np
import pytest
import torch

import harness_lthm as H
from oracle import embedding_oracle as O


def test_oracle_flavour_reproduces_reference_fixture(golden):
    g = golden("lthm_step")
    model = H.model_from_golden(g, H.oracle_layers())
    res = H.run_step(model, H.batch_from_golden(g), steps=2)
    H.compare_with_golden(g, model, res, loss_rtol=1e-7, out_atol=1e-7, out_rtol=1e-6, grad_tol=1e-6, w_tol=1e-6)
    assert res["output"]["current_token_ids"].shape == (256, 42)  # 8 all-pad leading columns trimmed
    assert float(g["margin"]) >= 2e-5


@pytest.mark.skipif(not H.REF.exists(), reason="/root/reference is only mounted in the build container")
def test_reference_flavour_live_equals_fixture_and_oracle(golden):
    g = golden("lthm_step")
    ref = H.model_from_golden(g, H.reference_layers())
    orc = H.model_from_golden(g, H.oracle_layers())
    batch = H.batch_from_golden(g)
    r1 = H.run_step(ref, batch, steps=2)
    r2 = H.run_step(orc, batch, steps=2)
    assert r1["losses"] == r2["losses"] == list(np.asarray(g["losses"]))
    for k in r1["output"]:
        assert torch.equal(r1["output"][k], r2["output"][k]), k
    s1, s2 = ref.state_dict(), orc.state_dict()
    assert set(s1) == set(s2)
    for k in s1:
        assert torch.equal(s1[k], s2[k]), k


def test_reference_trim_semantics():
    """query_tower.py:73-86 on hand-made masks: leading all-pad columns go, export_span columns stay."""
    m = torch.ones(3, 10, dtype=torch.bool)
    m[0, 6:] = False
    m[1, 8:] = False
    assert H.QueryTowerH.reference_trim(m, 2) == 6
    assert H.QueryTowerH.reference_trim(m, 5) == 5      # floor: keep >= export_span columns
    assert H.QueryTowerH.reference_trim(torch.ones(2, 7, dtype=torch.bool), 3) == 4  # everything padded
    m2 = torch.zeros(2, 7, dtype=torch.bool)
    assert H.QueryTowerH.reference_trim(m2, 3) == 0


def test_streaming_logq_oracle_vs_reference_fixture(golden):
    g = golden("streaming_logq")
    ids = torch.from_numpy(g["ids"])
    nb, offs = int(g["num_buckets"]), g["offsets"].tolist()
    b = [torch.full((nb,), 1.0 / float(g["p_init"])) for _ in offs]
    a = [torch.zeros(nb) for _ in offs]
    assert torch.equal(O.logq_forward(b, offs, ids), torch.from_numpy(g["fwd0"]))
    assert torch.allclose(O.logq_forward(b, offs, ids[:5]), torch.full((5,), -4.6052), atol=1e-4)
    for step in range(4):
        sub = ids[torch.randperm(ids.numel(), generator=torch.Generator().manual_seed(step))[:300]]
        O.logq_train_step(b, a, offs, sub, float(g["alpha"]), step)
        assert torch.equal(O.logq_forward(b, offs, ids), torch.from_numpy(g["fwd_steps"][step]))
    assert torch.equal(torch.stack(b), torch.from_numpy(g["b"]))
    assert torch.equal(torch.stack(a), torch.from_numpy(g["a"]))
