"""BASELINE configs[0] on the GPU: the repaired-harness LTHM training step (tests/harness_lthm.py) with
the recommendations_b200 modules swapped in -- KShiftEmbedding (product ids), CosineVectorEmbedding x3
(pooled bags + their backward), FlatEmbedding x2 (action / outcome), PatternFromTimelocal x3 (time
tables), CascadedStreamingLogQCorrectionModule -- against

  (1) the fixture the REFERENCE classes produced on CPU (tests/golden/lthm_step.npz): loss of both
      steps, next_token_emb / current_token_emb, mask, trimmed ids (bit-exact), gradients of every
      embedding table, tables and logQ buffers after the 2nd AdamW step;
  (2) the plain-torch flavour of the same harness run on the SAME GPU, which removes torch's own
      CPU-vs-CUDA differences in the dense layers from the comparison: what is left is the kernels."""
import pytest
import torch

import harness_lthm as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _run(g, L):
    model = H.model_from_golden(g, L, device=DEV)
    res = H.run_step(model, H.batch_from_golden(g, DEV), steps=2)
    return model, res


def test_b200_modules_reproduce_the_reference_lthm_step(golden):
    g = golden("lthm_step")
    torch.backends.cuda.matmul.allow_tf32 = False
    model, res = _run(g, H.b200_layers(DEV))
    rep = {}
    # loss / dense outputs carry torch's CPU-vs-CUDA differences of the transformer (fp32 GEMM order);
    # everything that is an embedding-path quantity sits at the north-star 1e-5
    H.compare_with_golden(g, model, res, loss_rtol=1e-5, out_atol=2e-5, out_rtol=1e-4, grad_tol=1e-5, w_tol=1e-5,
                          report=rep)
    print("deviation vs reference fixture:", {k: f"{v:.2e}" for k, v in rep.items()})


def test_b200_flavour_equals_torch_flavour_on_the_same_gpu(golden):
    g = golden("lthm_step")
    torch.backends.cuda.matmul.allow_tf32 = False
    m1, r1 = _run(g, H.b200_layers(DEV))
    m2, r2 = _run(g, H.oracle_layers())
    assert max(abs(a - b) / abs(b) for a, b in zip(r1["losses"], r2["losses"])) <= 1e-6
    for k in ("current_token_ids", "current_token_mask"):
        assert torch.equal(r1["output"][k], r2["output"][k]), k
    # current_token_emb = product_mapper(masked(emb_mapper(normalize(kshift)) + sum of cosine bags))
    torch.testing.assert_close(r1["output"]["current_token_emb"], r2["output"]["current_token_emb"],
                               rtol=1e-5, atol=1e-6)
    # two transformer layers amplify the 1e-7 differences of the embedding inputs
    torch.testing.assert_close(r1["output"]["next_token_emb"], r2["output"]["next_token_emb"], rtol=1e-4, atol=1e-5)
    for n in H.embedding_param_names(m1):
        a, b = r1["grads"][n], r2["grads"][n]
        assert (a - b).abs().max().item() <= 1e-5 * max(b.abs().max().item(), 1e-30), n
    s1, s2 = m1.state_dict(), m2.state_dict()
    for k in s2:
        if "_log_q_calc" in k:
            torch.testing.assert_close(s1[k], s2[k], rtol=1e-6, atol=0.0)


def test_lthm_step_fused_kshift_table(golden):
    """Same step with the product table in FUSED mode (no Parameter, no dense gradient): the table is
    detached in LTHM training (product_tower.py:47), so it must simply stay untouched and invisible to
    the `0.0 * sum |p|` scan and to AdamW (SURVEY a9 / 8b mode 2)."""
    import recommendations_b200 as R
    g = golden("lthm_step")
    L = H.b200_layers(DEV)
    base = L.KShiftEmbedding
    L.KShiftEmbedding = lambda *a, **k: base(*a, fused_optimizer=R.FusedOptimizerConfig(kind="adagrad"), **k)
    model = H.model_from_golden(g, L, device=DEV)
    assert not any("product_emb_module" in n for n, _ in model.named_parameters())
    w0 = model._model.product_emb_module.emb.weight.clone()
    res = H.run_step(model, H.batch_from_golden(g, DEV), steps=2)
    assert torch.equal(model._model.product_emb_module.emb.weight, w0)
    import numpy as np
    np.testing.assert_allclose(res["losses"], np.asarray(g["losses"]), rtol=1e-5)
