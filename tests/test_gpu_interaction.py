"""tcgen05 dot-interaction kernels vs the CPU restatement (canonical DLRM interaction; the
reference has no implementation: parity unpinned).  bf16 tolerance 1e-2 (north_star)."""
import pytest
import torch

import recommendations_b200 as R
from recommendations_b200 import ops
from oracle import embedding_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def check(got, want):
    scale = want.abs().max().item() + 1e-6
    torch.testing.assert_close(got.float().cpu(), want, rtol=1e-2, atol=1e-2 * scale)


@pytest.mark.parametrize("b,f,d", [(1, 2, 64), (4, 27, 128), (5, 27, 128), (7, 3, 256), (1001, 32, 64),
                                   (4099, 27, 128), (16384, 27, 128)])
def test_dot_interaction_fwd(b, f, d):
    torch.manual_seed(b + f)
    feats = torch.randn(b, f, d).bfloat16()
    got = ops.dot_interaction_fwd(feats.to(DEV))
    assert got.shape == (b, f * (f - 1) // 2) and got.dtype == torch.bfloat16
    check(got, O.dot_interaction(feats))


def test_dot_interaction_fwd_exact_on_small_integers():
    """Integer-valued bf16 inputs: every product and partial sum is exact in fp32 and the result is
    representable in bf16 -> bit-exact check of tile layout, swizzle and triangle packing."""
    b, f, d = 37, 27, 128
    feats = torch.randint(-2, 3, (b, f, d), generator=torch.Generator().manual_seed(3)).bfloat16()
    got = ops.dot_interaction_fwd(feats.to(DEV)).float().cpu()
    want = O.dot_interaction(feats)
    assert want.abs().max() <= 256  # representable exactly in bf16
    assert torch.equal(got, want)


@pytest.mark.parametrize("b,f,d", [(1, 2, 64), (5, 27, 128), (1001, 32, 64), (4099, 27, 128)])
def test_dot_interaction_bwd(b, f, d):
    torch.manual_seed(b)
    feats = torch.randn(b, f, d).bfloat16()
    go = torch.randn(b, f * (f - 1) // 2).bfloat16()
    got = ops.dot_interaction_bwd(feats.to(DEV), go.to(DEV))
    ref_in = feats.float().requires_grad_(True)
    O.dot_interaction(ref_in).backward(go.float())
    check(got, ref_in.grad)


def test_dot_interaction_bwd_exact_on_small_integers():
    b, f, d = 9, 27, 128
    g = torch.Generator().manual_seed(4)
    feats = torch.randint(-1, 2, (b, f, d), generator=g).bfloat16()
    go = torch.randint(-2, 3, (b, f * (f - 1) // 2), generator=g).bfloat16()
    got = ops.dot_interaction_bwd(feats.to(DEV), go.to(DEV)).float().cpu()
    ref_in = feats.float().requires_grad_(True)
    O.dot_interaction(ref_in).backward(go.float())
    assert torch.equal(got, ref_in.grad)


def test_module_forward_backward():
    b, f, d = 513, 26, 128
    torch.manual_seed(5)
    dense = torch.randn(b, d).bfloat16()
    sparse = torch.randn(b, f, d).bfloat16()
    dd, ss = dense.to(DEV).requires_grad_(True), sparse.to(DEV).requires_grad_(True)
    out = R.DotInteraction()(dd, ss)
    assert out.shape == (b, d + (f + 1) * f // 2)
    go = torch.randn(out.shape).bfloat16()
    out.backward(go.to(DEV))
    dr, sr = dense.float().requires_grad_(True), sparse.float().requires_grad_(True)
    t = torch.cat([dr.unsqueeze(1), sr], dim=1)
    ref = torch.cat([dr, O.dot_interaction(t)], dim=1)
    check(out, ref.detach())
    ref.backward(go.float())
    check(ss.grad, sr.grad)
    check(dd.grad, dr.grad)


@pytest.mark.parametrize("mode", ["sum", "mean"])
@pytest.mark.parametrize("fused", [False, True])
def test_pooled_interaction_pipeline_equals_lookup_permute_cat_interaction(mode, fused):
    """PooledInteraction (the pooled rows written feature-interleaved straight into the interaction's input,
    the interaction's gradient read in place by the pooled backward) == EmbeddingCollection -> permute ->
    DotInteraction (two concatenations): outputs bit-identical, gradients / fused updates equal."""
    import recommendations_b200 as R
    t, n_rows, dim, b, p = 5, 3001, 128, 300, 7
    g = torch.Generator().manual_seed(12)
    ids = torch.randint(-2 ** 40, 2 ** 40, (t, b, p), generator=g, dtype=torch.int64).to(DEV)
    lengths = torch.randint(0, p + 1, (t, b), generator=g).to(DEV)
    dense = torch.randn(b, dim, generator=g).to(torch.bfloat16).to(DEV)
    go = torch.randn(b, dim + (t + 1) * t // 2, generator=g).to(torch.bfloat16).to(DEV)
    opt = R.FusedOptimizerConfig(kind="rowwise_adagrad", lr=0.05) if fused else None
    mk = lambda: R.EmbeddingCollection(t, n_rows, dim, kind="pooled", mode=mode, dtype=torch.bfloat16, device=DEV,  # noqa: E731
                                       fused_optimizer=opt)
    c1, c2 = mk(), mk()
    c2.load_state_dict(c1.state_dict())
    d1 = dense.clone().requires_grad_(True)
    d2 = dense.clone().requires_grad_(True)
    out1 = R.PooledInteraction(c1)(d1, ids, lengths)
    out2 = R.DotInteraction()(d2, c2(ids, lengths).permute(1, 0, 2))
    assert out1.shape == out2.shape == (b, dim + (t + 1) * t // 2)
    assert torch.equal(out1, out2)
    out1.backward(go)
    out2.backward(go)
    torch.testing.assert_close(d1.grad.float(), d2.grad.float(), rtol=1e-2, atol=1e-2)
    if fused:
        torch.testing.assert_close(c1.table.weight.float(), c2.table.weight.float(), rtol=1e-2, atol=1e-2)
        torch.testing.assert_close(c1.table.opt_state1, c2.table.opt_state1, rtol=1e-2, atol=1e-4)
    else:
        for m1, m2 in zip(c1.members(), c2.members()):
            torch.testing.assert_close(m1.weight.grad.float(), m2.weight.grad.float(), rtol=1e-2, atol=1e-2)
