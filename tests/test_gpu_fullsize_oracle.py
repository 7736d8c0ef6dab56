"""FLOAT data at the full BASELINE sizes against the CPU oracle (one table / feature per config: the
oracle finishes each in a few seconds).  Gathered rows bit-exact; pooled sums, updated weights and
optimizer state inside the north-star 1e-5 (fp32) / 1e-2 (bf16).  Adagrad comparisons start from a
NON-ZERO accumulator (a torch.optim.Adagrad option) so that every element is well conditioned; the
zero-start case is compared with the demonstrated summation budget of tests/tolerances.py.
tests/test_gpu_fullsize.py holds the integer-valued property tests of the same sizes."""
import pytest
import torch

import recommendations_b200 as R
from recommendations_b200 import _native as N
from recommendations_b200 import ops
from oracle import embedding_oracle as O
from conftest import seeded_ids
from tolerances import EPS32, dense_grad64

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _adagrad64(w0, G, s0, lr, eps=1e-10):
    s = s0.double() + G * G
    return w0.double() - lr * G / (s.sqrt() + eps), s


@pytest.mark.parametrize("acc0", [0.1, 0.0])
def test_cfg2_one_table_float_parity(acc0):
    """cfg 2, table 0: FlatEmbedding(1M, 64) fp32, ids [8192, 200] (seed 1000), fwd + bwd + fused
    element-wise Adagrad(lr 0.5) -- embedding_module_gen.py:137's optimizer."""
    n_rows, dim, b, l, lr = 1_000_000, 64, 8192, 200, 0.5
    ids = seeded_ids(b * l, 1000, (b, l))
    torch.manual_seed(1234)
    w0 = torch.randn(n_rows, dim)
    go = torch.randn(b * l, dim, generator=torch.Generator().manual_seed(4321))
    m = R.FlatEmbedding(n_rows, dim, device=DEV, fused_optimizer=R.FusedOptimizerConfig(
        kind="adagrad", lr=lr, initial_accumulator_value=acc0))
    m.load_state_dict({"_emb_table.weight": w0})
    out = m(ids.to(DEV))
    rows = O.row_index(ids, n_rows, 0)
    assert torch.equal(out.cpu().view(-1, dim), w0[rows.view(-1)])           # gathered rows: bit-exact
    out.backward(go.view(b, l, dim).to(DEV))
    G, A = dense_grad64(rows, go, n_rows)
    w64, s64 = _adagrad64(w0, G, torch.full((n_rows, dim), acc0), lr)
    got_w, got_s = m._emb_table.weight.cpu().double(), m._emb_table.opt_state1.cpu().double()
    # state: sum of squares of a sum of ~1.6 terms -- always well conditioned relative to its own size
    torch.testing.assert_close(got_s, s64, rtol=1e-5, atol=1e-6)
    err = (got_w - w64).abs()
    if acc0 > 0:
        assert (err <= 1e-5 + 1e-5 * w64.abs()).all(), float(err.max())
    else:
        # zero start: the update is lr * G / |G|; fp32 summation noise of G matters only where |G| is
        # within ~ 8 eps32 * sum |g_i| of zero -- those elements may move by up to 2 * lr
        budget = torch.minimum(lr * 8 * EPS32 * A / (G.abs() + 1e-10), torch.full_like(A, 2 * lr))
        assert (err <= 1e-5 + 1e-5 * w64.abs() + budget).all(), float(err.max())
        assert (budget > 1e-5).float().mean().item() < 0.05  # worst-case (c = 8) budget: ~1 % of the elements
    untouched = torch.ones(n_rows, dtype=torch.bool)
    untouched[rows.view(-1)] = False
    assert torch.equal(m._emb_table.weight.cpu()[untouched], w0[untouched])   # never indexed: bit-identical


def test_cfg3_one_feature_bf16_parity():
    """cfg 3, feature 0: PooledEmbeddingBag(1M, 128) bf16, B 16384, P 20, ragged lengths, pooled sum fwd
    (fp32 accumulation in slot order -> bit-exact vs the oracle), bwd + fused row-wise Adagrad."""
    n_rows, dim, b, p, lr = 1_000_000, 128, 16384, 20, 0.1
    g = torch.Generator().manual_seed(3000)
    ids = torch.randint(0, n_rows, (b, p), generator=g, dtype=torch.int64)
    lengths = torch.randint(1, p + 1, (b,), generator=torch.Generator().manual_seed(300), dtype=torch.int32)
    torch.manual_seed(1234)
    w0 = torch.randn(n_rows, dim).bfloat16()
    go = torch.randn(b, dim, generator=torch.Generator().manual_seed(4321)).bfloat16()
    m = R.PooledEmbeddingBag(n_rows, dim, mode="sum", hash_ids=False, dtype=torch.bfloat16, device=DEV,
                             fused_optimizer=R.FusedOptimizerConfig(kind="rowwise_adagrad", lr=lr,
                                                                    initial_accumulator_value=0.1))
    m.load_state_dict({"emb.weight": w0})
    out = m(ids.to(DEV), lengths.to(DEV))
    want = O.pooled_bag(w0, ids, lengths=lengths, hash_ids=False)
    assert torch.equal(out.cpu(), want)
    out.backward(go.to(DEV))
    use = torch.arange(p).unsqueeze(0) < lengths.unsqueeze(1)
    rows = ids[use]
    gsl = go.float().unsqueeze(1).expand(b, p, dim)[use]
    G, _ = dense_grad64(rows, gsl, n_rows)
    touched = torch.zeros(n_rows, dtype=torch.bool)
    touched[rows] = True
    s64 = torch.full((n_rows,), 0.1, dtype=torch.float64)
    s64[touched] += (G[touched] ** 2).mean(dim=1)
    w64 = w0.double()
    w64[touched] -= lr * G[touched] / (s64[touched].sqrt() + 1e-10).unsqueeze(1)
    torch.testing.assert_close(m.emb.opt_state1.cpu().double(), s64, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(m.emb.weight.cpu().double(), w64, rtol=1e-2, atol=1e-2)
    # tighter than the north star where it can be: the bf16 table holds the fp32 result rounded once
    err = (m.emb.weight.cpu().double() - w64.float().bfloat16().double()).abs()
    assert (err <= 2.0 ** -7 * w64.abs() + 1e-6).all()                        # <= 1 bf16 ulp
    assert (err == 0).float().mean().item() > 0.99
    assert torch.equal(m.emb.weight.cpu()[~touched], w0[~touched])


def _zipf_rows(n, n_rows, alpha, seed):
    """SURVEY section 8d cfg 4: truncated Zipf(alpha) over ranks 1..N by inverse CDF in float64, ranks mapped
    through a fixed permutation so hot rows are scattered."""
    g = torch.Generator().manual_seed(seed)
    cdf = torch.cumsum(torch.arange(1, n_rows + 1, dtype=torch.float64) ** -alpha, 0)
    cdf /= cdf[-1].clone()
    ranks = torch.searchsorted(cdf, torch.rand(n, generator=g, dtype=torch.float64)).clamp_(max=n_rows - 1)
    return torch.randperm(n_rows, generator=g)[ranks]


def test_cfg4_one_table_zipf_float_parity():
    """cfg 4, table 0: FlatEmbedding(1M, 64) fp32, ids [4096, 1024] Zipf(1.05) (top row ~ 9.5 % of the 4.2 M
    lookups), fwd gather + bwd + fused row-wise Adagrad."""
    n_rows, dim, b, l, lr = 1_000_000, 64, 4096, 1024, 0.1
    ids = _zipf_rows(b * l, n_rows, 1.05, 2000).view(b, l)
    torch.manual_seed(1234)
    w0 = torch.randn(n_rows, dim)
    go = torch.randn(b * l, dim, generator=torch.Generator().manual_seed(4321))
    m = R.FlatEmbedding(n_rows, dim, device=DEV, fused_optimizer=R.FusedOptimizerConfig(
        kind="rowwise_adagrad", lr=lr, initial_accumulator_value=0.1))
    m.load_state_dict({"_emb_table.weight": w0})
    out = m(ids.to(DEV))
    assert torch.equal(out.cpu().view(-1, dim), w0[ids.view(-1)])
    out.backward(go.view(b, l, dim).to(DEV))
    G, A = dense_grad64(ids, go, n_rows)
    counts = torch.bincount(ids.view(-1), minlength=n_rows)
    assert counts.max().item() > 0.08 * b * l                                 # dedup-heavy: one row, ~400 k slots
    touched = counts > 0
    s64 = torch.full((n_rows,), 0.1, dtype=torch.float64)
    s64[touched] += (G[touched] ** 2).mean(dim=1)
    w64 = w0.double()
    w64[touched] -= lr * G[touched] / (s64[touched].sqrt() + 1e-10).unsqueeze(1)
    got_w, got_s = m._emb_table.weight.cpu().double(), m._emb_table.opt_state1.cpu().double()
    # fp32 sums of up to 400 k terms: 1e-5 relative + the summation bound 8 eps32 sum|g_i| mapped through
    # the update (d w = lr dG / sqrt(s)); the state is a mean of squares of those sums
    dG = 8 * EPS32 * A
    tol_w = 1e-5 + 1e-5 * w64.abs() + lr * dG / (s64.sqrt() + 1e-10).unsqueeze(1)
    assert ((got_w - w64).abs() <= tol_w).all(), float(((got_w - w64).abs() / tol_w).max())
    tol_s = 1e-6 + 1e-5 * s64 + (2 * G.abs() * dG).mean(dim=1)
    assert ((got_s - s64).abs() <= tol_s).all()
    # and the plain north-star statement on all rows with fewer than 1000 lookups (99.9 % of the touched rows)
    cool = touched & (counts < 1000)
    assert ((got_w - w64).abs()[cool] <= 2e-5 + 1e-5 * w64.abs()[cool]).all()
    assert cool.sum().item() > 0.99 * touched.sum().item()
    assert torch.equal(m._emb_table.weight.cpu()[~touched], w0[~touched])
