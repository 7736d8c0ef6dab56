"""Size-independent properties at the FULL sizes of BASELINE.json configs 3, 4 and the k-shift series
(the oracle cannot finish these in seconds): integer-valued tables / gradients make every fp32 sum
exact, so counts, checksums of checksums and linearity are bit-exact statements about the kernels."""
import pytest
import torch

import recommendations_b200 as R
from recommendations_b200 import _native as N
from recommendations_b200 import ops
from conftest import seeded_ids

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_cfg2_kshift_full_size_counts_and_checksums():
    """KShiftEmbedding(1M, 64, k = 8) on ids [8192, 200]: with w[r] = 1 the fused bag must equal k
    (before the 1/sqrt(k) scale); with grad == sqrt(k) and SGD lr 1 from zero, w[r] = -(slots on r):
    the row histogram of all 13.1 M (id, shift) slots, checked against recemb_row_index."""
    n_rows, dim, k, b, l = 1_000_000, 64, 8, 8192, 200
    ids = seeded_ids(b * l, 1000, (b, l)).to(DEV)
    m = R.KShiftEmbedding(n_rows, dim, num_shifts=k, device=DEV,
                          fused_optimizer=R.FusedOptimizerConfig(kind="sgd", lr=1.0))
    w = m.emb.weight
    w.fill_(1.0)
    out = m(ids)
    want = (torch.tensor(float(k)) / torch.tensor(float(k)).sqrt()).to(DEV)   # fp32: k ones / sqrt(k)
    assert torch.equal(out, want.expand_as(out))
    w.zero_()
    out.backward(torch.full_like(out, 4.0))
    counts = torch.zeros(n_rows, device=DEV)
    for c in range(k):
        counts += torch.bincount(ops.row_index(ids, N.HASH_ROTL_FLOORMOD, n_rows, c).view(-1), minlength=n_rows)
    assert float(counts.double().sum()) == float(b * l * k)
    # every slot adds -(4 / sqrt(8)) to its row: compare in float64 against count * that constant
    unit = float(torch.tensor(4.0) * (1.0 / torch.tensor(8.0).sqrt()))
    torch.testing.assert_close(w[:, 0].double(), -counts.double() * unit, rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(w[:, 63].double(), w[:, 0].double(), rtol=0, atol=0)
    # the collapse (commons/layers.py:182 arithmetic >>): about half of each shift >= 1 sits on 2^(c-1) rows
    assert float(counts[-1]) > 0.4 * b * l


def test_cfg3_full_size_pooled_counts_and_interaction():
    """cfg 3: 26 tables x [1M, 128] bf16, B 16384, P 20, ragged lengths.  Table rows = 1 -> pooled sum
    = the bag length exactly; the dot interaction of constant rows is length_i * length_j * 128 for
    feature pairs; backward with grad 1 and SGD: w = 1 - (slots on the row) exactly (small integers)."""
    f, n_rows, dim, b, p = 26, 1_000_000, 128, 16384, 20
    g = torch.Generator(device=DEV).manual_seed(3000)
    ids = torch.randint(0, n_rows, (f * b, p), generator=g, device=DEV, dtype=torch.int64)
    lengths = torch.randint(1, 9, (f * b,), generator=g, device=DEV, dtype=torch.int32)   # <= 8: products stay exact in bf16
    w = torch.ones(f * n_rows, dim, device=DEV, dtype=torch.bfloat16)
    pooled = ops.pool_fwd(w, ids, lengths=lengths, num_rows=n_rows, bags_per_table=b, num_tables=f,
                          hash_mode=N.HASH_IDENTITY)
    assert torch.equal(pooled[:, 0].float(), lengths.float()) and torch.equal(pooled[:, 127].float(), lengths.float())
    feats = pooled.view(f, b, dim).permute(1, 0, 2).contiguous()                 # [B, 26, 128]
    z = ops.dot_interaction_fwd(feats)                                          # [B, 325]
    len_bf = lengths.view(f, b).t().float()                                     # [B, 26]
    ii, jj = torch.tril_indices(f, f, -1, device=DEV)
    want = (len_bf[:, ii] * len_bf[:, jj] * dim).to(torch.bfloat16)             # <= 64 * 128 = 8192: exact in bf16
    assert torch.equal(z, want)
    # backward: one plan over all 26 tables, SGD lr 1, grad 1 per bag -> row r loses one per valid slot
    plan = ops.BackwardPlan.build(ids, num_rows=n_rows, hash_mode=N.HASH_IDENTITY, bag_size=p, lengths=lengths,
                                  ids_per_table=b * p, num_tables=f)
    n_valid, n_unique = (int(v) for v in plan.counters.cpu())
    assert n_valid == int(lengths.sum())
    go = torch.ones(f * b, dim, device=DEV, dtype=torch.bfloat16)
    ops.bwd_apply(plan, go, table=w, update=N.UPD_SGD, slots_per_grad_row=p, hp=ops.make_optim_params(lr=1.0))
    use = torch.arange(p, device=DEV).unsqueeze(0) < lengths.unsqueeze(1)
    t_of = (torch.arange(f * b, device=DEV) // b).unsqueeze(1)
    counts = torch.bincount((t_of * n_rows + ids)[use], minlength=f * n_rows).float()
    assert int((counts > 0).sum()) == n_unique
    assert torch.equal(w[:, 0].float(), 1.0 - counts) and torch.equal(w[:, 77].float(), 1.0 - counts)


def test_cfg4_full_size_zipf_rowwise_properties():
    """cfg 4: 10 tables x [1M, 64] fp32, B 4096 x L 1024 Zipf(1.05) rows (top row ~ 9.5 % of the
    lookups -> multi-level records).  Gather: every row carries its index.  Backward with SGD and
    grad 1: w = -(lookups of the row); row-wise Adagrad from zero state: state = count^2 and
    w = -lr * count / (count + eps) -> -lr for every touched row, untouched rows stay 0."""
    t, b, l, n_rows, dim = 10, 4096, 1024, 1_000_000, 64
    n = b * l
    g = torch.Generator(device=DEV).manual_seed(2000)
    ranks = torch.arange(1, n_rows + 1, device=DEV, dtype=torch.float64)
    cdf = torch.cumsum(ranks.pow(-1.05), 0)
    cdf /= cdf[-1].clone()
    perm = torch.randperm(n_rows, generator=g, device=DEV)
    ids = torch.cat([perm[torch.searchsorted(cdf, torch.rand(n, generator=g, device=DEV, dtype=torch.float64))
                          .clamp_(max=n_rows - 1)] for _ in range(t)])
    w = torch.arange(t * n_rows, device=DEV, dtype=torch.float32).remainder(n_rows).unsqueeze(1).expand(-1, dim).contiguous()
    out = torch.empty(t * n, dim, device=DEV)
    ops.gather_fwd(w, ids, out=out, ids_per_table=n, hash_mode=N.HASH_IDENTITY)
    assert torch.equal(out[:, 0].long(), ids) and torch.equal(out[:, 63].long(), ids)
    del out
    t_of = (torch.arange(t * n, device=DEV) // n)
    counts = torch.bincount(t_of * n_rows + ids, minlength=t * n_rows).float()
    assert float(counts.max()) > 0.05 * n                                        # the Zipf head really is hot
    plan = ops.BackwardPlan.build(ids, num_rows=n_rows, hash_mode=N.HASH_IDENTITY, ids_per_table=n)
    grad = torch.ones(t * n, dim, device=DEV)
    w.zero_()
    ops.bwd_apply(plan, grad, table=w, update=N.UPD_SGD, hp=ops.make_optim_params(lr=1.0))
    assert torch.equal(w[:, 0], -counts) and torch.equal(w[:, 31], -counts)     # exact: integers < 2^24
    assert float(w[:, 0].double().sum()) == -float(t * n)
    w.zero_()
    state = torch.zeros(t * n_rows, device=DEV)
    ops.bwd_apply(plan, grad, table=w, update=N.UPD_ROWWISE_ADAGRAD, state1=state,
                  hp=ops.make_optim_params(lr=0.25, eps=1e-10))
    torch.testing.assert_close(state, counts * counts, rtol=1e-5, atol=0)
    touched = counts > 0
    torch.testing.assert_close(w[touched][:, 5], torch.full((int(touched.sum()),), -0.25, device=DEV), rtol=1e-5, atol=0)
    assert float(w[~touched].abs().sum()) == 0.0
