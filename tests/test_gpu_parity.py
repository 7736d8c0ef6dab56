"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference-generated
golden vectors.  Bit-exact for indices, gathered rows, dedup; <= 1e-5 relative (fp32) /
1e-2 (bf16) for sums, normalised outputs and updated weights (BASELINE.json north_star)."""
import math

import numpy as np
import pytest
import torch

import recommendations_b200 as R
from recommendations_b200 import _native as N
from recommendations_b200 import ops
from oracle import embedding_oracle as O
from conftest import seeded_ids
from tolerances import (assert_cross_device_trajectory, assert_sums_close, dense_grad64 as _dense_grad64, kshift_adagrad_budget,
                        kshift_adagrad_error_bound as _kshift_adagrad_error_bound,
                        assert_adagrad_trajectory_close as _assert_adagrad_trajectory_close)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL32, ATOL32 = 1e-5, 1e-5  # atol: sums of O(10) unit-variance terms cancel to ~0
RTOL16, ATOL16 = 1e-2, 1e-2


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(got, want, bf16=False):
    torch.testing.assert_close(got.detach().cpu().float(), want.detach().float(),
                               rtol=RTOL16 if bf16 else RTOL32, atol=ATOL16 if bf16 else ATOL32)


# ------------------------------------------------------------------- hashing ----
def test_row_index_golden(golden):
    g = golden("row_index")
    ids = T(g["ids"]).to(DEV)
    for si, n_rows in enumerate(g["sizes"].tolist()):
        for ci, c in enumerate(g["cols"].tolist()):
            got = ops.row_index(ids, N.HASH_ROTL_FLOORMOD, n_rows, c).cpu()
            assert torch.equal(got, T(g["rows"][si, ci])), (n_rows, c)
        got0 = ops.row_index(ids, N.HASH_FLOORMOD, n_rows).cpu()
        assert torch.equal(got0, T(g["rows"][si, 0]))


@pytest.mark.parametrize("n_rows", [1, 2, 1000, 999983, 1 << 20, (1 << 31) - 1, (1 << 40) + 3, (1 << 62) + 1])
def test_row_index_random_vs_oracle(n_rows):
    ids = seeded_ids(100003, 31)
    for c in (0, 1, 13, 32, 63):
        got = ops.row_index(ids.to(DEV), N.HASH_ROTL_FLOORMOD, n_rows, c).cpu()
        assert torch.equal(got, O.row_index(ids, n_rows, c)), c


def test_qr_index_vs_oracle():
    ids = torch.cat([T(np.array([0, 1, -1, -2 ** 63, 2 ** 63 - 1])), seeded_ids(5000, 32)])
    for n_emb in (10007, 1 << 34, 5):
        q, r, d = O.qr_indices(ids, n_emb)
        assert torch.equal(ops.row_index(ids.to(DEV), N.HASH_QR_QUOTIENT, d, d).cpu(), q)
        assert torch.equal(ops.row_index(ids.to(DEV), N.HASH_QR_REMAINDER, d, d).cpu(), r)


def test_module_get_row_idx_matches_survey_known_answers():
    m = R.KShiftEmbedding(1000, 8, num_shifts=4, device=DEV)
    ids = torch.tensor([0, 1, -1, 2 ** 62, -2 ** 63, 2 ** 63 - 1, 12345678901234, -987654321], device=DEV)
    assert m.get_row_idx(ids, 0).tolist() == [0, 1, 999, 904, 192, 807, 234, 679]
    assert m.get_row_idx(ids, 1).tolist() == [0, 2, 999, 192, 999, 998, 468, 999]
    assert m.get_row_idx(ids, 3).tolist() == [0, 8, 999, 2, 996, 995, 872, 999]


# ------------------------------------------------------------- forward gather ----
def test_flat_embedding_golden(golden):
    g = golden("flat_embedding")
    m = R.FlatEmbedding(1000, 32, padding_idx=0, device=DEV)
    m.load_state_dict({"_emb_table.weight": T(g["weight"])})
    ids = T(g["ids"]).to(DEV)
    assert torch.equal(m(ids).cpu(), T(g["out"]))  # gathered rows: bit-exact
    mn = R.FlatEmbedding(1000, 32, normalize_output=True, device=DEV)
    mn.load_state_dict({"_emb_table.weight": T(g["weight"])})
    close(mn(ids), T(g["out_norm"]))


@pytest.mark.parametrize("dim,dtype", [(4, torch.float32), (8, torch.float32), (32, torch.float32),
                                       (64, torch.float32), (96, torch.float32), (128, torch.float32),
                                       (512, torch.float32), (1024, torch.float32),
                                       (8, torch.bfloat16), (64, torch.bfloat16), (128, torch.bfloat16)])
@pytest.mark.parametrize("n", [0, 1, 2, 255, 1024, 1025, 40001])
def test_gather_shapes_bit_exact(dim, dtype, n):
    torch.manual_seed(dim + n)
    w = torch.randn(3001, dim).to(dtype)
    ids = seeded_ids(n, 33 + n)
    out, _ = ops.gather_fwd(w.to(DEV), ids.to(DEV))
    assert out.shape == (n, dim)
    assert torch.equal(out.cpu(), O.flat_embedding(w, ids))


def test_gather_misaligned_ids_and_views():
    w = torch.randn(500, 64)
    base = seeded_ids(2051, 34).to(DEV)
    ids = base[1:]  # 8-byte aligned only -> non-bulk staging path
    assert ids.data_ptr() % 16 == 8
    out, _ = ops.gather_fwd(w.to(DEV), ids)
    assert torch.equal(out.cpu(), O.flat_embedding(w, ids.cpu()))
    ids2 = base.view(7, 293)  # multi-dim ids keep their shape
    out2, _ = ops.gather_fwd(w.to(DEV), ids2)
    assert out2.shape == (7, 293, 64)


def test_gather_fused_pad_mask():
    w = torch.randn(100, 32)
    ids = seeded_ids(4 * 50, 35, (4, 50))
    ids[:, 30:] = 0
    m = R.FlatEmbedding(100, 32, fused_pad_mask=True, device=DEV)
    m.load_state_dict({"_emb_table.weight": w})
    got = m(ids.to(DEV)).cpu()
    want = O.flat_embedding(w, ids).masked_fill((ids == 0).unsqueeze(-1), 0.0)
    assert torch.equal(got, want)


def test_qr_embedding_golden(golden):
    g = golden("qr_embedding")
    for norm, key in ((True, "out_norm"), (False, "out_plain")):
        m = R.QREmbedding(int(g["num_embeddings"]), 32, norm, device=DEV)
        m.load_state_dict({"emb_q.weight": T(g["weight_q"]), "emb_r.weight": T(g["weight_r"])})
        got = m(T(g["ids"]).to(DEV))
        if norm:
            close(got, T(g[key]))
        else:
            assert torch.equal(got.cpu(), T(g[key]))


# ----------------------------------------------------------------- k-shift fwd ----
def test_kshift_golden(golden):
    g = golden("kshift_embedding")
    ids = T(g["ids"]).to(DEV)
    for k, norm in ((4, False), (8, False), (16, True), (16, False)):
        m = R.KShiftEmbedding(1009, 32, num_shifts=k, normalize_output=norm, device=DEV)
        m.load_state_dict({"emb.weight": T(g["weight"])})
        close(m(ids), T(g[f"out_k{k}_{'norm' if norm else 'scale'}"]))


@pytest.mark.parametrize("dim,dtype,k", [(4, torch.float32, 16), (32, torch.float32, 16),
                                         (64, torch.float32, 8), (64, torch.float32, 1),
                                         (128, torch.float32, 3), (512, torch.float32, 8),
                                         (64, torch.bfloat16, 8), (128, torch.bfloat16, 16)])
def test_kshift_vs_oracle(dim, dtype, k):
    torch.manual_seed(k)
    w = torch.randn(7919, dim).to(dtype)
    ids = seeded_ids(5003, 36)
    bf = dtype == torch.bfloat16
    for norm in (False, True):
        out, inv = ops.kshift_fwd(w.to(DEV), ids.to(DEV), k, N.EPI_L2NORM if norm else N.EPI_RSQRT_K,
                                  want_inv_norm=True)
        close(out, O.kshift_embedding(w.float(), ids, k, norm), bf16=bf)
        if norm:
            pre = O.kshift_embedding(w.float(), ids, k, False) * math.sqrt(k)
            close(inv, 1.0 / pre.norm(dim=-1).clamp_min(1e-12), bf16=bf)
    # pre-epilogue sums are bit-exact in fp32: same order of adds as the reference
    if not bf:
        out, _ = ops.kshift_fwd(w.to(DEV), ids.to(DEV), k, N.EPI_NONE)
        ref = O.kshift_embedding(w, ids, k, False) * 0  # shape only
        acc = torch.nn.functional.embedding(O.row_index(ids, 7919, 0), w)
        for c in range(1, k):
            acc = acc + torch.nn.functional.embedding(O.row_index(ids, 7919, c), w)
        assert ref.shape == out.shape and torch.equal(out.cpu(), acc)


# ------------------------------------------------------------------ pooled fwd ----
@pytest.mark.parametrize("dim,dtype", [(16, torch.float32), (64, torch.float32), (128, torch.bfloat16),
                                       (512, torch.float32)])
@pytest.mark.parametrize("p", [1, 7, 20, 32, 33])
def test_pool_variants_vs_oracle(dim, dtype, p):
    torch.manual_seed(p)
    bf = dtype == torch.bfloat16
    w = torch.randn(2003, dim).to(dtype)
    m = 1237
    ids = seeded_ids(m * p, 37, (m, p))
    lengths = torch.randint(0, p + 1, (m,), generator=torch.Generator().manual_seed(p))
    wd = w.to(DEV)
    got = ops.pool_fwd(wd, ids.to(DEV))
    want = O.pooled_bag(w, ids)
    if bf:
        close(got, want, bf16=True)
    else:
        assert torch.equal(got.cpu(), want)  # in-order fp32 sum == CPU EmbeddingBag order
    got = ops.pool_fwd(wd, ids.to(DEV), lengths=lengths.to(DEV))
    close(got, O.pooled_bag(w, ids, lengths=lengths), bf16=bf)
    got = ops.pool_fwd(wd, ids.to(DEV), lengths=lengths.to(DEV), pool_mode=N.POOL_MEAN)
    close(got, O.pooled_bag(w, ids, lengths=lengths, mode="mean"), bf16=bf)
    got = ops.pool_fwd(wd, ids.to(DEV), lengths=lengths.to(DEV), last_n=3)
    close(got, O.pooled_bag(w, ids, lengths=lengths, last_n=3), bf16=bf)
    ids0 = ids.clone()
    ids0[:, ::3] = 0
    got = ops.pool_fwd(wd, ids0.to(DEV), zero_pad=True, pad_id=0)
    close(got, O.pooled_bag(w, ids0, skip_pad=True), bf16=bf)
    psw = torch.rand(m, p)
    got = ops.pool_fwd(wd, ids.to(DEV), per_slot_weight=psw.to(DEV))
    close(got, O.pooled_bag(w, ids, per_sample_weights=psw), bf16=bf)


def test_pool_equals_torch_embedding_bag():
    w = torch.randn(1000, 64)
    idx = torch.randint(0, 1000, (5000, 20))
    got = ops.pool_fwd(w.to(DEV), idx.to(DEV), hash_mode=N.HASH_IDENTITY)
    assert torch.equal(got.cpu(), O.embedding_bag_sum(w, idx))


def test_cosine_vector_embedding_golden(golden):
    g = golden("cosine_vector_embedding")
    m = R.CosineVectorEmbedding(32, 64, n_proj=32, num_bins=12, device=DEV)
    m.load_state_dict({k: T(g[k]) for k in ("projection_mat", "grid", "pos_offset")} |
                      {"emb.weight": T(g["weight"])})
    x = T(g["x"]).to(DEV)
    idxs = m.bucket_indices(x).cpu()
    agree = (idxs == T(g["idxs"])).float().mean().item()
    assert agree >= 0.999  # bucket edges are a float compare of a GPU vs CPU matmul
    # the bag-sum op itself on the reference's own indices: exact
    got = ops.pool_fwd(m.emb.weight.detach(), T(g["idxs"]).to(DEV), hash_mode=N.HASH_IDENTITY)
    assert torch.equal(got.cpu().view(g["out"].shape), T(g["out"]))
    # the module's own bag + backward on the reference's indices: unconditional
    out = m.bag(T(g["idxs"]).to(DEV)).view(g["out"].shape)
    assert torch.equal(out.cpu(), T(g["out"]))
    out.backward(T(g["grad_out"]).to(DEV))
    close(m.emb.weight.grad, T(g["grad_weight"]))
    if agree == 1.0:  # and end to end when no projection sits on a bucket edge of this GPU's matmul
        assert torch.equal(m(x).cpu(), T(g["out"]))


# -------------------------------------------------------------- backward: plan ----
@pytest.mark.parametrize("n,n_rows", [(1, 10), (33, 5), (1025, 1000), (70001, 997), (200000, 1 << 20)])
def test_plan_is_sorted_stable_and_complete(n, n_rows):
    ids = seeded_ids(n, 38)
    plan = ops.BackwardPlan.build(ids.to(DEV), num_rows=n_rows)
    rows, slots = plan.sorted_rows.cpu(), plan.sorted_slots.cpu()
    want_rows = O.row_index(ids, n_rows, 0)
    order = torch.argsort(want_rows, stable=True)
    assert torch.equal(rows, want_rows[order])       # sortedness + multiset
    assert torch.equal(slots, order)                 # stable in slot
    n_valid, n_unique = plan.counters.cpu().tolist()
    assert n_valid == n and n_unique == torch.unique(want_rows).numel()  # dedup count bit-exact


@pytest.mark.parametrize("n_rows", [1, 5, 4095, 4096, 4097, (1 << 24) - 1, (1 << 24) + 5, (1 << 31) + 11])
@pytest.mark.parametrize("n,skew", [(1, 0), (31, 0), (4097, 0), (70001, 0), (70001, 2), (3_000_000, 0), (3_000_000, 1)])
def test_plan_radix_sort_passes_and_skew(n, n_rows, skew):
    """The hand-written LSD radix sort (csrc/sort.cu) behind every plan: 1 / 2 / 3 passes (key widths
    1 ... 32 bits), one tile and many tiles, uniform keys, Zipf-like heads (skew 1: half of the slots on
    4 rows) and a single hot row (skew 2) -- sorted, stable in slot, complete."""
    g = torch.Generator().manual_seed(n + skew)
    rows = torch.randint(0, n_rows, (n,), generator=g, dtype=torch.int64)
    if skew == 1:
        hot = torch.randint(0, n_rows, (4,), generator=g, dtype=torch.int64)
        pick = torch.rand(n, generator=g) < 0.5
        rows[pick] = hot[torch.randint(0, 4, (int(pick.sum()),), generator=g)]
    elif skew == 2:
        rows[:] = n_rows - 1
        rows[::7] = 0
    plan = ops.BackwardPlan.build(rows.to(DEV), num_rows=n_rows, hash_mode=N.HASH_IDENTITY)
    order = torch.argsort(rows, stable=True)
    assert torch.equal(plan.sorted_rows.cpu(), rows[order])
    assert torch.equal(plan.sorted_slots.cpu(), order)
    n_valid, n_unique = plan.counters.cpu().tolist()
    assert n_valid == n and n_unique == torch.unique(rows).numel()


def test_plan_drops_padding_and_window():
    m, p, n_rows = 50, 8, 64
    ids = seeded_ids(m * p, 39, (m, p))
    ids[:, 5] = 0
    lengths = torch.randint(0, p + 1, (m,), generator=torch.Generator().manual_seed(1))
    plan = ops.BackwardPlan.build(ids.to(DEV), num_rows=n_rows, zero_pad=True, pad_id=0, pad_row=3,
                                  bag_size=p, lengths=lengths.to(DEV), last_n=4)
    rows = O.row_index(ids, n_rows, 0)
    pos = torch.arange(p).unsqueeze(0)
    hi = lengths.unsqueeze(1)
    keep = (pos < hi) & (pos >= (hi - 4).clamp(min=0)) & (ids != 0) & (rows != 3)
    got_rows = plan.sorted_rows.cpu()
    assert plan.counters.cpu()[0].item() == int(keep.sum())
    assert torch.equal(got_rows[got_rows < n_rows], torch.sort(rows[keep]).values)


def test_plan_kshift_slots():
    ids = seeded_ids(3000, 40)
    k, n_rows = 8, 1000
    plan = ops.BackwardPlan.build(ids.to(DEV), num_rows=n_rows, hash_mode=N.HASH_ROTL_FLOORMOD,
                                  slots_per_id=k)
    want = torch.stack([O.row_index(ids, n_rows, c) for c in range(k)], dim=1).reshape(-1)
    order = torch.argsort(want, stable=True)
    assert torch.equal(plan.sorted_rows.cpu(), want[order])
    assert torch.equal(plan.sorted_slots.cpu(), order)


# --------------------------------------------------- backward: dense gradients ----
def _dense_grad_gpu(ids, grad, n_rows, dim, **plan_kw):
    spg = plan_kw.pop("slots_per_grad_row", 1)
    plan = ops.BackwardPlan.build(ids.to(DEV), num_rows=n_rows, **plan_kw)
    gw = torch.zeros(n_rows, dim, device=DEV)
    ops.bwd_apply(plan, grad.to(DEV), table=gw, update=N.UPD_DENSE_GRAD, slots_per_grad_row=spg)
    return gw


@pytest.mark.parametrize("n,n_rows,dim", [(1, 7, 4), (32, 7, 8), (33, 3, 64), (1025, 11, 64),
                                          (70001, 997, 32), (200000, 50000, 64), (5000, 1, 128),
                                          (3000, 29, 512), (100000, 1000, 96)])
def test_dense_grad_vs_oracle(n, n_rows, dim):
    ids = seeded_ids(n, 41)
    grad = torch.randn(n, dim, generator=torch.Generator().manual_seed(n))
    got = _dense_grad_gpu(ids, grad, n_rows, dim)
    rows = O.row_index(ids, n_rows, 0)
    assert_sums_close(got, *_dense_grad64(rows, grad, n_rows))
    # and the oracle's own fp32 sum (torch index_add_) obeys the same bound -- it is not tighter than ours
    assert_sums_close(O.dense_grad(rows, grad, n_rows), *_dense_grad64(rows, grad, n_rows))
    untouched = torch.ones(n_rows, dtype=torch.bool)
    untouched[O.row_index(ids, n_rows, 0)] = False
    assert got.cpu()[untouched].abs().sum() == 0


def test_dense_grad_hot_rows_multi_level():
    """k-shift collapse: every id negative -> shift c >= 1 lands on <= 2^(c-1) rows."""
    n, k, n_rows, dim = 40000, 8, 100003, 32
    ids = -seeded_ids(n, 42).abs() - 1
    dx = torch.randn(n, dim, generator=torch.Generator().manual_seed(2))
    got = _dense_grad_gpu(ids, dx, n_rows, dim, hash_mode=N.HASH_ROTL_FLOORMOD, slots_per_id=k,
                          slots_per_grad_row=k)
    want = torch.zeros(n_rows, dim, dtype=torch.float64)
    asum = torch.zeros(n_rows, dim, dtype=torch.float64)
    for c in range(k):
        want.index_add_(0, O.row_index(ids, n_rows, c), dx.double())
        asum.index_add_(0, O.row_index(ids, n_rows, c), dx.double().abs())
    # sums over up to n addends: float64 oracle, fp32 summation bound on the collapse rows
    assert_sums_close(got, want, asum)
    assert got.cpu()[n_rows - 1].abs().sum() > 0  # the shift-1 collapse row


def test_dense_grad_is_deterministic():
    ids = torch.randint(0, 50, (100000,), generator=torch.Generator().manual_seed(3))
    grad = torch.randn(100000, 64, generator=torch.Generator().manual_seed(4))
    a = _dense_grad_gpu(ids, grad, 50, 64)
    b = _dense_grad_gpu(ids, grad, 50, 64)
    assert torch.equal(a, b)


def test_dense_grad_integer_exact():
    """With integer-valued gradients every partial sum is exact in fp32: bit-exact check of the
    whole reduction tree (chunk boundaries, record levels, slot order)."""
    n, n_rows, dim = 300000, 37, 64
    ids = torch.randint(0, n_rows, (n,), generator=torch.Generator().manual_seed(5))
    ids[:150000] = 5  # one run spanning thousands of chunks
    grad = torch.randint(-3, 4, (n, dim), generator=torch.Generator().manual_seed(6)).float()
    got = _dense_grad_gpu(ids, grad, n_rows, dim, hash_mode=N.HASH_IDENTITY)
    want = O.dense_grad(ids, grad, n_rows)
    assert torch.equal(got.cpu(), want)


def test_pooled_backward_mean_weights_and_window():
    m, p, n_rows, dim = 3001, 20, 503, 64
    ids = seeded_ids(m * p, 43, (m, p))
    lengths = torch.randint(0, p + 1, (m,), generator=torch.Generator().manual_seed(7))
    w = torch.randn(n_rows, dim)
    go = torch.randn(m, dim, generator=torch.Generator().manual_seed(8))
    mod = R.PooledEmbeddingBag(n_rows, dim, mode="mean", last_n=5, device=DEV)
    mod.load_state_dict({"emb.weight": w})
    out = mod(ids.to(DEV), lengths.to(DEV))
    out.backward(go.to(DEV))
    wr = w.clone().requires_grad_(True)
    rows = O.row_index(ids, n_rows, 0)
    pos = torch.arange(p).unsqueeze(0)
    hi = lengths.unsqueeze(1)
    use = ((pos < hi) & (pos >= (hi - 5).clamp(min=0))).float()
    ref = (wr[rows] * use.unsqueeze(-1)).sum(1) / use.sum(1, keepdim=True).clamp(min=1)
    close(out, ref)
    ref.backward(go)
    torch.testing.assert_close(mod.emb.weight.grad.cpu(), wr.grad, rtol=1e-5, atol=1e-5)


# ----------------------------------------------------- backward: fused updates ----
def test_flat_adagrad_golden_fused_and_torch_modes(golden):
    g = golden("flat_adagrad_train")
    ids, go = T(g["ids"]).to(DEV), T(g["grad_out"]).to(DEV)
    # fused mode
    m = R.FlatEmbedding(500, 32, device=DEV,
                        fused_optimizer=R.FusedOptimizerConfig(kind="adagrad", lr=float(g["lr"])))
    m.load_state_dict({"_emb_table.weight": T(g["weight0"])})
    for _ in range(2):
        m(ids).backward(go)
    close(m._emb_table.weight, T(g["weight2"]))
    close(m._emb_table.opt_state1, T(g["state_sum2"]))
    # torch-compatible mode: dense .grad + torch.optim.Adagrad on the device
    m2 = R.FlatEmbedding(500, 32, device=DEV)
    m2.load_state_dict({"_emb_table.weight": T(g["weight0"])})
    opt = torch.optim.Adagrad(m2.parameters(), lr=float(g["lr"]))
    for _ in range(2):
        opt.zero_grad()
        m2(ids).backward(go)
        opt.step()
    close(m2._emb_table.weight, T(g["weight2"]))
    # rows never indexed stay bit-identical
    rows = O.row_index(T(g["ids"]), 500, 0)
    untouched = torch.ones(500, dtype=torch.bool)
    untouched[rows.reshape(-1)] = False
    assert torch.equal(m._emb_table.weight.cpu()[untouched], T(g["weight0"])[untouched])


def _one_step_from(state_w, state_s, fused, lr, step_fn, n_rows, dim, k, normalize):
    """One optimizer step of a KShiftEmbedding that starts from the given table / Adagrad accumulator."""
    m = R.KShiftEmbedding(n_rows, dim, num_shifts=k, normalize_output=normalize, device=DEV)
    m.load_state_dict({"emb.weight": state_w})
    if fused:
        m.emb.enable_fused_optimizer(kind="adagrad", lr=lr)
        m.emb._ensure_state()
        m.emb.opt_state1.copy_(state_s)
        opt = R.FusedEmbeddingOptimizer([m.emb])
    else:
        opt = torch.optim.Adagrad(m.parameters(), lr=lr)
        opt.state[m.emb.weight]["sum"].copy_(state_s)
    opt.zero_grad()
    loss = step_fn(m)
    loss.backward()
    opt.step()
    return m.emb.weight.detach().cpu(), loss.item()


def test_kshift_train_loop_golden(golden):
    """embedding_module_gen.train_model body (MSE vs target, Adagrad lr 0.5), three steps.
      * whole trajectory vs the reference-generated CPU fixture: losses at 1e-5, table statistically (a
        trajectory of sign-like Adagrad steps amplifies any fp32 difference of an earlier step);
      * EVERY STEP element by element vs the same step in plain torch on this GPU started from the same
        state, within the demonstrated fp32 summation budget (tests/tolerances.py)."""
    g = golden("kshift_adagrad_train")
    k, lr = int(g["k"]), float(g["lr"])
    ids_c, target_c = T(g["ids"]), T(g["target"])
    ids, target = ids_c.to(DEV), target_c.to(DEV)
    n_rows = g["weight0"].shape[0]
    mse = torch.nn.functional.mse_loss
    # (a) whole trajectory vs the CPU fixture
    for fused in (True, False):
        m = R.KShiftEmbedding(n_rows, 32, num_shifts=k, normalize_output=True, device=DEV)
        m.load_state_dict({"emb.weight": T(g["weight0"])})
        if fused:
            m.emb.enable_fused_optimizer(kind="adagrad", lr=lr)
            opt = R.FusedEmbeddingOptimizer([m.emb])
        else:
            opt = torch.optim.Adagrad(m.parameters(), lr=lr)
        losses = []
        for _ in range(3):
            opt.zero_grad()
            loss = mse(m(ids), target)
            loss.backward()
            opt.step()
            losses.append(loss.item())
        np.testing.assert_allclose(losses, g["losses"], rtol=1e-5)
        assert_cross_device_trajectory(m.emb.weight.detach().cpu(), T(g["weight3"]), lr, 3, f"fused={fused} vs the CPU fixture")
    # (b) step by step vs plain torch on this GPU (the oracle's functions are device-agnostic)
    wt = torch.nn.Parameter(T(g["weight0"]).to(DEV))
    opt_t = torch.optim.Adagrad([wt], lr=lr)
    for step in range(3):
        w_before, s_before = wt.detach().clone(), opt_t.state[wt]["sum"].clone()
        opt_t.zero_grad()
        mse(O.kshift_embedding(wt, ids, k, True), target).backward()
        opt_t.step()
        budget = kshift_adagrad_budget(ids_c, target_c, w_before.cpu(), k, lr, steps=1, s0=s_before.cpu())
        for fused in (True, False):
            got, _ = _one_step_from(w_before, s_before, fused, lr, lambda m: mse(m(ids), target), n_rows, 32, k, True)
            _assert_adagrad_trajectory_close(got, wt.detach().cpu(), budget, f"step {step} fused={fused} vs torch on this GPU")


@pytest.mark.parametrize("kind", ["sgd", "adagrad", "rowwise_adagrad", "adam", "adamw"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_optimizers_vs_oracle(kind, dtype):
    bf = dtype == torch.bfloat16
    n, n_rows, dim = 20000, 3001, 64
    torch.manual_seed(11)
    w0 = torch.randn(n_rows, dim).to(dtype)
    wd = 0.01 if kind in ("adamw", "sgd") else 0.0
    # Adagrad from a zero accumulator is lr * g / |g| on the first step: a sign function of sums
    # whose fp32 order differs between implementations.  A non-zero initial accumulator (a
    # torch.optim.Adagrad option) keeps the randomized comparison well conditioned; the zero-start
    # case is pinned by the reference-generated golden loops above.
    acc0 = 0.1 if "adagrad" in kind else 0.0
    cfg = R.FusedOptimizerConfig(kind=kind, lr=0.05, eps=1e-8 if "adam" in kind else 1e-10,
                                 weight_decay=wd, initial_accumulator_value=acc0)
    m = R.FlatEmbedding(n_rows, dim, device=DEV, dtype=dtype, fused_optimizer=cfg)
    m.load_state_dict({"_emb_table.weight": w0})
    w = w0.float().clone()
    s1 = torch.full((n_rows,), acc0) if kind == "rowwise_adagrad" else torch.full((n_rows, dim), acc0)
    s2 = torch.zeros(n_rows, dim)
    for step in range(1, 4):
        ids = seeded_ids(n, 50 + step)
        go = torch.randn(n, dim, generator=torch.Generator().manual_seed(step)).to(dtype)
        m(ids.to(DEV)).backward(go.to(DEV))
        rows = O.row_index(ids, n_rows, 0)
        gw = O.dense_grad(rows, go.float(), n_rows)
        touched = torch.zeros(n_rows, dtype=torch.bool)
        touched[rows] = True
        if bf:
            w = w.bfloat16().float()  # the table is stored in bf16 between steps
        if kind == "sgd":
            g2 = gw.clone()
            g2[touched] += wd * w[touched]
            g2[~touched] = 0
            O.sgd_step(w, g2, cfg.lr)
        elif kind == "adagrad":
            O.adagrad_step(w, gw, s1, cfg.lr, cfg.eps, step=step)
        elif kind == "rowwise_adagrad":
            O.rowwise_adagrad_step(w, gw, touched, s1, cfg.lr, cfg.eps)
        else:
            O.lazy_adam_step(w, gw, touched, s1, s2, cfg.lr, cfg.betas, cfg.eps, wd, step,
                             decoupled=(kind == "adamw"))
    got = m._emb_table.weight.detach().cpu().float()
    # north-star tolerances: 1e-5 (fp32) / 1e-2 (bf16) on updated weights and optimizer state
    torch.testing.assert_close(got, w.bfloat16().float() if bf else w,
                               rtol=RTOL16 if bf else RTOL32, atol=ATOL16 if bf else ATOL32)
    if not bf and kind != "sgd":
        torch.testing.assert_close(m._emb_table.opt_state1.cpu(), s1, rtol=RTOL32, atol=1e-6)


def test_sparse_flag_gives_coo_grad():
    m = R.KShiftEmbedding(97, 8, num_shifts=4, sparse=True, device=DEV)
    ids = seeded_ids(50, 25)
    go = torch.randn(50, 8)
    m(ids.to(DEV)).backward(go.to(DEV))
    g = m.emb.weight.grad
    assert g.is_sparse and g._nnz() == 200
    w = m.emb.weight.detach().cpu().clone().requires_grad_(True)
    O.kshift_embedding(w, ids, 4).backward(go)
    torch.testing.assert_close(g.to_dense().cpu(), w.grad, rtol=1e-5, atol=1e-6)
    f = R.FlatEmbedding(97, 8, sparse=True, device=DEV)
    f(ids.to(DEV)).sum().backward()
    gf = f._emb_table.weight.grad
    assert gf.is_sparse and not gf.is_coalesced() and gf._nnz() == 50


def test_padding_idx_row_never_updated():
    m = R.FlatEmbedding(10, 4, padding_idx=0, device=DEV)
    ids = torch.tensor([[-1, 0, 13, 10, 20]], device=DEV)
    assert O.row_index(ids.cpu(), 10, 0).tolist() == [[9, 0, 3, 0, 0]]
    out = m(ids)
    assert out[0, 1].abs().sum() == 0
    out.sum().backward()
    g = m._emb_table.weight.grad
    assert g[0].abs().sum() == 0 and g[9].sum() == 4 and g[3].sum() == 4


# ----------------------------------------------- fused flip (right-pad -> left-pad) ----
@pytest.mark.parametrize("kind", ["flat", "flat_norm", "kshift", "kshift_sparse"])
def test_flip_sequences_equals_torch_flip(kind):
    b, l, n_rows, dim = 37, 50, 4001, 32
    torch.manual_seed(13)
    w = torch.randn(n_rows, dim)
    ids = seeded_ids(b * l, 90, (b, l))
    ids[:, 35:] = 0
    go = torch.randn(b, l, dim, generator=torch.Generator().manual_seed(14))
    wr = w.clone().requires_grad_(True)
    if kind.startswith("flat"):
        norm = kind == "flat_norm"
        m = R.FlatEmbedding(n_rows, dim, normalize_output=norm, flip_sequences=True, device=DEV)
        m.load_state_dict({"_emb_table.weight": w})
        ref = O.flat_embedding(wr, ids, normalize=norm).flip(1)
        param = m._emb_table.weight
    else:
        m = R.KShiftEmbedding(n_rows, dim, num_shifts=8, normalize_output=True, flip_sequences=True,
                              sparse=kind.endswith("sparse"), device=DEV)
        m.load_state_dict({"emb.weight": w})
        ref = O.kshift_embedding(wr, ids, 8, normalize=True).flip(1)
        param = m.emb.weight
    out = m(ids.to(DEV))
    if kind == "flat":
        assert torch.equal(out.cpu(), ref.detach())  # rows bit-exact, only re-addressed
    else:
        close(out, ref.detach())
    out.backward(go.to(DEV))
    ref.backward(go)
    g = param.grad.to_dense() if param.grad.is_sparse else param.grad
    torch.testing.assert_close(g.cpu(), wr.grad, rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------ table-batched mode ----
def test_table_batched_gather_plan_update_equal_per_table_calls():
    t_tables, n, n_rows, dim = 3, 5000, 1009, 64
    torch.manual_seed(21)
    w = torch.randn(t_tables * n_rows, dim)
    ids = seeded_ids(t_tables * n, 60)
    go = torch.randn(t_tables * n, dim, generator=torch.Generator().manual_seed(9))
    wd = w.to(DEV)
    out, _ = ops.gather_fwd(wd, ids.to(DEV), ids_per_table=n)
    for t in range(t_tables):
        want = O.flat_embedding(w[t * n_rows:(t + 1) * n_rows], ids[t * n:(t + 1) * n])
        assert torch.equal(out[t * n:(t + 1) * n].cpu(), want)
    plan = ops.BackwardPlan.build(ids.to(DEV), num_rows=n_rows, ids_per_table=n)
    assert plan.num_rows == t_tables * n_rows
    rows = torch.cat([O.row_index(ids[t * n:(t + 1) * n], n_rows, 0) + t * n_rows for t in range(t_tables)])
    order = torch.argsort(rows, stable=True)
    assert torch.equal(plan.sorted_rows.cpu(), rows[order]) and torch.equal(plan.sorted_slots.cpu(), order)
    assert plan.counters.cpu().tolist() == [t_tables * n, torch.unique(rows).numel()]
    state = torch.zeros_like(wd)
    ops.bwd_apply(plan, go.to(DEV), table=wd, update=N.UPD_ADAGRAD, state1=state,
                  hp=ops.make_optim_params(lr=0.5, eps=1e-10))
    w_ref, s_ref = w.clone(), torch.zeros_like(w)
    O.adagrad_step(w_ref, O.dense_grad(rows, go, t_tables * n_rows), s_ref, lr=0.5)
    close(wd, w_ref)
    close(state, s_ref)


# ------------------------------------------- full-size properties (BASELINE cfg 2) ----
def test_full_size_gather_and_update_properties():
    b, l, n_rows, dim = 8192, 200, 1_000_000, 64
    n = b * l
    ids = seeded_ids(n, 1000, (b, l)).to(DEV)
    m = R.FlatEmbedding(n_rows, dim, device=DEV, zero_init=True,
                        fused_optimizer=R.FusedOptimizerConfig(kind="sgd", lr=1.0))
    w = m._emb_table.weight
    w.copy_(torch.arange(n_rows, device=DEV, dtype=torch.float32).unsqueeze(1).expand(-1, dim))
    out = m(ids)
    rows = ops.row_index(ids, N.HASH_FLOORMOD, n_rows)
    # every gathered row carries its own row index in all 64 lanes
    assert torch.equal(out[..., 0].long(), rows) and torch.equal(out[..., 63].long(), rows)
    w.zero_()
    out.backward(torch.ones_like(out))
    # linearity: with grad == 1 and SGD lr 1 from zero, w[r] == -count(r) exactly
    counts = torch.bincount(rows.view(-1), minlength=n_rows).float()
    assert torch.equal(w[:, 0], -counts) and torch.equal(w[:, 37], -counts)
    assert float(w[:, 0].double().sum()) == -float(n)  # checksum of checksums


# -------------------- cfg 1 shapes: the LTHM product front-end the embedding path feeds ----
def test_cfg1_lthm_product_front_end():
    """BASELINE cfg 1 (B 256, L 50, KShiftEmbedding(10 000, 32, k = 8), right-padded histories):
    Encoder.forward -> ProductTower front-end restated with the oracle (encoder.py:44-54,
    product_tower.py:47-59): detached k-shift embeddings, pad mask ids == 0, six
    CosineVectorEmbedding bags summed, masked_fill, flip to left-padding; then the
    embedding_module_gen-style table update (fwd + bwd + Adagrad) on the same batch."""
    b, l, n_rows, d_in, d_out, k = 256, 50, 10_000, 32, 64, 8
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(-2 ** 63, 2 ** 63 - 1, (b, l), generator=g, dtype=torch.int64)
    valid = torch.randint(1, l + 1, (b,), generator=g)
    ids[torch.arange(l).unsqueeze(0) >= valid.unsqueeze(1)] = 0  # right padding with token 0
    torch.manual_seed(1234)
    w = torch.randn(n_rows, d_in)
    ks = R.KShiftEmbedding(n_rows, d_in, num_shifts=k, normalize_output=True, flip_sequences=True, device=DEV)
    ks.load_state_dict({"emb.weight": w})
    bins = [2, 4, 8, 12, 16, 20]
    cvs = [R.CosineVectorEmbedding(d_in, d_out, n_proj=32, num_bins=nb, device=DEV) for nb in bins]

    # --- B200 path: flipped gather, bags on the flipped embeddings, mask applied in flipped order
    x = ks(ids.to(DEV)).detach()                                   # [B, L, d_in], already left-padded
    mask = (ids == 0).flip(1).to(DEV)
    emb = sum(cv(x) for cv in cvs).masked_fill(mask.unsqueeze(-1), 0.0)

    # --- oracle restatement of the reference flow (flip last, as Encoder.flip_all does)
    x_ref = O.kshift_embedding(w, ids, k, normalize=True)
    emb_ref = torch.zeros(b, l, d_out)
    agree = 1.0
    for cv in cvs:
        idxs = O.cosine_bucket_indices(x_ref, cv.projection_mat.cpu(), cv.grid.cpu(), cv.pos_offset.cpu())
        got_idx = cv.bucket_indices(x.flip(1)).cpu()
        agree = min(agree, (got_idx == idxs).float().mean().item())
        emb_ref = emb_ref + O.embedding_bag_sum(cv.emb.weight.detach().cpu(), idxs).view(b, l, d_out)
    emb_ref = emb_ref.masked_fill((ids == 0).unsqueeze(-1), 0.0).flip(1)
    close(x, x_ref.flip(1))
    assert agree >= 0.999  # bucket edges: float compare of a GPU vs CPU matmul
    bad = (emb.cpu() - emb_ref).abs().amax(-1) > 1e-4          # positions whose bucket flipped at an edge
    assert bad.float().mean().item() <= 0.01

    # --- table update on the same batch (reconstruction loop body, embedding_module_gen.py:148-153)
    ks2 = R.KShiftEmbedding(n_rows, d_in, num_shifts=k, normalize_output=True, device=DEV,
                            fused_optimizer=R.FusedOptimizerConfig(kind="adagrad", lr=0.5))
    ks2.load_state_dict({"emb.weight": w})
    target = torch.nn.functional.normalize(torch.randn(b, l, d_in, generator=g), dim=-1)
    loss = torch.nn.functional.mse_loss(ks2(ids.to(DEV)), target.to(DEV))
    loss.backward()
    wr = w.clone().requires_grad_(True)
    loss_ref = torch.nn.functional.mse_loss(O.kshift_embedding(wr, ids, k, normalize=True), target)
    loss_ref.backward()
    w_ref, s_ref = w.clone(), torch.zeros_like(w)
    O.adagrad_step(w_ref, wr.grad, s_ref, lr=0.5)
    assert abs(loss.item() - loss_ref.item()) <= 1e-6 * abs(loss_ref.item()) + 1e-9
    # Adagrad from a zero accumulator moves an element by lr * G / |G|: 1e-5 wherever G is well above the
    # fp32 summation noise of its terms, the demonstrated budget (tests/tolerances.py) elsewhere -- a full
    # +-lr sign flip is only admitted where float64 says |G| is below that noise
    budget = kshift_adagrad_budget(ids, target, w, k, 0.5, steps=1)
    assert_cross_device_trajectory(ks2.emb.weight.cpu(), w_ref, 0.5, 1, "cfg1 table update vs the CPU oracle")
    wt = torch.nn.Parameter(w.to(DEV))          # the same step in plain torch on this GPU
    opt_t = torch.optim.Adagrad([wt], lr=0.5)
    torch.nn.functional.mse_loss(O.kshift_embedding(wt, ids.to(DEV), k, normalize=True), target.to(DEV)).backward()
    opt_t.step()
    _assert_adagrad_trajectory_close(ks2.emb.weight.cpu(), wt.detach().cpu(), budget, "cfg1 table update vs torch on this GPU")


@pytest.mark.parametrize("dim,dtype,k", [(64, torch.float32, 8), (32, torch.float32, 16), (128, torch.bfloat16, 4),
                                         (24, torch.float32, 3)])
@pytest.mark.parametrize("update", ["dense_grad", "adagrad", "rowwise_adagrad"])
def test_grad_div_equals_materialised_epilogue_backward(dim, dtype, k, update):
    """The x / sqrt(k) backward folded into the segmented reduction (grad_div: one multiply by the
    rounded reciprocal) equals running epilogue_bwd (true division) first and reducing its fp32 dx
    rows (commons/layers.py:170 autograd): bit-identical when sqrt(k) is a power of two (k = 4, 16),
    within 1 ulp per gradient element otherwise."""
    n, n_rows = 5003, 997
    ids = seeded_ids(n, 123).to(DEV)
    grad = torch.randn(n, dim, generator=torch.Generator().manual_seed(9)).to(dtype).to(DEV)
    plan = ops.BackwardPlan.build(ids, num_rows=n_rows, hash_mode=N.HASH_ROTL_FLOORMOD, slots_per_id=k)
    hp = ops.make_optim_params(lr=0.5, eps=1e-10)
    results = []
    for fused in (False, True):
        torch.manual_seed(5)
        w = (torch.zeros(n_rows, dim) if update == "dense_grad" else torch.randn(n_rows, dim)).to(dtype).to(DEV)
        st = None
        if update == "adagrad":
            st = torch.zeros(n_rows, dim, device=DEV)
        elif update == "rowwise_adagrad":
            st = torch.zeros(n_rows, device=DEV)
        if fused:
            ops.bwd_apply(plan, grad, table=w, update=N.UPDATE_BY_NAME[update], state1=st, slots_per_grad_row=k,
                          hp=hp, grad_div=math.sqrt(k))
        else:
            dx = ops.epilogue_bwd(grad, None, None, N.EPI_RSQRT_K, k)
            ops.bwd_apply(plan, dx, table=w, update=N.UPDATE_BY_NAME[update], state1=st, slots_per_grad_row=k, hp=hp)
        results.append((w, st))
    if math.sqrt(k) != int(math.sqrt(k)) and update != "rowwise_adagrad":
        # reciprocal multiply vs division: <= 1 ulp per TERM; the collapse rows sum thousands of
        # terms of magnitude ~1 that largely cancel, so the bound is absolute (1e-5 of the sum's scale)
        torch.testing.assert_close(results[0][0].float(), results[1][0].float(), rtol=1e-5, atol=2e-4)
        if results[0][1] is not None:   # state = (sum)^2: twice the relative error of a cancelling sum
            torch.testing.assert_close(results[0][1], results[1][1], rtol=1e-4, atol=1e-4)
        return
    if update == "rowwise_adagrad" and (dtype == torch.bfloat16 or math.sqrt(k) != int(math.sqrt(k))):
        # bf16 gradients take the 16-byte-per-lane instantiation: the row's mean of squares is
        # reduced over a different lane grouping than with fp32 dx rows -> equal within rounding
        torch.testing.assert_close(results[0][0].float(), results[1][0].float(), rtol=1e-2, atol=1e-2)
        torch.testing.assert_close(results[0][1], results[1][1], rtol=1e-4, atol=1e-4)
        return
    assert torch.equal(results[0][0], results[1][0])
    if results[0][1] is not None:
        assert torch.equal(results[0][1], results[1][1])


def test_mask_model_train_loop_golden(golden):
    """embedding_module_gen.train_mask_model body (:70-118): nn.Sequential(KShiftEmbedding(N, 4, k = 16),
    MLP(4, 1, [64])) + BCEWithLogits + Adagrad(lr 0.5) over positives and uniform-random int64 negatives,
    three steps.  D = 4: 16-byte rows, one lane per row.  Whole trajectory vs the reference-generated fixture
    (losses 1e-5, table statistically); every step's table update vs the same step in plain torch on this GPU
    from the same state within the fp32 summation budget."""
    import copy
    from tolerances import mask_mlp, mask_step_budget
    g = golden("mask_model_train")
    k, lr = int(g["k"]), float(g["lr"])
    n_rows = g["sd0/0.emb.weight"].shape[0]
    bce = torch.nn.functional.binary_cross_entropy_with_logits

    def batch(step):
        ids = T(g["ids"][step])
        return ids, torch.cat([torch.ones(ids.numel() // 2), torch.zeros(ids.numel() // 2)])

    # (a) whole trajectory vs the CPU fixture
    for fused in (False, True):
        ks = R.KShiftEmbedding(n_rows, 4, num_shifts=k, normalize_output=False, device=DEV)
        model = torch.nn.Sequential(ks, mask_mlp(4).to(DEV))
        model.load_state_dict({n[4:]: T(g[n]) for n in g.files if n.startswith("sd0/")})
        if fused:
            ks.emb.enable_fused_optimizer(kind="adagrad", lr=lr)
            opts = [R.FusedEmbeddingOptimizer([ks.emb]), torch.optim.Adagrad(model[1].parameters(), lr=lr)]
        else:
            opts = [torch.optim.Adagrad(model.parameters(), lr=lr)]
        losses = []
        for step in range(3):
            ids, target = batch(step)
            loss = bce(model(ids.to(DEV)).squeeze(1), target.to(DEV))
            loss.backward()
            for o in opts:
                o.step()
                o.zero_grad()
            losses.append(loss.item())
        np.testing.assert_allclose(losses, g["losses"], rtol=1e-5)
        assert_cross_device_trajectory(ks.emb.weight.detach().cpu(), T(g["sd3/0.emb.weight"]), lr, 3,
                                       f"mask model fused={fused} vs the CPU fixture")
    # (b) step by step vs plain torch on this GPU (k = 16: the 1/sqrt(k) scale is exact, so the forward and the
    # upstream gradient are bit-identical -- what differs is the table reduction and the update)
    wt = torch.nn.Parameter(T(g["sd0/0.emb.weight"]).to(DEV))
    head = mask_mlp(4).to(DEV)
    head.load_state_dict({n[6:]: T(g[n]) for n in g.files if n.startswith("sd0/1.")})
    opt_t = torch.optim.Adagrad([wt, *head.parameters()], lr=lr)
    for step in range(3):
        ids, target = batch(step)
        w_before, s_before, head_before = wt.detach().clone(), opt_t.state[wt]["sum"].clone(), copy.deepcopy(head)
        bce(head(O.kshift_embedding(wt, ids.to(DEV), k, False)).squeeze(1), target.to(DEV)).backward()
        opt_t.step()
        opt_t.zero_grad()
        budget = mask_step_budget(ids, w_before.cpu(), s_before.cpu(), head_before, k, lr)
        for fused in (False, True):
            frozen = copy.deepcopy(head_before)
            got, _ = _one_step_from(w_before, s_before, fused, lr,
                                    lambda m: bce(frozen(m(ids.to(DEV))).squeeze(1), target.to(DEV)), n_rows, 4, k, False)
            _assert_adagrad_trajectory_close(got, wt.detach().cpu(), budget,
                                             f"mask model step {step} fused={fused} vs torch on this GPU")


def test_hand_written_radix_sort_selected_by_environment():
    """RECEMB_PLAN_SORT=own swaps the library sort of the plan for the hand-written LSD radix sort
    (csrc/sort.cu); the choice is read once per process, so the plan tests are re-run in a child."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, RECEMB_PLAN_SORT="own")
    res = subprocess.run([sys.executable, "-m", "pytest", __file__, "-q", "-x", "-k",
                          "test_plan_ or test_dense_grad_integer_exact or test_dense_grad_hot_rows"],
                         env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:]
    assert " passed" in res.stdout
