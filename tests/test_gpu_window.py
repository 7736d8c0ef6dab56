"""a7 -- QueryTower's batch-wide trim (models/lthm/sequence/query_tower.py:73-86) on the device and the
windowed sequence gather / k-shift that never moves the trimmed columns.

  * recemb_sequence_window vs the reference rule (restated in tests/harness_lthm.py::reference_trim, itself
    checked against the imported QueryTower in tests/test_lthm_step.py) on masks with leading, interior and
    total padding;
  * windowed forward == full forward, flipped, then sliced (bit-exact), both orientations;
  * windowed backward == backward through the slice of the full forward;
  * the pre-trim on `ids == 0` followed by the reference rule on the window == the reference rule on the full
    mask, for masks that contain more padding than the ids show (the norm-threshold part of
    product_tower.py:49)."""
import pytest
import torch

import harness_lthm as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _right_padded_ids(b, length, max_valid, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(-2 ** 62, 2 ** 62, (b, length), generator=g, dtype=torch.int64)
    ids[ids == 0] = 1
    valid = torch.randint(1, max_valid + 1, (b,), generator=g)
    valid[0] = max_valid
    ids[torch.arange(length).unsqueeze(0) >= valid.unsqueeze(1)] = 0
    return ids


@pytest.mark.parametrize("case", ["leading", "interior", "all_pad", "none", "span_floor"])
def test_sequence_trim_equals_the_reference_rule(case):
    from recommendations_b200.sequence import sequence_trim, SequenceWindow
    b, length, span = 37, 50, 3
    g = torch.Generator().manual_seed(3)
    mask = torch.rand(b, length, generator=g) < 0.3
    if case == "leading":
        mask[:, :11] = True
    elif case == "interior":
        mask[:, :4] = True
        mask[:, 20:29] = True
    elif case == "all_pad":
        mask[:] = True
    elif case == "span_floor":           # more all-pad columns than L - span, only some of them leading
        mask[:] = True
        mask[5, 2] = False
    want = H.QueryTowerH.reference_trim(mask, span)
    assert sequence_trim(mask.to(DEV), span) == want
    # the same rule on right-padded data: the mirrored mask, columns dropped at the tail
    w = SequenceWindow.from_mask(mask.flip(1).contiguous().to(DEV), span, drop="tail")
    assert w.trim == want and w.keep == length - want


@pytest.mark.parametrize("flip", [True, False])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_windowed_kshift_equals_the_trimmed_full_lookup(flip, dtype):
    import recommendations_b200 as R
    from recommendations_b200.sequence import SequenceWindow
    b, length, span = 64, 50, 2
    ids = _right_padded_ids(b, length, 41, 11).to(DEV)
    mod = R.KShiftEmbedding(10_000, 32, num_shifts=8, dtype=dtype, device=DEV, flip_sequences=flip)
    full = mod(ids)                                     # [B, L, D], flipped when flip
    w = SequenceWindow.from_ids(ids, span)
    assert w.keep == 41 and w.trim == 9
    got = mod(ids, window=w)
    want = full[:, w.trim:] if flip else full[:, :w.keep]
    assert got.shape == (b, w.keep, 32)
    assert torch.equal(got, want)
    # backward through the window == backward through the slice of the full lookup
    go = torch.randn_like(got)
    gw_win, = torch.autograd.grad(got, mod.emb.weight, go)
    gw_full, = torch.autograd.grad(want, mod.emb.weight, go)
    # same summands per row; the dropped slots move the chunk boundaries of the segmented reduction, so the
    # fp32 sums are re-associated: north-star 1e-5 plus the summation budget ~ eps32 * sum |g_i| of the
    # collapse rows, which add up thousands of cancelling gradient rows (tests/tolerances.py)
    if dtype == torch.float32:
        asum, = torch.autograd.grad(mod(ids)[:, w.trim:] if flip else mod(ids)[:, :w.keep], mod.emb.weight, go.abs())
        bound = 1e-5 * gw_full.abs() + 16 * 2.0 ** -24 * asum
        assert bool(((gw_win - gw_full).abs() <= bound).all())
    else:
        torch.testing.assert_close(gw_win, gw_full, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("drop", ["tail", "head"])
def test_windowed_flat_gather_with_fused_pad_mask(drop):
    import recommendations_b200 as R
    from recommendations_b200.sequence import SequenceWindow
    b, length, span = 33, 70, 5
    ids = _right_padded_ids(b, length, 52, 5)
    if drop == "head":
        ids = ids.flip(1).contiguous()
    ids = ids.to(DEV)
    mod = R.FlatEmbedding(5003, 64, device=DEV, fused_pad_mask=True)
    w = SequenceWindow.from_ids(ids, span, drop=drop)
    assert w.keep == 52
    got = mod(ids, window=w)
    want = w.narrow(mod(ids))
    assert torch.equal(got, want)
    go = torch.randn_like(got)
    gw_win, = torch.autograd.grad(got, mod._emb_table.weight, go)
    gw_full, = torch.autograd.grad(want, mod._emb_table.weight, go)
    torch.testing.assert_close(gw_win, gw_full, rtol=1e-5, atol=1e-5)


def test_windowed_fused_update_equals_the_unwindowed_one():
    """Fused mode: the plan drops the slots outside the window, the update is the one of the sliced output."""
    import recommendations_b200 as R
    from recommendations_b200.sequence import SequenceWindow
    b, length = 48, 40
    ids = _right_padded_ids(b, length, 29, 9).to(DEV)
    cfgs = dict(kind="adagrad", lr=0.1, initial_accumulator_value=0.1)
    m1 = R.KShiftEmbedding(4001, 32, num_shifts=4, device=DEV, fused_optimizer=R.FusedOptimizerConfig(**cfgs))
    m2 = R.KShiftEmbedding(4001, 32, num_shifts=4, device=DEV, fused_optimizer=R.FusedOptimizerConfig(**cfgs))
    m2.load_state_dict(m1.state_dict())
    w = SequenceWindow.from_ids(ids, 1)
    go = torch.randn(b, w.keep, 32, device=DEV)
    m1(ids, window=w).backward(go)
    m2(ids)[:, :w.keep].backward(go)
    # 1e-5 everywhere except the handful of k-shift collapse rows, whose gradient is a cancelling sum of
    # hundreds of rows (re-associated when the dropped slots move the chunk boundaries): those are bounded
    for a, b_ in ((m1.emb.weight, m2.emb.weight), (m1.emb.opt_state1, m2.emb.opt_state1)):
        err = (a - b_).abs()
        bad = err > 1e-5 * b_.abs() + 1e-5
        assert bad.float().mean().item() <= 1e-4
        assert bool((err <= 1e-3 * b_.abs().clamp_min(1.0)).all())


def test_pretrim_on_ids_then_reference_rule_equals_reference_rule():
    """product_tower.py:49: mask = (||x|| < threshold) | (ids == 0) is a superset of ids == 0."""
    from recommendations_b200.sequence import SequenceWindow
    g = torch.Generator().manual_seed(17)
    for trial in range(40):
        b, length = 9, 24
        span = int(torch.randint(1, 6, (1,), generator=g))
        max_valid = int(torch.randint(1, length + 1, (1,), generator=g))
        ids = _right_padded_ids(b, length, max_valid, 100 + trial)
        extra = torch.rand(b, length, generator=g) < float(torch.rand(1, generator=g))   # norm-threshold padding
        if trial % 5 == 0:
            extra[:, int(torch.randint(0, length, (1,), generator=g)):] = True
        mask_full = (ids == 0) | extra
        # the reference: flip, then trim on the full mask
        true_trim = H.QueryTowerH.reference_trim(mask_full.flip(1), span)
        w = SequenceWindow.from_ids(ids.to(DEV), span)             # before the flip: all-pad columns at the tail
        mask_win = w.narrow(mask_full).flip(1)
        rest = H.QueryTowerH.reference_trim(mask_win, span)
        assert w.trim + rest == true_trim, (trial, w.trim, rest, true_trim)


def test_lthm_step_with_the_window_decided_before_the_gather(golden):
    """The cfg 1 harness step with the product lookup windowed BEFORE the rows are moved reproduces the
    reference fixture: same loss, outputs, trimmed ids."""
    g = golden("lthm_step")
    torch.backends.cuda.matmul.allow_tf32 = False
    model = H.model_from_golden(g, H.b200_layers(DEV, pretrim=True), device=DEV)
    res = H.run_step(model, H.batch_from_golden(g, DEV), steps=2)
    H.compare_with_golden(g, model, res, loss_rtol=1e-5, out_atol=2e-5, out_rtol=1e-4, grad_tol=1e-5, w_tol=1e-5)
    assert model.last_window_keep is not None and model.last_window_keep < model.cfg.hist


def test_fused_lookup_sum_equals_the_chain_of_lookups_and_adds():
    """query_tower.py:89-104 as one kernel: bit-exact forward (fp32 adds in the reference's order), gradients of
    the dense term, the pad row and the four tiny tables equal to autograd through the separate modules."""
    import recommendations_b200 as R
    b, length, dim = 48, 37, 64
    g = torch.Generator().manual_seed(2)
    base = torch.randn(b, length, dim, generator=g).to(DEV).requires_grad_(True)
    base2 = base.detach().clone().requires_grad_(True)
    labels = torch.randint(0, 4, (b, length), generator=g).to(DEV)
    ts = torch.randint(1_600_000_000, 1_700_000_000, (b, length), generator=g).to(DEV)
    mask = (torch.rand(b, length, generator=g) < 0.3).to(DEV)
    pad = torch.nn.Parameter(torch.randn(1, 1, dim, generator=g).to(DEV))
    pad2 = torch.nn.Parameter(pad.detach().clone())
    def mods():
        torch.manual_seed(5)
        return (R.FlatEmbedding(4, dim, device=DEV), R.PatternFromTimelocal(3600, 24, dim, device=DEV),
                R.PatternFromTimelocal(3600, 168, dim, device=DEV), R.PatternFromTimelocal(86400, 7, dim, device=DEV))
    m1, m2 = mods(), mods()
    for a, c in zip(m1, m2):
        c.load_state_dict(a.state_dict())
    got = R.fused_lookup_sum(base, [(m1[0], labels), (m1[1], ts), (m1[2], ts), (m1[3], ts)], mask=mask, masked_row=pad)
    x = base2 + m2[0](labels) + m2[1](ts) + m2[2](ts) + m2[3](ts)
    want = torch.where(mask.unsqueeze(-1), pad2.expand(b, length, -1), x)
    assert torch.equal(got, want)
    go = torch.randn(b, length, dim, generator=g).to(DEV)
    got.backward(go)
    want.backward(go)
    assert torch.equal(base.grad, base2.grad)
    torch.testing.assert_close(pad.grad, pad2.grad, rtol=1e-5, atol=1e-5)
    for a, c in zip(m1, m2):
        wa = next(a.parameters()).grad
        wc = next(c.parameters()).grad
        torch.testing.assert_close(wa, wc, rtol=1e-5, atol=1e-4)


def test_lthm_step_with_fused_front_end(golden):
    """cfg 1 with every piece of the B200 sequence front end switched on: the window decided before the product
    lookup, QueryTower's trim through sequence_trim, its input sum through fused_lookup_sum."""
    import recommendations_b200 as R
    g = golden("lthm_step")
    torch.backends.cuda.matmul.allow_tf32 = False
    L = H.b200_layers(DEV, trim_fn=R.sequence_trim, pretrim=True, fused_front_end=True)
    model = H.model_from_golden(g, L, device=DEV)
    res = H.run_step(model, H.batch_from_golden(g, DEV), steps=2)
    H.compare_with_golden(g, model, res, loss_rtol=1e-5, out_atol=2e-5, out_rtol=1e-4, grad_tol=1e-5, w_tol=1e-5)
