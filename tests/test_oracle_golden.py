"""Oracle vs. the committed golden vectors (generated FROM THE REFERENCE by
oracle/make_golden.py).  CPU only."""
import numpy as np
import torch

from oracle import embedding_oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_row_index_known_answers(golden):
    g = golden("row_index")
    ids = T(g["ids"])
    for si, n_rows in enumerate(g["sizes"].tolist()):
        for ci, c in enumerate(g["cols"].tolist()):
            assert torch.equal(O.row_index(ids, n_rows, c), T(g["rows"][si, ci])), (n_rows, c)


def test_row_index_numpy_restatement_agrees(golden):
    g = golden("row_index")
    ids = g["ids"][:64]
    for si, n_rows in enumerate(g["sizes"].tolist()):
        for ci, c in enumerate(g["cols"].tolist()):
            assert np.array_equal(O.row_index_np(ids, n_rows, c), g["rows"][si, ci][:64]), (n_rows, c)


def test_survey_known_answers():
    # KShiftEmbedding(1000, 8, 4).get_row_idx
    ids = torch.tensor([0, 1, -1, 2 ** 62, -2 ** 63, 2 ** 63 - 1, 12345678901234, -987654321])
    want = {0: [0, 1, 999, 904, 192, 807, 234, 679], 1: [0, 2, 999, 192, 999, 998, 468, 999],
            2: [0, 4, 999, 1, 998, 997, 936, 999], 3: [0, 8, 999, 2, 996, 995, 872, 999]}
    for c, rows in want.items():
        assert O.row_index(ids, 1000, c).tolist() == rows
        assert O.row_index_np(ids.numpy(), 1000, c).tolist() == rows
    assert O.hash_feature_name("product_id") == 396283771
    assert O.hash_string_to_id("12345", 396283771, False) == -7448648811083631205
    assert O.pad_history([5, 6, 7], 5).tolist() == [5, 6, 7, 0, 0]
    assert O.pad_history(range(10), 4).tolist() == [0, 1, 2, 3]


def test_flat_embedding(golden):
    g = golden("flat_embedding")
    w, ids = T(g["weight"]), T(g["ids"])
    assert torch.equal(O.flat_embedding(w, ids, padding_idx=0), T(g["out"]))
    assert torch.equal(O.flat_embedding(w, ids, normalize=True), T(g["out_norm"]))
    assert torch.count_nonzero(w[0]) == 0  # padding row zeroed by nn.Embedding


def test_kshift_embedding(golden):
    g = golden("kshift_embedding")
    w, ids = T(g["weight"]), T(g["ids"])
    for k, norm in ((4, False), (8, False), (16, True), (16, False)):
        want = T(g[f"out_k{k}_{'norm' if norm else 'scale'}"])
        assert torch.equal(O.kshift_embedding(w, ids, k, normalize=norm), want), (k, norm)


def test_qr_embedding(golden):
    g = golden("qr_embedding")
    ids = T(g["ids"])
    n = int(g["num_embeddings"])
    assert torch.equal(O.qr_embedding(T(g["weight_q"]), T(g["weight_r"]), ids, n, True), T(g["out_norm"]))
    assert torch.equal(O.qr_embedding(T(g["weight_q"]), T(g["weight_r"]), ids, n, False), T(g["out_plain"]))


def test_cosine_vector_embedding(golden):
    g = golden("cosine_vector_embedding")
    idxs = O.cosine_bucket_indices(T(g["x"]), T(g["projection_mat"]), T(g["grid"]), T(g["pos_offset"]))
    assert torch.equal(idxs, T(g["idxs"]))
    out = O.embedding_bag_sum(T(g["weight"]), idxs)
    assert torch.equal(out.view(g["out"].shape), T(g["out"]))
    # the build-defined pooled bag is the same in-order fp32 sum
    pooled = O.pooled_bag(T(g["weight"]), idxs, hash_ids=False)
    assert torch.equal(pooled.view(g["out"].shape), T(g["out"]))
    # backward: dense gradient of the bag sum
    go = T(g["grad_out"]).reshape(-1, g["out"].shape[-1])
    gw = O.dense_grad(idxs, go.unsqueeze(1).expand(-1, idxs.shape[1], -1), g["weight"].shape[0])
    torch.testing.assert_close(gw, T(g["grad_weight"]), rtol=1e-5, atol=1e-6)


def test_flat_adagrad_train(golden):
    g = golden("flat_adagrad_train")
    w = T(g["weight0"]).clone()
    ids, go = T(g["ids"]), T(g["grad_out"])
    state = torch.zeros_like(w)
    rows = O.row_index(ids, w.shape[0], 0)
    for step in range(2):
        gw = O.dense_grad(rows, go, w.shape[0])
        O.adagrad_step(w, gw, state, lr=float(g["lr"]), step=step + 1)
    torch.testing.assert_close(w, T(g["weight2"]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(state, T(g["state_sum2"]), rtol=1e-5, atol=1e-6)


def test_kshift_adagrad_train(golden):
    """embedding_module_gen.train_model loop body restated with the oracle pieces."""
    g = golden("kshift_adagrad_train")
    k, lr = int(g["k"]), float(g["lr"])
    w = T(g["weight0"]).clone().requires_grad_(True)
    ids, target = T(g["ids"]), T(g["target"])
    state = torch.zeros_like(w)
    losses = []
    for step in range(3):
        y = O.kshift_embedding(w, ids, k, normalize=True)
        loss = torch.nn.functional.mse_loss(y, target)
        (gw,) = torch.autograd.grad(loss, w)
        with torch.no_grad():
            O.adagrad_step(w, gw, state, lr=lr, step=step + 1)
        losses.append(loss.item())
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-5)
    torch.testing.assert_close(w.detach(), T(g["weight3"]), rtol=1e-5, atol=1e-6)


def test_feature_utils(golden):
    g = golden("feature_utils")
    seed = int(g["seed"])
    assert seed == O.hash_feature_name("product_id")
    for s, want, want_l in zip(g["strings"].tolist(), g["ids"].tolist(), g["ids_lower"].tolist()):
        assert O.hash_string_to_id(s, seed, False) == want
        assert O.hash_string_to_id(s, seed, True) == want_l
    assert O.pad_history([5, 6, 7], 5).tolist() == g["pad_short"].tolist()
    assert O.pad_history(list(range(10)), 4).tolist() == g["pad_long"].tolist()


def test_pooled_bag_variants_against_embedding_bag():
    torch.manual_seed(0)
    w = torch.randn(50, 16)
    ids = torch.randint(0, 50, (9, 7))
    lengths = torch.tensor([7, 0, 1, 3, 7, 5, 2, 6, 4])
    # sum with lengths == EmbeddingBag over the ragged offsets form
    flat = torch.cat([ids[b, :lengths[b]] for b in range(9)])
    offsets = torch.cat([torch.zeros(1, dtype=torch.long), lengths.cumsum(0)[:-1]])
    want = torch.nn.functional.embedding_bag(flat, w, offsets, mode="sum")
    assert torch.equal(O.pooled_bag(w, ids, lengths=lengths, hash_ids=False), want)
    want_mean = torch.nn.functional.embedding_bag(flat, w, offsets, mode="mean")
    torch.testing.assert_close(O.pooled_bag(w, ids, lengths=lengths, mode="mean", hash_ids=False), want_mean)
    # last-2 window
    got = O.pooled_bag(w, ids, lengths=lengths, last_n=2, hash_ids=False)
    for b in range(9):
        lo = max(0, int(lengths[b]) - 2)
        torch.testing.assert_close(got[b], w[ids[b, lo:lengths[b]]].sum(0))
    # padding ids skipped
    ids2 = ids.clone()
    ids2[:, ::2] = 0
    want_pad = torch.nn.functional.embedding_bag(ids2, w, mode="sum", padding_idx=0)
    torch.testing.assert_close(O.pooled_bag(w, ids2, hash_ids=False, skip_pad=True), want_pad)


def test_dot_interaction_restatement():
    torch.manual_seed(0)
    f = torch.randn(5, 27, 16).bfloat16()
    out = O.dot_interaction(f)
    assert out.shape == (5, 351)
    # element (i, j), j < i sits at i*(i-1)/2 + j
    for i, j in ((1, 0), (2, 1), (26, 25), (13, 4)):
        want = (f[:, i].float() * f[:, j].float()).sum(-1)
        torch.testing.assert_close(out[:, i * (i - 1) // 2 + j], want)


def _py_row(x: int, c: int, n: int) -> int:
    """commons/layers.py:174-185 in unbounded Python integers: wrapping << and ARITHMETIC >> on a
    two's-complement int64, then Python's floor modulo (== torch.remainder)."""
    def wrap(v):
        v &= (1 << 64) - 1
        return v - (1 << 64) if v >= (1 << 63) else v
    if c:
        x = wrap(wrap(x << c) | (x >> (64 - c)))        # Python >> on a negative int is arithmetic
    return x % n


def test_row_index_against_big_integer_model():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.integers(-2 ** 63, 2 ** 63 - 1), min_size=1, max_size=16),
           st.integers(0, 63), st.integers(1, 2 ** 40))
    def check(ids, c, n):
        want = [_py_row(x, c, n) for x in ids]
        assert O.row_index(torch.tensor(ids, dtype=torch.int64), n, c).tolist() == want
        assert O.row_index_np(np.array(ids, dtype=np.int64), n, c).tolist() == want

    check()
    # the collapse of SURVEY section 0.5: a negative id and shift c >= 1 land in the last 2^(c-1) rows
    for c in (1, 2, 8, 15):
        for x in (-1, -2 ** 63, -987654321, -(2 ** 40) - 7):
            assert _py_row(x, c, 1_000_003) >= 1_000_003 - 2 ** (c - 1)
