"""Oracle vs. the reference classes imported live from /root/reference (build container
only -- skipped on the GPU box, where the committed golden vectors stand in)."""
import sys
from pathlib import Path

import pytest
import torch

from oracle import embedding_oracle as O
from conftest import seeded_ids

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="/root/reference not mounted")


@pytest.fixture(scope="module")
def ref():
    sys.path.insert(0, str(REF))
    import commons.layers as cl
    import commons.transformers.layers as tl
    import commons.feature_utils as fu
    yield cl, tl, fu
    sys.path.remove(str(REF))


@pytest.mark.parametrize("n_rows", [1, 3, 1000, 1 << 20, 1234567, (1 << 34)])
def test_row_index_random(ref, n_rows):
    cl, _, _ = ref
    m = cl.KShiftEmbedding(4, 2, num_shifts=2)
    m._num_embeddings = n_rows
    ids = seeded_ids(5000, 21)
    for c in (0, 1, 5, 15, 17, 40, 63):
        assert torch.equal(O.row_index(ids, n_rows, c), m.get_row_idx(ids, c))


def test_negative_ids_collapse(ref):
    """for id < 0 and shift c >= 1 the row lands in [N - 2^(c-1), N-1]."""
    ids = -seeded_ids(4000, 22).abs() - 1
    n_rows = 1_000_000
    for c in (1, 2, 8, 15):
        r = O.row_index(ids, n_rows, c)
        assert int(r.min()) >= n_rows - 2 ** (c - 1)


@pytest.mark.parametrize("normalize", [False, True])
def test_flat_and_kshift_modules(ref, normalize):
    cl, _, _ = ref
    torch.manual_seed(3)
    ids = seeded_ids(7 * 19, 23, (7, 19))
    fe = cl.FlatEmbedding(777, 24, normalize_output=normalize)
    assert torch.equal(O.flat_embedding(fe._emb_table.weight.detach(), ids, normalize), fe(ids))
    for k in (1, 2, 8, 16):
        ks = cl.KShiftEmbedding(777, 24, num_shifts=k, normalize_output=normalize)
        assert torch.equal(O.kshift_embedding(ks.emb.weight.detach(), ids, k, normalize), ks(ids))


def test_kshift_backward_matches_autograd_of_reference(ref):
    cl, _, _ = ref
    torch.manual_seed(4)
    ids = seeded_ids(300, 24)
    ks = cl.KShiftEmbedding(97, 8, num_shifts=8, normalize_output=True)
    go = torch.randn(300, 8)
    ks(ids).backward(go)
    w = ks.emb.weight.detach().clone().requires_grad_(True)
    O.kshift_embedding(w, ids, 8, True).backward(go)
    assert torch.equal(w.grad, ks.emb.weight.grad)


def test_sparse_grad_semantics(ref):
    """sparse=True (commons/layers.py:146): one lookup gives an uncoalesced COO with nnz == n;
    the k-shift sum of k such grads is whatever autograd's sparse add yields -- only the dense
    value is contractual, and it equals the dense-mode gradient."""
    cl, _, _ = ref
    ids = seeded_ids(50, 25)
    one = cl.KShiftEmbedding(97, 8, num_shifts=1, sparse=True)
    one(ids).sum().backward()
    g1 = one.emb.weight.grad
    assert g1.is_sparse and not g1.is_coalesced() and g1._nnz() == 50
    ks = cl.KShiftEmbedding(97, 8, num_shifts=4, sparse=True)
    go = torch.randn(50, 8)
    ks(ids).backward(go)
    w = ks.emb.weight.detach().clone().requires_grad_(True)
    O.kshift_embedding(w, ids, 4).backward(go)
    assert ks.emb.weight.grad.is_sparse
    torch.testing.assert_close(ks.emb.weight.grad.to_dense(), w.grad, rtol=1e-6, atol=1e-6)


def test_cosine_vector_embedding(ref):
    _, tl, _ = ref
    torch.manual_seed(5)
    cv = tl.CosineVectorEmbedding(32, 48, n_proj=16, num_bins=20)
    x = torch.randn(4, 9, 32)
    idxs = O.cosine_bucket_indices(x, cv.projection_mat, cv.grid, cv.pos_offset)
    assert torch.equal(O.embedding_bag_sum(cv.emb.weight.detach(), idxs).view(4, 9, 48), cv(x))


def test_feature_utils(ref):
    _, _, fu = ref
    seed = fu.hash_feature_name_to_int("Product_ID")
    assert seed == O.hash_feature_name("Product_ID")
    for s in ("x", "Hello", "42", 42, ""):
        for lower in (False, True):
            assert fu.hash_string_to_long(s, seed, lower) == O.hash_string_to_id(s, seed, lower)
    for arr, size in (([1, 2, 3], 6), ([], 3), (list(range(9)), 4)):
        assert fu.pad_array(arr, size).tolist() == O.pad_history(arr, size).tolist()
