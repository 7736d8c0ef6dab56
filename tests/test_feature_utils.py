"""Id pipeline: host xxh32 seed and GPU xxh64 / history padding vs the reference semantics."""
import numpy as np
import pytest
import torch

from recommendations_b200 import feature_utils as FU
from oracle import embedding_oracle as O


def test_host_xxh32_seed_known_answers(golden):
    assert FU.hash_feature_name_to_int("product_id") == 396283771 == int(golden("feature_utils")["seed"])
    for name in ("Product_ID", "x", "a_very_long_feature_name_over_sixteen_bytes", ""):
        assert FU.hash_feature_name_to_int(name) == O.hash_feature_name(name)


@pytest.mark.gpu
def test_gpu_xxh64_golden_and_random(golden):
    g = golden("feature_utils")
    seed = int(g["seed"])
    strings = g["strings"].tolist()
    assert FU.hash_strings_to_long(strings, seed, False).cpu().tolist() == g["ids"].tolist()
    assert FU.hash_strings_to_long(strings, seed, True).cpu().tolist() == g["ids_lower"].tolist()
    assert FU.hash_strings_to_long(["12345"], 396283771, False).item() == -7448648811083631205
    rng = np.random.default_rng(0)
    alphabet = np.array(list("abcXYZ0123456789-_ /"))
    vals = ["".join(rng.choice(alphabet, size=n)) for n in list(range(0, 80)) + rng.integers(0, 200, 500).tolist()]
    for lower in (False, True):
        got = FU.hash_strings_to_long(vals, seed, lower).cpu().tolist()
        assert got == [O.hash_string_to_id(v, seed, lower) for v in vals]
    assert FU.hash_strings_to_long([123, 4.5], 7, False).cpu().tolist() == \
        [O.hash_string_to_id(123, 7), O.hash_string_to_id(4.5, 7)]


@pytest.mark.gpu
def test_gpu_xxh64_unicode_lower_matches_python_str_lower():
    """hash_string_to_long lower-cases with Python's Unicode str.lower() (commons/feature_utils.py:43-44):
    non-ASCII upper-case text must give the reference id, a pre-packed non-ASCII blob is refused."""
    vals = ["ÀÉÎÕÜ", "Ünïcode-MIXED", "ΑΒΓ δ", "ascii ONLY", "İstanbul", "ǅ", "日本語ABC", ""]
    seed = 396283771
    for lower in (False, True):
        assert FU.hash_strings_to_long(vals, seed, lower).cpu().tolist() == \
            [O.hash_string_to_id(v, seed, lower) for v in vals]
    blob, off = FU.pack_strings(vals, "cuda")
    with pytest.raises(RuntimeError, match="non-ASCII"):
        FU.hash_packed_to_long(blob, off, seed, True)
    assert FU.hash_packed_to_long(blob, off, seed, False).cpu().tolist() == \
        [O.hash_string_to_id(v, seed, False) for v in vals]


@pytest.mark.gpu
def test_gpu_pad_histories_matches_reference_loop():
    rng = np.random.default_rng(1)
    rows, L = 300, 20
    lens = rng.integers(0, 60, rows)
    hist = [rng.integers(-5, 6, n).astype(np.int64) for n in lens]
    target = rng.integers(-5, 6, rows).astype(np.int64)
    offsets = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(lens, out=offsets[1:])
    values = torch.from_numpy(np.concatenate(hist) if offsets[-1] else np.zeros(0, np.int64)).cuda()
    off = torch.from_numpy(offsets).cuda()
    got = FU.pad_histories(values, off, L).cpu().numpy()
    want = np.stack([O.pad_history(h, L) for h in hist])
    assert np.array_equal(got, want)
    got = FU.pad_histories(values, off, L, remove_ids=torch.from_numpy(target).cuda()).cpu().numpy()
    want = np.stack([O.pad_history([v for v in h if v != t], L) for h, t in zip(hist, target)])
    assert np.array_equal(got, want)
