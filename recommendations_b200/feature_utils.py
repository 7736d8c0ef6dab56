"""GPU id pipeline mirroring commons/feature_utils.py (the feeder of the embedding path):
xxhash ids and history pad / truncate, bit-exact with the reference's per-value Python loops."""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N

MAX_LONG_VALUE_PLUS_ONE = 2 ** 63          # commons/feature_utils.py:6
CATEGORICAL_VAR_HASH_PAD_TOKEN = 0         # commons/feature_utils.py:8

_P32 = (2654435761, 2246822519, 3266489917, 668265263, 374761393)
_M32 = 0xFFFFFFFF


def _rotl32(x, r):
    return ((x << r) | (x >> (32 - r))) & _M32


def xxh32(data: bytes, seed: int = 0) -> int:
    """XXH32 on the host (only ever hashes a feature name: hash_feature_name_to_int)."""
    p1, p2, p3, p4, p5 = _P32
    n, i = len(data), 0
    if n >= 16:
        v = [(seed + p1 + p2) & _M32, (seed + p2) & _M32, seed & _M32, (seed - p1) & _M32]
        while i + 16 <= n:
            for k in range(4):
                w = int.from_bytes(data[i + 4 * k:i + 4 * k + 4], "little")
                v[k] = (_rotl32((v[k] + w * p2) & _M32, 13) * p1) & _M32
            i += 16
        h = (_rotl32(v[0], 1) + _rotl32(v[1], 7) + _rotl32(v[2], 12) + _rotl32(v[3], 18)) & _M32
    else:
        h = (seed + p5) & _M32
    h = (h + n) & _M32
    while i + 4 <= n:
        h = (h + int.from_bytes(data[i:i + 4], "little") * p3) & _M32
        h = (_rotl32(h, 17) * p4) & _M32
        i += 4
    while i < n:
        h = (h + data[i] * p5) & _M32
        h = (_rotl32(h, 11) * p1) & _M32
        i += 1
    h ^= h >> 15
    h = (h * p2) & _M32
    h ^= h >> 13
    h = (h * p3) & _M32
    h ^= h >> 16
    return h


def hash_feature_name_to_int(feature_name: str) -> int:
    """commons/feature_utils.py:36-37."""
    return xxh32(feature_name.lower().encode("utf-8"), 0)


def pack_strings(values: Iterable, device, lower_non_ascii: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """str(value) UTF-8 bytes back to back + offsets [n + 1] on `device`.  lower_non_ascii: values
    with non-ASCII characters are lower-cased HERE with Python's Unicode str.lower() (what the
    reference calls, commons/feature_utils.py:43-44); the kernel only folds ASCII A-Z."""
    strs = [str(v) for v in values]
    if lower_non_ascii:
        strs = [s if s.isascii() else s.lower() for s in strs]
    enc = [s.encode("utf-8") for s in strs]
    offsets = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in enc], out=offsets[1:])
    blob = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8).copy()
    return torch.from_numpy(blob).to(device), torch.from_numpy(offsets).to(device)


def hash_strings_to_long(values: Sequence, seed: int, value_to_lower: bool, device="cuda") -> torch.Tensor:
    """Vector form of hash_string_to_long (commons/feature_utils.py:40-46): int64 ids on `device`."""
    blob, offsets = pack_strings(values, device, lower_non_ascii=value_to_lower)
    return hash_packed_to_long(blob, offsets, seed, value_to_lower, assume_ascii=True)


def hash_packed_to_long(blob: torch.Tensor, offsets: torch.Tensor, seed: int, value_to_lower: bool,
                        assume_ascii: bool = False) -> torch.Tensor:
    """ids of pre-packed UTF-8 strings.  With value_to_lower the kernel folds ASCII A-Z only, so a
    blob holding non-ASCII bytes is refused (one device reduction + sync) unless the caller states
    that those strings were already lower-cased (`assume_ascii=True`, as hash_strings_to_long does
    after lower-casing them with str.lower() on the host)."""
    dev = N.require_cuda(blob, offsets)
    if value_to_lower and not assume_ascii and bool((blob >= 128).any().item()):
        raise N.NativeError("value_to_lower on a blob with non-ASCII bytes: lower-case those strings on the host "
                            "(str.lower() is Unicode-aware, the kernel folds ASCII only) and pass assume_ascii=True")
    n = offsets.numel() - 1
    out = torch.empty((n,), dtype=torch.int64, device=blob.device)
    N.check(N.load().recemb_xxh64_ids(N.ptr(blob), N.ptr(offsets), n, seed, int(value_to_lower), N.ptr(out), dev,
                                      N.stream_ptr(dev)), "recemb_xxh64_ids")
    return out


def pad_histories(values: torch.Tensor, offsets: torch.Tensor, history_length: int,
                  remove_ids: Optional[torch.Tensor] = None,
                  pad_token: int = CATEGORICAL_VAR_HASH_PAD_TOKEN) -> torch.Tensor:
    """Ragged histories (values[offsets[r]:offsets[r+1]]) -> [rows, history_length] int64: truncated,
    right-padded with 0, optionally without the row's own id (handle_categorical_history_feature,
    commons/feature_utils.py:149-179; pad_array :21-25)."""
    dev = N.require_cuda(values, offsets, remove_ids)
    rows = offsets.numel() - 1
    out = torch.empty((rows, history_length), dtype=torch.int64, device=offsets.device)
    N.check(N.load().recemb_pad_histories(N.ptr(values), N.ptr(offsets), N.ptr(remove_ids), rows, history_length,
                                          pad_token, N.ptr(out), dev, N.stream_ptr(dev)), "recemb_pad_histories")
    return out
