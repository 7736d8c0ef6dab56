"""The C-ABI library loads, exports every symbol include/recemb_b200.h declares, and its
argument validation answers without a GPU (no compute calls here)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from recommendations_b200 import _native as N

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "recemb_b200.h").read_text()


def declared_symbols():
    return sorted(set(re.findall(r"RECEMB_API\s+[\w\s\*]+?\b(recemb_\w+)\s*\(", HEADER)))


def test_header_declares_what_python_binds():
    assert declared_symbols() == sorted(N.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = N.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.recemb_abi_version() == 1


def test_enum_values_match_header():
    def enum(name):
        return int(re.search(rf"\b{name}\s*=\s*(-?\d+)", HEADER).group(1))
    assert (N.F32, N.BF16) == (enum("RECEMB_F32"), enum("RECEMB_BF16"))
    assert N.HASH_ROTL_FLOORMOD == enum("RECEMB_HASH_ROTL_FLOORMOD")
    assert N.HASH_QR_REMAINDER == enum("RECEMB_HASH_QR_REMAINDER")
    assert N.EPI_RSQRT_K == enum("RECEMB_EPI_RSQRT_K")
    assert N.UPD_ADAMW == enum("RECEMB_UPD_ADAMW")
    assert N.UPD_ROWWISE_ADAGRAD == enum("RECEMB_UPD_ROWWISE_ADAGRAD")
    assert C.sizeof(N.OptimParams) == 32


def test_argument_errors_are_codes_not_crashes():
    lib = N.load()
    # n == 0 is a no-op everywhere
    assert lib.recemb_row_index(None, 0, N.HASH_FLOORMOD, 10, 0, None, 0, None) == 0
    # null pointers with n > 0
    assert lib.recemb_row_index(None, 4, N.HASH_FLOORMOD, 10, 0, None, 0, None) == -1
    assert b"null" in lib.recemb_last_error()
    # row bytes not a multiple of 16
    rc = lib.recemb_gather_fwd(1 << 20, 10, None, 0, 3, N.F32, 1 << 20, 4, None, N.HASH_FLOORMOD, 0, 0,
                               N.EPI_NONE, 0, 0, 1 << 20, None, 0, None)
    assert rc == -3 and b"16 bytes" in lib.recemb_last_error()
    # unknown hash mode
    rc = lib.recemb_gather_fwd(1 << 20, 10, None, 0, 4, N.F32, 1 << 20, 4, None, 99, 0, 0, N.EPI_NONE, 0, 0,
                               1 << 20, None, 0, None)
    assert rc == -1
    # bad k
    rc = lib.recemb_kshift_fwd(1 << 20, 10, 4, N.F32, 1 << 20, 4, 64, N.EPI_RSQRT_K, 0, 1 << 20, None, 0, None)
    assert rc == -1 and b"num_shifts" in lib.recemb_last_error()
    # workspace sizing is pure host arithmetic
    assert lib.recemb_bwd_apply_workspace_bytes(1_638_400, 64) > 2 * (1_638_400 // 64) * 64 * 4
    assert lib.recemb_bwd_apply_workspace_bytes(0, 64) == 256


def test_missing_library_raises(monkeypatch, tmp_path):
    monkeypatch.setattr(N, "_lib", None)
    monkeypatch.setattr(N, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(N.NativeLibraryMissing):
        N.load()
