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


def test_peer_arena_layout_is_pure_host_arithmetic():
    """recemb_peer_arena_layout: regions in order, 256-byte aligned, sized for world x cap entries and
    two [world, bags, dim] row buffers (gathered gradients, partial pools)."""
    lib = N.load()
    a = N.PeerArena()
    assert lib.recemb_peer_arena_layout(8, 245824, 65536, 128, N.BF16, C.byref(a)) == 0
    offs = [a.off_flags, a.off_epoch, a.off_status, a.off_counts, a.off_inbox, a.off_grads, a.off_parts, a.bytes]
    assert offs == sorted(offs) and len(set(offs)) == len(offs)
    assert all(o % 128 == 0 for o in offs) and a.off_inbox % 256 == 0 and a.off_grads % 256 == 0
    assert a.off_epoch - a.off_flags >= 2 * 16 * 8                  # 2 barrier channels x 16 ranks x u64
    assert a.off_grads - a.off_inbox >= 8 * 245824 * 8
    assert a.off_parts - a.off_grads >= 8 * 65536 * 128 * 2
    assert a.bytes - a.off_parts >= 8 * 65536 * 128 * 2
    assert (a.cap, a.bags_total) == (245824, 65536)
    # bad arguments are codes, not crashes
    assert lib.recemb_peer_arena_layout(0, 16, 16, 128, N.BF16, C.byref(a)) == -1
    assert lib.recemb_peer_arena_layout(17, 16, 16, 128, N.BF16, C.byref(a)) == -1
    assert lib.recemb_peer_arena_layout(2, 16, 16, 3, N.F32, C.byref(a)) == -3    # row not a 16-byte multiple
    assert C.sizeof(N.PeerGroupStruct) == 8 + 2 * 16 * 8
    assert int(re.search(r"#define RECEMB_MAX_PEERS (\d+)", HEADER).group(1)) == N.MAX_PEERS


def test_peer_entry_points_validate_before_touching_the_gpu():
    lib = N.load()
    g = N.PeerGroupStruct()
    a = N.PeerArena()
    lib.recemb_peer_arena_layout(2, 64, 16, 64, N.F32, C.byref(a))
    g.world, g.rank = 2, 5                                            # rank outside the group
    assert lib.recemb_peer_barrier(C.byref(g), C.byref(a), 0, 0, None) == -1
    g.rank = 0                                                        # arenas not mapped
    assert lib.recemb_peer_barrier(C.byref(g), C.byref(a), 0, 0, None) == -1
    assert b"arena" in lib.recemb_last_error()
    assert lib.recemb_peer_barrier(C.byref(g), C.byref(a), 7, 0, None) == -1     # channel out of range
    assert lib.recemb_peer_pool_fwd(None, 10, 64, N.F32, None, 4, 2, None, 0, None, N.HASH_FLOORMOD, 0,
                                    N.POOL_SUM, 0, 0, None, None, 0, None) == -1
    assert lib.recemb_peer_allgather_push(C.byref(g), None, 32, 0, 0, None) == -1
    assert lib.recemb_peer_plan(None, None, 10, None, 0, 0, None) == -1


def test_sharded_module_peer_capacity_and_modes():
    """Host logic of the peer exchange that needs no GPU: inbox capacity, forward-mode default."""
    from recommendations_b200.sharded import RowWiseShardedEmbeddingBag, SingleProcess

    class FakeComm(SingleProcess):
        def __init__(self, world, rank):
            super().__init__()
            self.world, self.rank = world, rank

    one = RowWiseShardedEmbeddingBag(1000, 16, exchange="peer", comm=FakeComm(1, 0))
    assert one.peer_forward == "pull" and one.peer_capacity(1000) == 1000
    eight = RowWiseShardedEmbeddingBag(1000, 16, exchange="peer", comm=FakeComm(8, 3))
    assert eight.peer_forward == "push"
    cap = eight.peer_capacity(1_310_720)
    assert cap % 2 == 0 and 1.5 * 1_310_720 / 8 <= cap <= 1.5 * 1_310_720 / 8 + 66
    skew = RowWiseShardedEmbeddingBag(1000, 16, exchange="peer", comm=FakeComm(8, 3), capacity_factor=100.0)
    assert skew.peer_capacity(4096) == 4096                           # never more than a rank can send
    with pytest.raises(ValueError):
        RowWiseShardedEmbeddingBag(1000, 16, exchange="peer", comm=FakeComm(2, 0), peer_forward="sideways")
    with pytest.raises(N.NativeError):
        one.peer_group()                                              # built on the first forward only


def test_round2_entry_points_validate_before_touching_the_gpu():
    """Argument errors of the sequence window, the fused lookup sum, the sequence-mode / table-wise / fused-push
    peer entry points are error codes + a message, not crashes (no GPU involved)."""
    lib = N.load()
    dummy = (C.c_int64 * 64)()
    out2 = (C.c_int32 * 2)()
    ws = (C.c_uint8 * 4096)()
    assert lib.recemb_sequence_window_workspace_bytes(50) >= 51 * 4
    assert lib.recemb_sequence_window(None, 0, 4, 16, 0, 2, 0, ws, 4096, out2, 0, None) == -1          # null data
    assert lib.recemb_sequence_window(dummy, 2, 4, 16, 0, 2, 0, ws, 4096, out2, 0, None) == -1         # bad kind
    assert lib.recemb_sequence_window(dummy, 0, 4, 16, 0, 2, 3, ws, 4096, out2, 0, None) == -1         # bad side
    assert lib.recemb_sequence_window(dummy, 0, 4, 100_000, 0, 2, 0, ws, 4096, out2, 0, None) == -3    # too long
    assert lib.recemb_sequence_window(dummy, 0, 4, 16, 0, 2, 0, ws, 8, out2, 0, None) == -1            # workspace
    terms = (N.GatherTerm * 9)()
    assert lib.recemb_multi_gather_add_fwd(None, terms, 9, 4, 64, N.F32, None, None, dummy, 0, None) == -1   # > 8 terms
    assert lib.recemb_multi_gather_add_fwd(None, terms, 1, 4, 64, N.F32, None, None, dummy, 0, None) == -1   # null table
    assert lib.recemb_multi_gather_add_fwd(None, terms, 0, 4, 3, N.F32, None, None, dummy, 0, None) == -3    # 12-byte row
    assert lib.recemb_multi_gather_add_fwd(None, terms, 0, 4, 64, N.F32, dummy, None, dummy, 0, None) == -1  # mask, no row
    g = N.PeerGroupStruct()
    a = N.PeerArena()
    lib.recemb_peer_arena_layout(2, 64, 16, 64, N.F32, C.byref(a))
    assert a.off_gate > a.off_counts and a.off_inbox - a.off_gate >= 64 + 64 * 4 + 64 * 8
    g.world, g.rank = 2, 0
    hp = N.OptimParams()
    # sequence mode needs an arena laid out with bags_total == cap
    assert lib.recemb_peer_bucket_push_rows(C.byref(g), C.byref(a), dummy, 8, None, N.HASH_FLOORMOD, 100, 0, 0, 0,
                                            dummy, ws, 4096, 0, None) == -1
    assert b"bags_total == cap" in lib.recemb_last_error()
    assert lib.recemb_peer_rows_scatter_push(C.byref(g), C.byref(a), dummy, 8, 64, N.F32, dummy, 0, None) == -1
    assert lib.recemb_peer_pool_push_tablewise(C.byref(g), C.byref(a), 64, N.F32, 5, 0, None) == -1    # 16 % 5 != 0
    for fn in (lib.recemb_peer_bwd_apply_fused, lib.recemb_peer_bwd_apply_fused_tablewise):
        # tables x bags_per_table != bags_total
        assert fn(C.byref(g), C.byref(a), dummy, 512, dummy, 3, 4, 64, N.F32, N.UPD_SGD, dummy, 300, 100, None,
                  C.byref(hp), ws, 4096, 16, 0, None) == -1
    # row-wise Adagrad / SGD on 256- or 512-byte rows only (elementwise Adagrad keeps push -> barrier -> update)
    assert lib.recemb_peer_bwd_apply_fused(C.byref(g), C.byref(a), dummy, 512, dummy, 4, 4, 64, N.F32, N.UPD_ADAGRAD,
                                           dummy, 400, 100, None, C.byref(hp), ws, 4096, 16, 0, None) in (-1, -3)
    assert lib.recemb_peer_signal(C.byref(g), C.byref(a), 9, 0, None) == -1                           # channel


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of the header's structs (recemb_layout, recemb_optim_params, recemb_peer_group,
    recemb_peer_arena, recemb_gather_term) have the size and field offsets a C compiler gives the header."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    structs = {
        "recemb_layout": N.Layout, "recemb_optim_params": N.OptimParams, "recemb_peer_group": N.PeerGroupStruct,
        "recemb_peer_arena": N.PeerArena, "recemb_gather_term": N.GatherTerm,
    }
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "recemb_b200.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c11", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, what, val = line.split()
        cls = structs[cname]
        want = C.sizeof(cls) if what == "size" else getattr(cls, what).offset
        assert int(val) == want, line
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())


def test_prototype_arity_matches_the_ctypes_signatures():
    """Every RECEMB_API prototype has as many parameters as its ctypes argtypes list, and pointer / scalar
    parameters line up (a pointer in the header is a c_void_p / POINTER / array there)."""
    header = re.sub(r"/\*.*?\*/", " ", HEADER, flags=re.S)          # comments may contain commas
    protos = re.findall(r"RECEMB_API\s+[\w\s\*]+?\b(recemb_\w+)\s*\(([^;]*?)\)\s*;", header, flags=re.S)
    assert len(protos) == len(N.SIGNATURES)
    for name, params in protos:
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        _, argtypes = N.SIGNATURES[name]
        assert len(plist) == len(argtypes), (name, len(plist), len(argtypes))
        for p, t in zip(plist, argtypes):
            is_ptr_c = "*" in p or "[" in p or "recemb_stream_t" in p
            is_ptr_py = t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or hasattr(t, "_length_")
            assert is_ptr_c == is_ptr_py, (name, p, t)
