"""Peer-memory group for the row-wise sharded exchange (include/recemb_b200.h, "peer" section).

One process per GPU.  Every rank owns a table shard and an exchange arena (barrier flags, inbox
for routed backward entries, gather buffer for the pooled gradients); both allocations are
exported with CUDA IPC, the 64-byte handles travel over torch.distributed (host plumbing, once),
and every rank maps every other rank's memory.  After that the lookup kernels load table rows
and store entries / gradients straight through NVLink -- no collective library call, no host
synchronisation on the data path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from . import _native as N

STATUS_INBOX_OVERFLOW = 1
STATUS_BARRIER_TIMEOUT = 2


def arena_layout(world: int, cap: int, bags_total: int, dim: int, dtype: torch.dtype) -> N.PeerArena:
    out = N.PeerArena()
    N.check(N.load().recemb_peer_arena_layout(world, cap, bags_total, dim, N.dtype_code(dtype), C.byref(out)),
            "recemb_peer_arena_layout")
    return out


def export_handle(t: torch.Tensor):
    """(handle bytes, offset of t inside the exported allocation) for a CUDA tensor."""
    dev = N.require_cuda(t)
    handle = (C.c_uint8 * N.PEER_HANDLE_BYTES)()
    off, size = C.c_int64(0), C.c_int64(0)
    N.check(N.load().recemb_peer_export(t.data_ptr(), handle, C.byref(off), C.byref(size), dev),
            "recemb_peer_export")
    return bytes(handle), int(off.value)


# A CUDA IPC handle names a whole allocation and may be opened once per process: two exported tensors
# that the caching allocator carved out of one segment (e.g. the arenas of two batch shapes) share a
# handle.  Process-wide table: (device, handle) -> [mapped base, reference count].
_OPENED: dict = {}


def _open_shared(handle: bytes, dev: int) -> int:
    key = (dev, handle)
    ent = _OPENED.get(key)
    if ent is None:
        base = C.c_void_p()
        buf = (C.c_uint8 * N.PEER_HANDLE_BYTES).from_buffer_copy(handle)
        N.check(N.load().recemb_peer_open(buf, C.byref(base), dev), "recemb_peer_open")
        ent = _OPENED[key] = [int(base.value), 0]
    ent[1] += 1
    return ent[0]


def _close_shared(handle: bytes, dev: int) -> None:
    key = (dev, handle)
    ent = _OPENED.get(key)
    if ent is None:
        return
    ent[1] -= 1
    if ent[1] <= 0:
        N.load().recemb_peer_close(ent[0], dev)
        del _OPENED[key]


class PeerGroup:
    """Mapped views of every rank's table shard and arena + the arena layout.

    `PeerGroup.connect(...)` is the multi-process constructor; `PeerGroup.local(...)` builds a
    group whose "peers" are allocations on ONE device (single-GPU emulation of W ranks for the
    kernel tests -- the kernels cannot tell the difference, only the barrier must be skipped)."""

    def __init__(self, world: int, rank: int, arenas: Sequence[int], tables: Sequence[int],
                 layout: N.PeerArena, arena: torch.Tensor, device: int, opened: Optional[List[bytes]] = None):
        if not (1 <= world <= N.MAX_PEERS):
            raise N.NativeError(f"peer exchange supports 1..{N.MAX_PEERS} ranks, got {world}")
        self.world, self.rank, self.layout, self.arena, self.device = world, rank, layout, arena, device
        self.struct = N.PeerGroupStruct()
        self.struct.world, self.struct.rank = world, rank
        for i in range(world):
            self.struct.arena[i] = arenas[i]
            self.struct.table[i] = tables[i]
        self._opened = opened or []
        self._status_host = torch.zeros(1, dtype=torch.int32).pin_memory() if torch.cuda.is_available() else None

    # ------------------------------------------------------------ builders ----
    @staticmethod
    def new_arena(layout: N.PeerArena, device) -> torch.Tensor:
        return torch.zeros((int(layout.bytes),), dtype=torch.uint8, device=device)

    @classmethod
    def connect(cls, table: torch.Tensor, *, cap: int, bags_total: int, group=None,
                table_ptrs: Optional[Sequence[int]] = None) -> "PeerGroup":
        """Collective: every rank of `group` calls it with its own shard.  `table_ptrs` (the mapped
        shard pointers of an earlier group over the same table) skips the export / mapping of the
        table: a second batch shape only needs a second arena."""
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        dev = N.require_cuda(table)
        layout = arena_layout(world, cap, bags_total, table.shape[1], table.dtype)
        arena = cls.new_arena(layout, table.device)
        mine = (export_handle(arena), export_handle(table) if table_ptrs is None else (b"", 0))
        everyone: List = [None] * world
        dist.all_gather_object(everyone, mine, group=group)
        mapped = {}            # handle bytes -> mapped base, references held by THIS group
        opened: List[bytes] = []

        def resolve(handle: bytes, offset: int) -> int:
            if handle not in mapped:
                mapped[handle] = _open_shared(handle, dev)
                opened.append(handle)
            return mapped[handle] + offset

        arenas, tables = [], []
        for r, ((ah, ao), (th, to)) in enumerate(everyone):
            if r == rank:
                arenas.append(arena.data_ptr())
                tables.append(table.data_ptr())
            else:
                arenas.append(resolve(ah, ao))
                tables.append(resolve(th, to) if table_ptrs is None else int(table_ptrs[r]))
        self = cls(world, rank, arenas, tables, layout, arena, dev, opened)
        # nobody may touch a peer arena before every rank has zero-filled its own
        torch.cuda.synchronize(dev)
        dist.barrier(group=group)
        return self

    @classmethod
    def local(cls, world: int, rank: int, arenas: Sequence[torch.Tensor], tables: Sequence[torch.Tensor],
              layout: N.PeerArena) -> "PeerGroup":
        dev = N.require_cuda(*arenas, *tables)
        return cls(world, rank, [a.data_ptr() for a in arenas], [t.data_ptr() for t in tables], layout,
                   arenas[rank], dev)

    # --------------------------------------------------------------- views ----
    def grads_view(self, dim: int, dtype: torch.dtype) -> torch.Tensor:
        """[world * bags_total, dim] view of my gather buffer (filled by every rank's push)."""
        rows = self.world * int(self.layout.bags_total)
        nbytes = rows * dim * dtype.itemsize
        off = int(self.layout.off_grads)
        return self.arena[off:off + nbytes].view(dtype).view(rows, dim)

    def parts_view(self, dim: int, dtype: torch.dtype) -> torch.Tensor:
        """[world, bags_total, dim] view of my partial-pool region (slice o written by owner o)."""
        rows = int(self.layout.bags_total)
        nbytes = self.world * rows * dim * dtype.itemsize
        off = int(self.layout.off_parts)
        return self.arena[off:off + nbytes].view(dtype).view(self.world, rows, dim)

    def counts_view(self) -> torch.Tensor:
        off = int(self.layout.off_counts)
        return self.arena[off:off + 8 * self.world].view(torch.int64)

    def inbox_view(self) -> torch.Tensor:
        off = int(self.layout.off_inbox)
        n = self.world * int(self.layout.cap)
        return self.arena[off:off + 8 * n].view(torch.int64).view(self.world, int(self.layout.cap))

    def status_word(self) -> torch.Tensor:
        off = int(self.layout.off_status)
        return self.arena[off:off + 4].view(torch.int32)

    # -------------------------------------------------------------- status ----
    def snapshot_status(self) -> None:
        """Asynchronous copy of the sticky status word to pinned host memory (no sync)."""
        self._status_host.copy_(self.status_word(), non_blocking=True)

    def raise_on_status(self, synchronize: bool = False) -> None:
        if synchronize:
            self.snapshot_status()
            torch.cuda.synchronize(self.device)
        st = int(self._status_host[0])
        if st & STATUS_INBOX_OVERFLOW:
            raise N.NativeError(
                "peer exchange: an owner's inbox overflowed (entries were dropped, the last update is incomplete); "
                "raise capacity_factor -- the ids are more skewed across ranks than the inbox allows")
        if st & STATUS_BARRIER_TIMEOUT:
            raise N.NativeError("peer exchange: a device barrier or gradient gate timed out (a rank did not reach it within RECEMB_PEER_BARRIER_TIMEOUT_S, default 600 s); the update of that step was skipped")

    def table_ptrs(self) -> List[int]:
        return [int(self.struct.table[i] or 0) for i in range(self.world)]

    def close(self) -> None:
        for handle in self._opened:
            _close_shared(handle, self.device)
        self._opened = []
