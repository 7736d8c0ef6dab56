#!/usr/bin/env python
"""Probe of the peer-memory plumbing on N GPUs (torchrun): CUDA IPC export / open through the C ABI,
a peer read + the device barrier, and (informational) torch's symmetric-memory rendezvous."""
import os
import sys
import traceback
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recommendations_b200 import ops  # noqa: E402
from recommendations_b200.peer import PeerGroup  # noqa: E402


def main():
    world, rank = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    try:
        table = torch.full((1024, 64), float(rank + 1), device=dev)
        pg = PeerGroup.connect(table, cap=64, bags_total=16)
        # rows r (global) live on rank r % world: pooling ids 0..world-1 must give sum(1..world)
        ids = torch.arange(world, device=dev, dtype=torch.int64).view(1, world)
        for it in range(3):
            ops.peer_barrier(pg)
            out = ops.peer_pool_fwd(pg, ids, num_rows=1024 * world, dim=64, dtype=torch.float32)
            torch.cuda.synchronize()
            want = float(sum(range(1, world + 1))) + it * world
            assert float(out[0, 0]) == want, (float(out[0, 0]), want)
            ops.peer_barrier(pg)          # everybody has read
            table += 1.0                  # owner-side update, visible to the next pull after the barrier
        pg.raise_on_status(synchronize=True)
        print(f"[probe] rank {rank}: cuda-ipc peer read + barrier ok", flush=True)
        dist.barrier()
        pg.close()
    except Exception:
        ok = False
        traceback.print_exc()
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(4096, dtype=torch.uint8, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
        print(f"[probe] rank {rank}: symm_mem ok, ptrs={len(hdl.buffer_ptrs)} multicast={hdl.multicast_ptr != 0}", flush=True)
    except Exception as e:
        print(f"[probe] rank {rank}: symm_mem unavailable: {type(e).__name__}: {e}", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
