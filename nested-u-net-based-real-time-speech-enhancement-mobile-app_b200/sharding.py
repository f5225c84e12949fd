"""Multi-GPU plumbing of the path: clips and streams are independent, so ranks take contiguous
ranges and the only collective is one broadcast of the packed weight blob at start-up.  Works on any
`torch.distributed` backend (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import hashlib
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) of `n_items` clips / streams owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(n_items, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def broadcast_blob(blob: Optional[bytes], src: int = 0, device: Optional[torch.device] = None) -> bytes:
    """Rank `src` passes its packed weight blob, every rank returns identical bytes (length, payload and a SHA-256
    check travel in three broadcasts at init; steady state has no collective)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        if blob is None:
            raise ValueError("single process: the blob must be given")
        return blob
    rank = dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    n = torch.tensor([len(blob) if rank == src else 0], dtype=torch.int64, device=dev)
    dist.broadcast(n, src=src)
    if rank == src:
        payload = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        digest = torch.frombuffer(bytearray(hashlib.sha256(blob).digest()), dtype=torch.uint8).to(dev)
    else:
        payload = torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
        digest = torch.empty(32, dtype=torch.uint8, device=dev)
    dist.broadcast(payload, src=src)
    dist.broadcast(digest, src=src)
    out = bytes(payload.cpu().numpy().tobytes())
    if hashlib.sha256(out).digest() != bytes(digest.cpu().numpy().tobytes()):
        raise RuntimeError("weight blob corrupted in the broadcast")
    return out
