"""Host-side helpers of the multi-GPU path (one process per GPU, torch.distributed for the plumbing).

The data path shards the node array by sector sub-tree and needs exactly one exchange step: every
rank ends up with the concatenation of all ranks' visible lists.  The library does that natively with
NCCL (`fyx_allgather_visible`); `allgather_varlen` is the same protocol on torch tensors (gloo or nccl)
— counts first, then fixed max-count slots, then packing — used by the CPU tests of the N>1 logic and
usable as a drop-in when the lists already live in torch tensors.
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def allgather_varlen(local: torch.Tensor, group=None) -> Tuple[torch.Tensor, List[int]]:
    """All-gather 1-D tensors of different lengths.  Returns (concatenation in rank order, per-rank counts)."""
    assert local.dim() == 1
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    counts = torch.zeros(world, dtype=torch.int64, device=local.device)
    dist.all_gather_into_tensor(counts, n, group=group)
    counts_l = [int(c) for c in counts.tolist()]
    maxc = max(counts_l) if counts_l else 0
    if maxc == 0:
        return local.new_empty(0), counts_l
    slot = local.new_zeros(maxc)
    slot[: local.numel()] = local
    padded = local.new_empty(world * maxc)
    dist.all_gather_into_tensor(padded, slot, group=group)
    parts = [padded[r * maxc: r * maxc + counts_l[r]] for r in range(world)]
    return torch.cat(parts), counts_l


def broadcast_bytes(data: bytes, src: int = 0, device="cpu", group=None) -> bytes:
    """Broadcast a small byte string (the 128-byte NCCL unique id of fyx_comm_init)."""
    n = len(data)
    t = torch.zeros(n, dtype=torch.uint8, device=device)
    if dist.get_rank(group) == src:
        t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
    dist.broadcast(t, src, group=group)
    return bytes(t.cpu().numpy().tobytes())
