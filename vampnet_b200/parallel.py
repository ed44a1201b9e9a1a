"""Multi-GPU plumbing for the hot path: independent clips shard across ranks with NO data-path collective
("replicas + batch shards").  The only collective is the one-off weight broadcast from rank 0
over NCCL/NVLink at start-up; everything else is local.  Works with the gloo backend on CPU tensors too, which
is how the host logic is tested without GPUs (tests/test_parallel_cpu.py).
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous split of n_items clips: the first (n_items % world) ranks get one extra."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def broadcast_module_weights(modules: Iterable[torch.nn.Module], src: int = 0) -> int:
    """Broadcast every parameter/buffer of `modules` from `src` as ONE flat blob per module (few large messages
    instead of hundreds of small ones).  Returns the number of bytes broadcast.  No-op without a process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    total = 0
    for m in modules:
        tensors = [p.data for p in m.parameters()] + [b.data for b in m.buffers()]
        if not tensors:
            continue
        flat = torch.cat([t.reshape(-1).to(torch.float32) for t in tensors])
        dist.broadcast(flat, src=src)
        off = 0
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
            off += n
        total += flat.numel() * 4
        if hasattr(m, "_invalidate"):
            m._invalidate()  # packed bf16 weights must be rebuilt from the new parameters
    return total


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed numbers are reported as the max over ranks."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
