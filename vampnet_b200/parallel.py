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


BUCKET_BYTES = 64 << 20


def broadcast_module_weights(modules: Iterable[torch.nn.Module], src: int = 0, bucket_bytes: int = BUCKET_BYTES) -> int:
    """Broadcast every parameter/buffer of `modules` from `src`, in the tensors' OWN dtype.  Tensors of at least
    `bucket_bytes` go out in place (no staging copy); smaller ones are coalesced per dtype into buckets of at most
    `bucket_bytes`, so a 20-layer model is a few dozen large messages and the peak temporary is one bucket, not a
    second copy of the model.  Derived state (packed bf16 weights, codec packs) is dropped afterwards through the
    module's `_invalidate()`.  Returns the number of bytes broadcast.  No-op without a process group."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    total = 0

    def flush(bucket):
        nonlocal total
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.broadcast(flat, src=src)
        off = 0
        for t in bucket:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
        total += flat.numel() * flat.element_size()
        bucket.clear()

    for m in modules:
        pending = {}  # (dtype, device) -> [tensors], bytes
        seen = set()
        for t in [p.data for p in m.parameters()] + [b.data for b in m.buffers()]:
            if t.data_ptr() in seen or t.numel() == 0:
                continue
            seen.add(t.data_ptr())
            nbytes = t.numel() * t.element_size()
            if nbytes >= bucket_bytes and t.is_contiguous():
                dist.broadcast(t, src=src)
                total += nbytes
                continue
            key = (t.dtype, t.device)
            bucket, size = pending.setdefault(key, ([], 0))
            if size + nbytes > bucket_bytes:
                flush(bucket)
                size = 0
            bucket.append(t)
            pending[key] = (bucket, size + nbytes)
        for bucket, _ in pending.values():
            flush(bucket)
        for sub in m.modules():
            if hasattr(sub, "_invalidate"):
                sub._invalidate()  # packed weights must be rebuilt from the new parameters
    return total


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed numbers are reported as the max over ranks."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
