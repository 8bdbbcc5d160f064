"""Instance sharding across GPUs and the one collective of the path.

A single VRAM is an order-dependent read-modify-write stream and is never split across GPUs ("replicas only");
the batched workload shards by INSTANCE: rank r of R owns the contiguous block
``[r*N/R, (r+1)*N/R)`` of the N instances, with its VRAM pool, command buffers and hash slots resident only on
its own GPU.  Nothing crosses NVLink while rendering; the only exchange is one all-gather of the per-instance
64-bit VRAM hashes (8 bytes per instance) after the hash kernel (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


def instance_range(rank: int, world: int, n_total: int) -> tuple[int, int]:
    """Static block partition: [first, last) of the instances owned by ``rank``."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    first = (n_total * rank) // world
    last = (n_total * (rank + 1)) // world
    return first, last


def owner_of(instance: int, world: int, n_total: int) -> int:
    """Inverse of :func:`instance_range`."""
    if not (0 <= instance < n_total):
        raise ValueError("instance out of range")
    r = (instance * world) // n_total
    while instance < instance_range(r, world, n_total)[0]:
        r -= 1
    while instance >= instance_range(r, world, n_total)[1]:
        r += 1
    return r


def gather_hashes(local_hashes, world: int, n_total: int, dist=None):
    """All-gathers per-instance hashes (int64 tensor of this rank's block) into a length-``n_total`` tensor ordered by
    global instance index.  Blocks may be ragged (n_total not divisible by world): they are padded to the largest
    block for the collective and trimmed afterwards."""
    import torch

    if world == 1 or dist is None:
        return local_hashes.clone()
    sizes = [instance_range(r, world, n_total)[1] - instance_range(r, world, n_total)[0] for r in range(world)]
    biggest = max(sizes)
    padded = torch.zeros(biggest, dtype=local_hashes.dtype, device=local_hashes.device)
    padded[: local_hashes.numel()] = local_hashes
    out = [torch.zeros_like(padded) for _ in range(world)]
    dist.all_gather(out, padded)
    return torch.cat([o[:n] for o, n in zip(out, sizes)])
