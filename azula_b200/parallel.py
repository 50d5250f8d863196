r"""Batch-sharded sampling over several GPUs (one process per GPU).

Samples of a batch are independent on the generation path -- nothing in ``Sampler.__call__``,
the denoisers or the backbones mixes batch entries (per-sample norms, per-sample attention) -- so
the path shards over the batch with NO per-step collective.  The only exchange is the one-off
broadcast of the weights from rank 0 at initialisation (NCCL over NVLink on GPUs, gloo on CPU).
The reference has no multi-device support at all.
"""

from __future__ import annotations

__all__ = ["broadcast_parameters", "shard_range", "shard_of"]

import torch
import torch.distributed as dist
import torch.nn as nn

from torch import Tensor


def broadcast_parameters(module: nn.Module, src: int = 0, group=None) -> int:
    r"""Overwrites every parameter and floating-point buffer of :py:`module` with rank
    :py:`src`'s values using ONE flat broadcast per dtype; returns the number of bytes sent."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers() if b.is_floating_point()]
    sent = 0
    for dtype in sorted({t.dtype for t in tensors}, key=str):
        same = [t for t in tensors if t.dtype == dtype]
        flat = torch._utils._flatten_dense_tensors(same)
        dist.broadcast(flat, src=src, group=group)
        for t, f in zip(same, torch._utils._unflatten_dense_tensors(flat, same)):
            t.copy_(f)
        sent += flat.numel() * flat.element_size()
    return sent


def shard_range(global_batch: int, rank: int, world: int) -> range:
    r"""Contiguous slice of the global batch owned by :py:`rank` (equal slices required)."""
    if global_batch % world:
        raise ValueError(f"global batch {global_batch} is not divisible by {world} ranks")
    per = global_batch // world
    return range(rank * per, (rank + 1) * per)


def shard_of(x: Tensor, rank: int, world: int) -> Tensor:
    r"""Rank :py:`rank`'s slice of a global batch tensor."""
    r = shard_range(x.shape[0], rank, world)
    return x[r.start : r.stop]
