"""Batch sharding across the GPUs of one box and the logits gather.

Sequences are independent (LayerNorm only, attention within a sequence, eval
mode), so the encoder shards along the batch with *zero* data-path
communication; the only collective is one all-gather of the per-frame logits
that CTC decoding consumes.  The reference itself has no
parallelism (its NCCL group is initialised and never used, ``utils.py:237-265``).

Works with ``torch.distributed`` over NCCL (GPU) and gloo (CPU tests).
"""

from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def partition(batch: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous ``[start, end)`` slice of the batch per rank; sizes differ by at most one."""
    base, extra = divmod(batch, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((start, start + n))
        start += n
    return out


def partition_by_length(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-first assignment of sequence indices to ranks balancing the
    sum of valid lengths (ragged batches); returns sorted index lists per rank."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    loads = [0] * world
    buckets: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        buckets[r].append(i)
        loads[r] += int(lengths[i])
    return [sorted(b) for b in buckets]


def gather_logits(local: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather ``local [B_loc, T', V]`` along the batch.  Every rank must hold
    the same ``B_loc`` (pad the last shard); the result is ``[world * B_loc, T', V]``
    in rank order, i.e. the unsharded batch order of :func:`partition`."""
    world = dist.get_world_size(group)
    if world == 1:
        return local
    local = local.detach().contiguous()
    out = torch.empty((world * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    with torch.no_grad():
        dist.all_gather_into_tensor(out, local, group=group)
    return out


def sharded_encoder_forward(model, keypoints: torch.Tensor, mask: torch.Tensor, group=None, head: str = "fuse_coord_gloss_logits"):
    """Run ``model`` on this rank's slice of the global batch and all-gather the
    logits of ``head``.  ``keypoints`` / ``mask`` are the *global* batch (host or
    device); only the local slice is touched."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    batch = keypoints.shape[0]
    per = -(-batch // world)
    lo, hi = min(rank * per, batch), min((rank + 1) * per, batch)
    dev = next(model.parameters()).device
    kp = torch.zeros((per,) + tuple(keypoints.shape[1:]), dtype=torch.float32, device=dev)
    mk = torch.zeros((per, mask.shape[1]), dtype=mask.dtype, device=dev)
    if hi > lo:
        kp[: hi - lo].copy_(keypoints[lo:hi], non_blocking=True)
        mk[: hi - lo].copy_(mask[lo:hi], non_blocking=True)
    out = model(kp, mk)
    full = gather_logits(out[head], group)
    return full[:batch], out
