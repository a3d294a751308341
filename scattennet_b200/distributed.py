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


class PeerGather:
    """All-gather of equally shaped shards over NVLink peer memory: one kernel per call (``scatt_peer_allgather``)
    pushes this rank's shard into every peer's buffer and runs the barrier, instead of NCCL's ring of ``world - 1``
    dependent hops.  Buffers are ``torch.distributed`` symmetric memory (CUDA VMM handles exchanged once, in
    ``__init__`` - a collective call); two of them alternate so a shard never lands in a buffer a peer still reads.

    ``gather(local)`` returns a view ``[world * B_loc, ...]`` of the current buffer, valid until the call after
    the next one.  Construction raises if symmetric memory cannot be set up (callers fall back to
    :func:`gather_logits`, the NCCL path)."""

    def __init__(self, shard_shape, dtype=torch.float32, device=None, group=None):
        import ctypes as C

        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib as L

        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if self.world > 8:
            raise RuntimeError("PeerGather covers the GPUs of one NVSwitch box (<= 8)")
        self.shard_shape = tuple(shard_shape)
        self.dtype = dtype
        esz = torch.empty(0, dtype=dtype).element_size()
        n = 1
        for d in self.shard_shape:
            n *= int(d)
        self.shard_bytes = n * esz
        if self.shard_bytes % 16:
            raise ValueError("PeerGather: the shard must be a multiple of 16 bytes")
        self.buf_bytes = self.world * self.shard_bytes
        pad = 256  # flag pad: uint64[world], kept apart from the data
        total = 2 * self.buf_bytes + pad
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.mem = symm_mem.empty(total, dtype=torch.uint8, device=dev)
        self.mem.zero_()
        self.handle = symm_mem.rendezvous(self.mem, self.group.group_name)
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)  # every pad is zero before anybody signals
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self._bufs = [(C.c_void_p * self.world)(*[p + par * self.buf_bytes for p in ptrs]) for par in range(2)]
        self._flags = (C.c_void_p * self.world)(*[p + 2 * self.buf_bytes for p in ptrs])
        self.counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self.seq = 0
        self._lib = L

    def gather(self, local: torch.Tensor) -> torch.Tensor:
        if tuple(local.shape) != self.shard_shape or local.dtype != self.dtype:
            raise ValueError(f"PeerGather: expected a {self.shard_shape} {self.dtype} shard, got {tuple(local.shape)} {local.dtype}")
        local = local.detach().contiguous()
        self.seq += 1
        par = self.seq & 1
        L = self._lib
        L.check(L.load().scatt_peer_allgather(local.data_ptr(), self.shard_bytes, self._bufs[par], self._flags, self.world, self.rank,
                                              self.counter.data_ptr(), self.seq, torch.cuda.current_stream().cuda_stream),
                "scatt_peer_allgather")
        view = self.mem[par * self.buf_bytes:(par + 1) * self.buf_bytes].view(self.dtype)
        return view.view((self.world * self.shard_shape[0],) + self.shard_shape[1:])


_peer_gathers = {}


def gather_logits_peer(local: torch.Tensor, group=None) -> torch.Tensor:
    """:func:`gather_logits` over NVLink peer memory when it can be set up (one box, CUDA tensors, symmetric
    memory available; ``SCATT_PEER_GATHER=0`` switches it off), else the NCCL all-gather.  The first call for a
    shard shape is collective (it allocates and exchanges the buffers)."""
    import os

    world = dist.get_world_size(group)
    if world == 1:
        return local
    key = (tuple(local.shape), local.dtype, str(local.device), id(group))
    pg = _peer_gathers.get(key)
    if pg is None:
        # Agree on feasibility BEFORE the collective set-up (symmetric-memory allocation, rendezvous and barrier block
        # until every rank takes part): each rank probes locally, the minimum decides, and only then do all ranks
        # construct - a rank that cannot use peer memory sends everybody down the NCCL route instead of hanging them.
        can = bool(local.is_cuda and os.environ.get("SCATT_PEER_GATHER", "1") != "0" and world <= 8
                   and local.numel() * local.element_size() % 16 == 0)
        if can:
            try:
                import importlib

                importlib.import_module("torch.distributed._symmetric_memory")
            except Exception:
                can = False
        ok = torch.tensor([1 if can else 0], device=local.device if local.is_cuda else "cpu")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        pg = False
        if int(ok.item()) == 1:
            try:
                pg = PeerGather(local.shape, local.dtype, local.device, group)
            except Exception as exc:  # set-up failed on this rank (said once, on stderr); the vote below settles the route
                import sys

                print(f"[scattennet_b200] peer-memory gather unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
                pg = False
            ok = torch.tensor([1 if pg else 0], device=local.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                pg = False
        _peer_gathers[key] = pg
    if pg is False:
        return gather_logits(local, group)
    return pg.gather(local)


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
