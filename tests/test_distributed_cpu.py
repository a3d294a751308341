"""Batch sharding + logits gather on CPU (gloo, world_size 2): concatenating the
per-rank outputs must equal the unsharded output exactly (SURVEY.md section 4,
'distributed without a cluster')."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from scattennet_b200.distributed import (gather_logits, gather_logits_peer, partition, partition_by_length,
                                        sharded_encoder_forward)


class StubEncoder(torch.nn.Module):
    """Sequence-independent stand-in for MSCAEncoder (CPU): per-frame 'logits' from the keypoints."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.arange(12, dtype=torch.float32).view(3, 4) / 7.0)

    def forward(self, kp, mask):
        b, t = kp.shape[:2]
        feat = kp[:, : (t // 4) * 4, :4, 0].reshape(b, t // 4, 4, 4).mean(2)  # [B, T/4, 4]
        return {"fuse_coord_gloss_logits": feat @ self.w.t() * mask[:, : t // 4, None].to(feat.dtype)}


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, batch, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        kp = torch.rand(batch, 16, 6, 2, generator=g)
        mask = (torch.rand(batch, 16, generator=g) > 0.2).long()
        model = StubEncoder()
        full, local = sharded_encoder_forward(model, kp, mask)
        ref = model(kp, mask)["fuse_coord_gloss_logits"]
        ok = torch.equal(full, ref)
        same_shape = gather_logits(torch.full((2, 3), float(rank))).tolist()
        # host tensors cannot take the NVLink peer-memory route: both ranks must agree on the NCCL / gloo one
        routed = gather_logits_peer(torch.full((2, 3), float(rank))).tolist()
        ret[rank] = (ok and routed == same_shape, same_shape)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [8, 5, 1])
def test_gloo_sharded_forward_equals_unsharded(batch):
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, batch, ret), nprocs=world, join=True)
    for r in range(world):
        ok, gathered = ret[r]
        assert ok, f"rank {r}: gathered logits differ from the unsharded run"
        assert gathered == [[0.0] * 3] * 2 + [[1.0] * 3] * 2


def test_partition_helpers():
    assert partition(8, 4) == [(0, 2), (2, 4), (4, 6), (6, 8)]
    assert partition(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert partition(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    parts = partition_by_length([200, 187, 160, 200, 133, 96, 200, 64], 2)
    assert sorted(sum(parts, [])) == list(range(8))
    loads = [sum([200, 187, 160, 200, 133, 96, 200, 64][i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 64
