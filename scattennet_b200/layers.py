"""``LearningPositionEmbedding``, ``FeedForward`` and ``CoordinateMapping`` with
the interface of the reference ``model/layers.py`` (the three classes the live
model uses; the unused positional encodings are out of scope)."""

from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib as L
from . import functional as F_
from .functional import Act


class LearningPositionEmbedding(nn.Embedding):
    """Learned table with ``offset = 2`` rows of slack: ``x + table[t + 2]``
    (reference ``model/layers.py:15-30``).  On the hot path the containers fuse
    this with the LayerNorm that follows; the standalone ``forward`` is kept
    for interface parity and runs the add as a row-wise kernel."""

    def __init__(self, num_embeddings, embedding_dim):
        self.offset = 2
        super().__init__(num_embeddings + self.offset, embedding_dim)

    def forward(self, inputs_embeds):
        F_.require_cuda(inputs_embeds, self.weight)
        bsz, seq_len = inputs_embeds.shape[:2]
        if seq_len + self.offset > self.weight.shape[0]:
            raise IndexError("index out of range in self")
        prec = F_.get_precision("fp32")
        x = F_.as_f32_2d(inputs_embeds)
        pos = self.weight[self.offset : self.offset + seq_len].detach().float()
        pos = pos.unsqueeze(0).expand(bsz, seq_len, -1).reshape(bsz * seq_len, -1).contiguous()
        out = F_.rowwise(prec, x, F_.make_epilogue(residual_mode=L.RES_AFTER_LN), residual=pos)
        return out.f32.view_as(inputs_embeds).to(inputs_embeds.dtype)


class FeedForward(nn.Module):
    """``fc2(GELU(fc1 x))`` (exact-erf GELU; reference ``model/layers.py:94-108``)."""

    def __init__(self, in_dim, out_dim, dropout):
        super().__init__()
        self.fc1 = nn.Linear(in_dim, out_dim)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(out_dim, in_dim)
        self.dropout = dropout
        self.precision: Optional[str] = None

    def forward(self, x):
        if self.training and self.dropout > 0:
            raise RuntimeError("scattennet_b200 is inference-only: call .eval()")
        F_.require_cuda(x)
        prec = F_.get_precision(self.precision)
        h = F_.linear(prec, [Act.from_f32(x)], [F_.pack_of(self, "fc1", [self.fc1])], F_.make_epilogue(act_pre=L.ACT_GELU),
                      out_f32=not prec.uses_planes)
        y = F_.linear(prec, h, [F_.pack_of(self, "fc2", [self.fc2])], F_.make_epilogue(), out_planes=False)
        return y[0].f32.view_as(x).to(x.dtype)


class CoordinateMapping(nn.Module):
    """Two ``Linear(K_s -> D)`` on the x and y coordinates (reference
    ``model/layers.py:111-123``).  ``KeypointModule`` fuses this into the front-end
    kernel; the standalone ``forward`` pads K_s up to the GEMM granularity."""

    def __init__(self, in_feat, out_feat):
        super().__init__()
        self.mapping_x = nn.Linear(in_feat, out_feat)
        self.mapping_y = nn.Linear(in_feat, out_feat)

    def _one(self, lin: nn.Linear, coord: torch.Tensor) -> torch.Tensor:
        F_.require_cuda(coord, lin.weight)
        prec = F_.get_precision("fp32")
        k = coord.shape[-1]
        kp = (k + 15) // 16 * 16
        x = torch.zeros(coord.numel() // k, kp, dtype=torch.float32, device=coord.device)
        x[:, :k] = coord.reshape(-1, k)
        padded = nn.Linear(kp, lin.out_features).to(coord.device)
        with torch.no_grad():
            padded.weight.zero_()
            padded.weight[:, :k] = lin.weight
            padded.bias.copy_(lin.bias)
        y = F_.linear(prec, [Act(x)], [F_.PackedLinear([padded], None, None)], F_.make_epilogue(), out_planes=False)
        return y[0].f32.view(*coord.shape[:-1], lin.out_features).to(coord.dtype)

    def forward(self, x_coord, y_coord):
        return self._one(self.mapping_x, x_coord), self._one(self.mapping_y, y_coord)
