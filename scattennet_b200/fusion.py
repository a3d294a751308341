"""``CoordinatesFusion`` / ``InvertedResidual`` with the interface of the
reference ``model/fusion.py``: three squeeze linears + GELU, an unmasked,
unscaled single-head attention (queries = right, keys = left, values = body),
``out_proj`` + LayerNorm, then ``LN(GELU(W1 x) + x) -> W3 GELU(W2 .)``."""

from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib as L
from . import functional as F_
from .functional import Act, Precision


class InvertedResidual(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, in_dim)
        self.linear_2 = nn.Linear(in_dim, in_dim * 3)
        self.linear_3 = nn.Linear(in_dim * 3, out_dim)
        self.gelu = nn.GELU()
        self.bn1 = nn.LayerNorm(in_dim)  # a LayerNorm despite the name (reference model/fusion.py:65)
        self.precision: Optional[str] = None

    def forward(self, x):
        F_.require_cuda(x)
        prec = F_.get_precision(self.precision)
        out = inverted_residual_forward(prec, self, Act.from_f32(x))
        return out.f32.view(*x.shape[:-1], -1).to(x.dtype)


def inverted_residual_forward(prec: Precision, m: InvertedResidual, x: Act, out_planes: bool = False) -> Act:
    h = F_.linear(prec, [x], [F_.pack_of(m, "linear_1", [m.linear_1])],
                  F_.make_epilogue(act_pre=L.ACT_GELU, residual_mode=L.RES_BEFORE_LN, layer_norm=True), residuals=[x.f32],
                  lns=[m.bn1], out_f32=not prec.uses_planes)
    u = F_.linear(prec, h, [F_.pack_of(m, "linear_2", [m.linear_2])], F_.make_epilogue(act_pre=L.ACT_GELU),
                  out_f32=not prec.uses_planes)
    return F_.linear(prec, u, [F_.pack_of(m, "linear_3", [m.linear_3])], F_.make_epilogue(), out_planes=out_planes)[0]


class CoordinatesFusion(nn.Module):
    def __init__(self, in_feat, out_feat, drop_rate=0.0):
        super().__init__()
        self.left_se = nn.Linear(in_feat, out_feat)
        self.right_se = nn.Linear(in_feat, out_feat)
        self.body_se = nn.Linear(in_feat, out_feat)
        self.out_proj = nn.Linear(out_feat, out_feat)
        self.norm = nn.LayerNorm(out_feat)
        self.gelu = nn.GELU()
        self.inverted_res = InvertedResidual(out_feat, out_feat)
        self.drop_rate = drop_rate
        self.precision: Optional[str] = None

    def forward(self, left_embed, right_embed, body_embed):
        if self.training and self.drop_rate > 0:
            raise RuntimeError("scattennet_b200 is inference-only: call .eval()")
        F_.require_cuda(left_embed, right_embed, body_embed)
        prec = F_.get_precision(self.precision)
        b, t, _ = left_embed.shape
        out = coordinates_fusion_forward(prec, self, Act.from_f32(left_embed), Act.from_f32(right_embed),
                                         Act.from_f32(body_embed), b, t)
        return out.f32.view(b, t, -1).to(left_embed.dtype)


def coordinates_fusion_forward(prec: Precision, m: CoordinatesFusion, left: Act, right: Act, body: Act, B: int, T: int,
                               out_planes: bool = False) -> Act:
    packs = [F_.pack_of(m, "left_se", [m.left_se]), F_.pack_of(m, "right_se", [m.right_se]),
             F_.pack_of(m, "body_se", [m.body_se])]
    if F_.fusion_attention_planes_supported(prec, T, m.out_proj.in_features):
        # tensor-core contraction: the squeeze GEMMs write the split planes the attention kernel's TMA loads consume
        se = F_.linear(prec, [left, right, body], packs, F_.make_epilogue(act_pre=L.ACT_GELU), out_f32=False, out_planes=True)
        l, r, bd = (s.planes for s in se)
        a = F_.fusion_attention_planes(prec, r, l, bd, B, T)
    else:
        se = F_.linear(prec, [left, right, body], packs, F_.make_epilogue(act_pre=L.ACT_GELU), out_planes=False)
        l, r, bd = (s.f32 for s in se)
        a = F_.fusion_attention(prec, r, l, bd, B, T)
    f = F_.linear(prec, [a], [F_.pack_of(m, "out_proj", [m.out_proj])], F_.make_epilogue(layer_norm=True), lns=[m.norm])[0]
    return inverted_residual_forward(prec, m.inverted_res, f, out_planes)
