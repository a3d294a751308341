"""``EncoderLayer`` / ``Encoder`` with the interface of the reference
``model/encoder.py`` (constructor config keys, forward signatures, state-dict
layout).  The reference's live model never instantiates it;
its data flow is the self branch of the SCA, so it runs on the same kernels.
"""

from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import functional as F_
from .attention import SelfAttention, attention_core
from .functional import Act, Precision
from .layers import LearningPositionEmbedding
from .utils import create_attention_mask  # noqa: F401  (interface parity)


class EncoderLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.d_model = config["d_model"]
        self.self_attn = SelfAttention(d_model=self.d_model, num_heads=config["encoder_attention_heads"],
                                       dropout=config["attention_dropout"])
        self.self_attn_layer_norm = nn.LayerNorm(self.d_model)
        self.dropout = config["dropout"]
        self.activation_fn = nn.GELU()
        self.activation_dropout = config["activation_dropout"]
        self.fc1 = nn.Linear(self.d_model, config["encoder_ffn_dim"])
        self.fc2 = nn.Linear(config["encoder_ffn_dim"], self.d_model)
        self.final_layer_norm = nn.LayerNorm(self.d_model)
        self.precision: Optional[str] = None

    def forward(self, hidden_states, attention_mask):
        """``attention_mask``: the expanded additive ``[B,1,T,T]`` mask, as in the reference."""
        if self.training and max(self.dropout, self.activation_dropout) > 0:
            raise RuntimeError("scattennet_b200 is inference-only: call .eval()")
        F_.require_cuda(hidden_states, attention_mask)
        prec = F_.get_precision(self.precision)
        b, t, _ = hidden_states.shape
        add = attention_mask.to(torch.float32).expand(b, 1, t, t).contiguous()
        out = encoder_layer_forward(prec, self, Act.from_f32(hidden_states), b, t, None, add)
        return out.f32.view_as(hidden_states).to(hidden_states.dtype)


def encoder_layer_forward(prec: Precision, m: EncoderLayer, x: Act, B: int, T: int, key_mask, additive=None) -> Act:
    """reference ``model/encoder.py:26-57``."""
    ctx = attention_core(prec, [m.self_attn], [x], None, B, T, T, L.ATTN_SELF, key_mask, additive)
    h = F_.linear(prec, ctx, [F_.pack_of(m.self_attn, "out", [m.self_attn.out_proj])],
                  F_.make_epilogue(residual_mode=L.RES_BEFORE_LN, layer_norm=True), residuals=[x.f32],
                  lns=[m.self_attn_layer_norm])
    f = F_.linear(prec, h, [F_.pack_of(m, "fc1", [m.fc1])], F_.make_epilogue(act_pre=L.ACT_GELU), out_f32=not prec.uses_planes)
    return F_.linear(prec, f, [F_.pack_of(m, "fc2", [m.fc2])], F_.make_epilogue(residual_mode=L.RES_BEFORE_LN, layer_norm=True),
                     residuals=[h[0].f32], lns=[m.final_layer_norm])[0]


class Encoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dropout = config["dropout"]
        self.layerdrop = config["encoder_layerdrop"]
        embed_dim = config["d_model"]
        self.embed_positions = LearningPositionEmbedding(config["max_position_embeddings"], embed_dim)
        self.layers = nn.ModuleList([EncoderLayer(config) for _ in range(config["encoder_layers"])])
        self.layernorm_embedding = nn.LayerNorm(embed_dim)
        self.precision: Optional[str] = None

    def forward(self, x_embed, attention_mask):
        """``attention_mask``: ``[B,T]`` 0/1 key mask (reference ``model/encoder.py:79-92``)."""
        if self.training and max(self.dropout, self.layerdrop) > 0:
            raise RuntimeError("scattennet_b200 is inference-only: call .eval()")
        F_.require_cuda(x_embed, attention_mask)
        prec = F_.get_precision(self.precision)
        b, t, _ = x_embed.shape
        h = F_.posembed_layernorm(prec, x_embed, self.embed_positions.weight, self.layernorm_embedding, b, t)
        km = F_.key_mask_u8(attention_mask)
        for layer in self.layers:
            h = encoder_layer_forward(prec, layer, h, b, t, km)
        return h.f32.view_as(x_embed).to(x_embed.dtype)
