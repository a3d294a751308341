"""Attention primitives with the interface of the reference ``model/attention.py``.

Same constructor arguments, ``forward`` signatures and state-dict layout
(``k_proj, v_proj, q_proj, out_proj`` - the order ``BaseAttention.__init__``
registers them, reference ``model/attention.py:23-26``); the computation runs
in ``libscatt.so``: one grouped projection GEMM (q scaled by ``head_dim**-0.5``
*after* the bias, ``:49``; the cross variant's ``key_value_states / 2`` folded
into ``W_v``, ``:103``), a flash-style score/softmax/PV kernel that never
materialises ``[B,H,T,T]``, and the output projection.

The low-level classes take the reference's already-expanded additive mask
``[B,1,Tq,Tk]`` and add it verbatim; the containers above them
(``SeparativeCoordinateAttention``, ``Encoder``) use the ``[B,T]`` key mask
fast path through :func:`attention_core`.
"""

from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _lib as L
from . import functional as F_
from .functional import Act, Precision


class BaseAttention(nn.Module):
    def __init__(self, d_model, num_heads, dropout=0.0, bias=True):
        super().__init__()
        self.d_model = d_model
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = d_model // num_heads
        if (self.head_dim * num_heads) != self.d_model:
            raise ValueError(
                f"d_model must be divisible by num_heads (got `d_model`: {self.d_model}"
                f" and `num_heads`: {num_heads})."
            )
        self.scaling = self.head_dim**-0.5
        self.k_proj = nn.Linear(d_model, d_model, bias=bias)
        self.v_proj = nn.Linear(d_model, d_model, bias=bias)
        self.q_proj = nn.Linear(d_model, d_model, bias=bias)
        self.out_proj = nn.Linear(d_model, d_model, bias=bias)
        self.precision: Optional[str] = None  # None -> functional.get_precision()

    kind = L.ATTN_SELF

    def _check_eval(self):
        if self.training and self.dropout > 0:
            raise RuntimeError("scattennet_b200 is inference-only: call .eval() (attention dropout > 0 in training mode)")

    def _forward(self, hidden_states, key_value_states, attention_mask):
        self._check_eval()
        F_.require_cuda(hidden_states, key_value_states, attention_mask)
        prec = F_.get_precision(self.precision)
        b, tq, d = hidden_states.shape
        tk = tq if key_value_states is None else key_value_states.size(1)
        xq = Act.from_f32(hidden_states)
        xkv = None if key_value_states is None else Act.from_f32(key_value_states)
        additive = None
        if attention_mask is not None:
            additive = attention_mask.to(torch.float32).expand(b, 1, tq, tk).contiguous()
        ctx = attention_core(prec, [self], [xq], None if xkv is None else [xkv], b, tq, tk, self.kind, None, additive)
        out = F_.linear(prec, ctx, [F_.pack_of(self, "out", [self.out_proj])], F_.make_epilogue(), out_planes=False)
        return out[0].f32.view(b, tq, d).to(hidden_states.dtype)


def attention_core(prec: Precision, mods: Sequence[BaseAttention], xq: List[Act], xkv: Optional[List[Act]], B: int, Tq: int,
                   Tk: int, kind: int, key_mask: Optional[torch.Tensor], additive: Optional[torch.Tensor] = None,
                   kv_views=None, q_planes: Optional[List[Act]] = None) -> List[Act]:
    """Projections + softmax(QK^T + mask) V for a group of same-shaped modules
    (one per anatomical stream).  Returns the per-head context ``[B*Tq, D]``
    *before* ``out_proj``.  ``kv_views`` supplies precomputed ``(k, v)`` fp32
    views (the merge ladder projects K/V of all layers in one GEMM); ``q_planes`` an already projected and scaled q
    (fast path only)."""
    d, h = mods[0].d_model, mods[0].num_heads
    scale = mods[0].scaling
    if (prec.uses_planes and additive is None and d // h == 16 and Tk <= F_.ATTN_PLANES_MAX_T
            and (kv_views is None or isinstance(kv_views[0][0], tuple))):
        # fast path: the projection GEMMs write split planes, the attention kernel TMA-loads them as they are
        ep_q = F_.make_epilogue(scale_cols=d, scale=scale)
        if xkv is None and kv_views is None:
            packs = [F_.pack_of(m, "qkv", [m.q_proj, m.k_proj, m.v_proj]) for m in mods]
            qkv = F_.linear(prec, xq, packs, ep_q, out_f32=False)
            qs = [(t.planes, 0) for t in qkv]
            ks = [(t.planes, d) for t in qkv]
            vs = [(t.planes, 2 * d) for t in qkv]
        else:
            if q_planes is not None:  # already projected and scaled (scatt_attn_out_q: fused into the producing layer's tail)
                q = q_planes
            else:
                packs = [F_.pack_of(m, "q", [m.q_proj]) for m in mods]
                q = F_.linear(prec, xq, packs, ep_q, out_f32=False)
            qs = [(t.planes, 0) for t in q]
            if kv_views is None:
                kv = cross_kv(prec, mods, xkv, planes=True)
                ks = [(t.planes, 0) for t in kv]
                vs = [(t.planes, d) for t in kv]
            else:
                ks, vs = [kv[0] for kv in kv_views], [kv[1] for kv in kv_views]
        return F_.stream_attention_planes(prec, qs, ks, vs, B, Tq, Tk, h, kind, key_mask)
    if q_planes is not None:
        raise ValueError("attention_core: a precomputed q needs the plane path (tensor-core engine, key mask, T within range)")
    if xkv is None and kv_views is None:  # self / causal: one N = 3D GEMM
        packs = [F_.pack_of(m, "qkv", [m.q_proj, m.k_proj, m.v_proj]) for m in mods]
        qkv = F_.linear(prec, xq, packs, F_.make_epilogue(scale_cols=d, scale=scale), out_planes=False)
        qs = [t.f32[:, 0:d] for t in qkv]
        ks = [t.f32[:, d : 2 * d] for t in qkv]
        vs = [t.f32[:, 2 * d : 3 * d] for t in qkv]
    else:
        packs = [F_.pack_of(m, "q", [m.q_proj]) for m in mods]
        q = F_.linear(prec, xq, packs, F_.make_epilogue(scale_cols=d, scale=scale), out_planes=False)
        qs = [t.f32 for t in q]
        if kv_views is None:
            kv = cross_kv(prec, mods, xkv)
            ks = [t.f32[:, 0:d] for t in kv]
            vs = [t.f32[:, d : 2 * d] for t in kv]
        else:
            ks, vs = [kv[0] for kv in kv_views], [kv[1] for kv in kv_views]
    return F_.stream_attention(prec, qs, ks, vs, B, Tq, Tk, h, kind, key_mask, additive)


def cross_kv(prec: Precision, mods: Sequence[BaseAttention], xkv: List[Act], planes: bool = False) -> List[Act]:
    """``[k_proj(x) | v_proj(x / 2)]`` as one N = 2D GEMM; the halving is folded
    into ``W_v`` (exact: a power of two), the bias is not halved (``:103``)."""
    packs = [F_.pack_of(m, "kv", [m.k_proj, m.v_proj], scales=[1.0, 0.5]) for m in mods]
    return F_.linear(prec, xkv, packs, F_.make_epilogue(), out_f32=not planes, out_planes=planes)


class SelfAttention(BaseAttention):
    kind = L.ATTN_SELF

    def forward(self, hidden_states, attention_mask):
        return self._forward(hidden_states, None, attention_mask)


class CrossAttention(BaseAttention):
    kind = L.ATTN_CROSS

    def forward(self, hidden_states, key_value_states, attention_mask):
        return self._forward(hidden_states, key_value_states, attention_mask)


class SelfCausalAttention(BaseAttention):
    kind = L.ATTN_CAUSAL

    def forward(self, hidden_states, attention_mask):
        return self._forward(hidden_states, None, attention_mask)
