"""TEST INFRASTRUCTURE - not part of the product path.

CPU restatement of the SCAttenNet encoder forward (inference) as plain
functions over a flat state dict.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
file; ``scattennet_b200`` never does (it fails loudly without its CUDA library).

Why torch-on-CPU and not numpy/C: the reference is pure PyTorch and all of its
arithmetic is ATen ("third-party arithmetic"), so ATen
fp32 on the host is the closest statement of "what the reference computes";
``dtype=torch.float64`` re-evaluates the same algorithm in double precision to
separate this package's error from the reference's own fp32 rounding.

Parity pin: the reference ships no tests or golden vectors (SURVEY.md section
4), so this oracle is pinned against outputs of the *reference itself*,
generated in the build container by ``tests/golden/make_golden.py`` (which
imports ``/root/reference``) and committed under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function here against them, and
``tests/test_oracle_vs_reference.py`` re-checks live whenever
``/root/reference`` is present.

Each function cites the reference lines it restates (paths relative to the
reference repo root).
"""

from __future__ import annotations

import math
from typing import Mapping, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Mapping[str, Tensor]


def _w(sd: SD, name: str, dtype) -> Tensor:
    return sd[name].to(dtype)


def linear(sd: SD, p: str, x: Tensor) -> Tensor:
    """``nn.Linear`` stored at ``p.weight`` / ``p.bias``."""
    return F.linear(x, _w(sd, p + ".weight", x.dtype), _w(sd, p + ".bias", x.dtype))


def layer_norm(sd: SD, p: str, x: Tensor) -> Tensor:
    """``nn.LayerNorm`` defaults: eps 1e-5, biased variance, affine."""
    return F.layer_norm(x, (x.shape[-1],), _w(sd, p + ".weight", x.dtype), _w(sd, p + ".bias", x.dtype), 1e-5)


# --------------------------------------------------------------------------- masks


def key_padding_additive(mask: Tensor, dtype, tgt_len: Optional[int] = None) -> Tensor:
    """model/utils.py:3-12 - ``[B,1,tgt,src]`` additive mask: 0 for a valid key,
    ``finfo(dtype).min`` for a padded key."""
    b, src = mask.shape
    tgt = src if tgt_len is None else tgt_len
    keep = mask.to(torch.bool)[:, None, None, :].expand(b, 1, tgt, src)
    out = torch.zeros(b, 1, tgt, src, dtype=dtype)
    return out.masked_fill(~keep, torch.finfo(dtype).min)


def causal_additive(mask: Tensor, t: int, dtype) -> Tensor:
    """model/utils.py:15-28 - key padding mask plus **+1.0 on the lower
    triangle** (sic); true causality comes from the -inf fill in
    :func:`attention` (model/attention.py:165-171)."""
    out = key_padding_additive(mask, dtype, t)
    return out + torch.tril(torch.ones(t, t, dtype=dtype))[None, None]


# --------------------------------------------------------------------------- attention


def attention(sd: SD, p: str, x_q: Tensor, x_kv: Tensor, additive: Tensor, heads: int, kind: str) -> Tensor:
    """model/attention.py:46-76 (self), :97-128 (cross), :148-182 (causal).

    q is scaled *after* the bias; the cross variant feeds ``x_kv / 2`` to
    ``v_proj``; the causal variant fills ``j > i`` with -inf *before* adding
    the additive mask.
    """
    b, tq, d = x_q.shape
    tk = x_kv.shape[1]
    hd = d // heads
    q = linear(sd, p + ".q_proj", x_q) * (hd ** -0.5)
    k = linear(sd, p + ".k_proj", x_kv)
    v = linear(sd, p + ".v_proj", x_kv / 2 if kind == "cross" else x_kv)
    q = q.view(b, tq, heads, hd).transpose(1, 2)
    k = k.view(b, tk, heads, hd).transpose(1, 2)
    v = v.view(b, tk, heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if kind == "causal":
        future = torch.ones(tq, tk, dtype=torch.bool).triu(1)
        s = s.masked_fill(future[None, None], float("-inf"))
    s = s + additive
    o = torch.softmax(s, dim=-1) @ v
    o = o.transpose(1, 2).reshape(b, tq, d)
    return linear(sd, p + ".out_proj", o)


def feed_forward(sd: SD, p: str, x: Tensor) -> Tensor:
    """model/layers.py:94-108 - fc2(GELU_erf(fc1 x)); dropout is identity in eval."""
    return linear(sd, p + ".fc2", F.gelu(linear(sd, p + ".fc1", x)))


def coordinate_attention(sd: SD, p: str, x: Tensor, additive: Tensor, heads: int, attn_type: str) -> Tensor:
    """model/keypoint_module.py:61-80 - LN(x + attn(x)); the self variant adds
    LN(h + FFN(h)); the causal variant has Identity mlp / last_layer_norm."""
    kind = {"self_attn": "self", "causal_attn": "causal"}[attn_type]
    h = layer_norm(sd, p + ".attn_layer_norm", x + attention(sd, p + ".attn", x, x, additive, heads, kind))
    if attn_type == "self_attn":
        h = layer_norm(sd, p + ".last_layer_norm", h + feed_forward(sd, p + ".mlp", h))
    return h


def coordinates_merge(sd: SD, p: str, y: Tensor, x: Tensor, additive: Tensor, heads: int) -> Tensor:
    """model/keypoint_module.py:97-115 - LN(y + cross(y, x)) then LN(h + FFN(h))."""
    h = layer_norm(sd, p + ".attn_layer_norm", y + attention(sd, p + ".attn", y, x, additive, heads, "cross"))
    return layer_norm(sd, p + ".last_layer_norm", h + feed_forward(sd, p + ".mlp", h))


def position_embed(sd: SD, p: str, x: Tensor) -> Tensor:
    """model/layers.py:15-30 - ``x + table[t + 2]``; IndexError when
    ``T > max_position_embeddings``."""
    t = x.shape[1]
    table = _w(sd, p + ".weight", x.dtype)
    if t + 2 > table.shape[0]:
        raise IndexError("index out of range in self")
    return x + table[2 : t + 2][None]


def sca(sd: SD, p: str, x_embed: Tensor, y_embed: Tensor, mask: Tensor, cfg: Mapping, return_maps: bool = False):
    """model/keypoint_module.py:153-198 - SeparativeCoordinateAttention."""
    heads, layers = cfg["attention_heads"], cfg["attn_layers"]
    if cfg.get("self_attn_x", True):
        s_in, c_in = x_embed, y_embed
    else:
        s_in, c_in = y_embed, x_embed
    s = layer_norm(sd, p + ".first_self_norm", position_embed(sd, p + ".self_pos_embed", s_in))
    c = layer_norm(sd, p + ".first_causal_norm", position_embed(sd, p + ".causal_pos_embed", c_in))
    t = c.shape[1]
    pad = key_padding_additive(mask, s.dtype)
    causal = causal_additive(mask, t, c.dtype)
    cross = key_padding_additive(mask, c.dtype, t)
    for i in range(layers):
        s = coordinate_attention(sd, f"{p}.self_attn_layers.{i}", s, pad, heads, "self_attn")
    for i in range(layers):
        c = coordinate_attention(sd, f"{p}.causal_attn_layers.{i}", c, causal, heads, "causal_attn")
        c = coordinates_merge(sd, f"{p}.coordinates_merge.{i}", c, s, cross, heads)
    if return_maps:
        return {"outputs": c, "self_attn_map": s, "causal_attn_map": c}
    return c


# --------------------------------------------------------------------------- residual / pooling


def max_pool_pairs(x: Tensor) -> Tensor:
    """``MaxPool1d(2, 2)`` over time on ``[B,T,C]`` (model/residual.py:40-43): floor(T/2) frames."""
    b, t, c = x.shape
    return x[:, : (t // 2) * 2].reshape(b, t // 2, 2, c).amax(dim=2)


def residual_block(sd: SD, p: str, x: Tensor, downsample: bool) -> Tensor:
    """model/residual.py:25-45."""
    res = linear(sd, p + ".projection", x) if (p + ".projection.weight") in sd else x
    h = torch.relu(layer_norm(sd, p + ".norm1", linear(sd, p + ".linear1", x)))
    h = layer_norm(sd, p + ".norm2", linear(sd, p + ".linear2", h))
    h = torch.relu(h + res)
    return max_pool_pairs(h) if downsample else h


def residual_network(sd: SD, p: str, x: Tensor, blocks: Sequence[int]):
    """model/residual.py:92-118.  The shortcut branches are evaluated only to
    decide, from shapes, whether they would be added (they never are for T>=2,
    SURVEY.md Appendix A.11); when shapes do match they are added as in the
    reference."""
    outputs = []
    history = [x]
    for i in range(len(blocks)):
        y = residual_block(sd, f"{p}.blocks.{i}", x, downsample=(i % 2 == 0))
        if i > 0:
            src = history[i - 2 if i > 1 else 0]
            key = f"{p}.shortcuts.{i - 1}.projection.weight"
            need_proj = key in sd
            need_pool = (i % 2 == 0) and ((i - 1) % 2 == 1)
            sc_t = src.shape[1] // 2 if need_pool else src.shape[1]
            sc_c = sd[key].shape[0] if need_proj else src.shape[2]
            if (src.shape[0], sc_t, sc_c) == tuple(y.shape):
                sc = linear(sd, f"{p}.shortcuts.{i - 1}.projection", src) if need_proj else src
                y = y + (max_pool_pairs(sc) if need_pool else sc)
        x = y
        outputs.append(x)
        history.append(x)
    return x, outputs


# --------------------------------------------------------------------------- stream / fusion / heads


def region_gather(keypoints: Tensor, idx: Sequence[int]) -> Tensor:
    """model/__init__.py:133-142 - ``keypoints[:, :, idx, :]`` (exact copy)."""
    return keypoints[:, :, list(idx), :]


def coordinate_mapping(sd: SD, p: str, kp: Tensor):
    """model/keypoint_module.py:23-24 + model/layers.py:118-123."""
    return linear(sd, p + ".mapping_x", kp[..., 0]), linear(sd, p + ".mapping_y", kp[..., 1])


def keypoint_module(sd: SD, p: str, kp: Tensor, mask: Tensor, cfg: Mapping) -> Tensor:
    """model/keypoint_module.py:22-31; ``kp`` is the already gathered ``[B,T,K_s,2]``."""
    x_embed, y_embed = coordinate_mapping(sd, p + ".coordinate_mapping", kp)
    h = sca(sd, p + ".sca", x_embed, y_embed, mask, cfg)
    out, _ = residual_network(sd, p + ".residual", h, cfg["residual_blocks"])
    return out


def inverted_residual(sd: SD, p: str, x: Tensor) -> Tensor:
    """model/fusion.py:67-78 - LN(GELU(W1 x) + x) -> W3 GELU(W2 .); ``bn1`` is a LayerNorm."""
    h = layer_norm(sd, p + ".bn1", F.gelu(linear(sd, p + ".linear_1", x)) + x)
    return linear(sd, p + ".linear_3", F.gelu(linear(sd, p + ".linear_2", h)))


def coordinates_fusion(sd: SD, p: str, left: Tensor, right: Tensor, body: Tensor) -> Tensor:
    """model/fusion.py:36-55 - queries = right, keys = left, values = body; no
    mask and no 1/sqrt(d) scaling."""
    l = F.gelu(linear(sd, p + ".left_se", left))
    r = F.gelu(linear(sd, p + ".right_se", right))
    bd = F.gelu(linear(sd, p + ".body_se", body))
    a = torch.softmax(r @ l.transpose(1, 2), dim=-1)
    f = layer_norm(sd, p + ".norm", linear(sd, p + ".out_proj", a @ bd))
    return inverted_residual(sd, p + ".inverted_res", f)


def linear_heads(sd: SD, p: str, left: Tensor, right: Tensor, fuse: Tensor, body: Tensor) -> dict:
    """model/__init__.py:49-60 - the four linear classifiers with the +-50 clamp
    (the BiLSTM alignment head is outside the path)."""
    clamp = lambda z: torch.clamp(z, min=-50, max=50)
    return {
        "left": clamp(linear(sd, p + ".left_gloss_classifier", left)),
        "right": clamp(linear(sd, p + ".right_gloss_classifier", right)),
        "body": clamp(linear(sd, p + ".body_gloss_classifier", body)),
        "fuse_coord_gloss_logits": clamp(linear(sd, p + ".fuse_coord_classifier", fuse)),
    }


def encoder_forward(sd: SD, cfg: Mapping, keypoints: Tensor, mask: Tensor, dtype=torch.float32, heads: bool = True,
                    alignment: bool = False) -> dict:
    """model/__init__.py:126-159 restricted to the encoder path: region split,
    three keypoint streams, coordinate fusion and the four linear heads.
    ``sd`` uses the ``MSCA_Net`` key prefixes."""
    kp = keypoints.to(dtype)
    out = {}
    for part in ("body", "left", "right"):
        out[part + "_embed"] = keypoint_module(sd, part + "_encoder", region_gather(kp, cfg[part + "_idx"]), mask, cfg)
    out["fuse_embed"] = coordinates_fusion(sd, "coordinates_fusion", out["left_embed"], out["right_embed"], out["body_embed"])
    if heads:
        out.update(linear_heads(sd, "recognition_head", out["left_embed"], out["right_embed"], out["fuse_embed"], out["body_embed"]))
    if alignment:
        out["alignment_gloss_logits"] = alignment_head(sd, "recognition_head", out["fuse_embed"])
    return out


# --------------------------------------------------------------------------- consumers of the path


def lstm_direction(w_ih: Tensor, w_hh: Tensor, b_ih: Tensor, b_hh: Tensor, x: Tensor, reverse: bool) -> Tensor:
    """One direction of one ``nn.LSTM`` layer, written out (the arithmetic itself is ATen's, which the
    reference reaches through ``nn.LSTM`` at model/alignment_module.py:25-31).  ``x [T, B, in]`` ->
    ``[T, B, H]``; zero initial state; gate rows ordered (input, forget, cell, output)."""
    t_len, b, _ = x.shape
    hid = w_hh.shape[1]
    h = x.new_zeros(b, hid)
    c = x.new_zeros(b, hid)
    out = x.new_empty(t_len, b, hid)
    steps = range(t_len - 1, -1, -1) if reverse else range(t_len)
    for t in steps:
        g = x[t] @ w_ih.T + b_ih + h @ w_hh.T + b_hh
        i, f, gg, o = g.split(hid, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return out


def alignment_module(sd: SD, p: str, x: Tensor, num_layers: int = 2) -> Tensor:
    """model/alignment_module.py:66-72 - ``x [T, B, 1024]`` -> stacked bidirectional LSTM (each layer's
    input is the concatenation forward | reverse of the previous one; inter-layer dropout is identity in
    eval) -> ``permute(1, 0, 2)`` -> ``gloss_layer``.  Returns ``[B, T, V]`` (unclamped)."""
    h = x
    for layer in range(num_layers):
        outs = []
        for sfx, rev in ((f"_l{layer}", False), (f"_l{layer}_reverse", True)):
            w = lambda n: _w(sd, f"{p}.rnn.{n}{sfx}", h.dtype)
            outs.append(lstm_direction(w("weight_ih"), w("weight_hh"), w("bias_ih"), w("bias_hh"), h, rev))
        h = torch.cat(outs, dim=2)
    return linear(sd, p + ".gloss_layer", h.permute(1, 0, 2))


def alignment_head(sd: SD, p: str, fuse: Tensor) -> Tensor:
    """model/__init__.py:51,56 - ``fuse_alignment_head(fuse.permute(1, 0, 2))`` clamped to +-50."""
    return torch.clamp(alignment_module(sd, p + ".fuse_alignment_head", fuse.permute(1, 0, 2)), min=-50, max=50)


def ctc_log_probs(logits: Tensor) -> Tensor:
    """model/__init__.py:243-250 (``compute_loss``) - ``[B,T,V]`` logits -> time-major ``[T,B,V]``
    ``clamp(log_softmax, -100, 0)``, the tensor handed to ``nn.CTCLoss``."""
    return torch.clamp(F.log_softmax(logits.permute(1, 0, 2), dim=-1), min=-100, max=0)


def non_finite_mask(tensors: Sequence[Tensor]) -> int:
    """model/__init__.py:130-167 - the ``isnan(...).any() or isinf(...).any()`` checks, as a bit set."""
    bits = 0
    for i, t in enumerate(tensors):
        if bool(torch.isnan(t).any()) or bool(torch.isinf(t).any()):
            bits |= 1 << i
    return bits


def generic_encoder(sd: SD, p: str, x_embed: Tensor, mask: Tensor, cfg: Mapping) -> Tensor:
    """model/encoder.py:79-92 + :26-57 - the (dead in the live model) generic
    Encoder: pos-embed, LN, N x (self-attn, +res, LN, fc1-GELU-fc2, +res, LN)."""
    pref = p + "." if p else ""
    h = layer_norm(sd, pref + "layernorm_embedding", position_embed(sd, pref + "embed_positions", x_embed))
    add = key_padding_additive(mask, h.dtype)
    for i in range(cfg["encoder_layers"]):
        lp = f"{pref}layers.{i}"
        a = attention(sd, lp + ".self_attn", h, h, add, cfg["encoder_attention_heads"], "self")
        h = layer_norm(sd, lp + ".self_attn_layer_norm", h + a)
        f = linear(sd, lp + ".fc2", F.gelu(linear(sd, lp + ".fc1", h)))
        h = layer_norm(sd, lp + ".final_layer_norm", h + f)
    return h


def encoder_flops_per_frame(cfg: Mapping, t: int, vocab: int = 0) -> float:
    """Algorithmic flops (2*MAC, dense attention, dead shortcuts excluded) per
    *input* frame of :func:`encoder_forward`; BASELINE.md section 4."""
    d, f, L = cfg["d_model"], cfg["ff_dim"], cfg["attn_layers"]
    blocks = cfg["residual_blocks"]
    total = 0.0
    for part in ("body", "left", "right"):
        total += 4 * len(cfg[part + "_idx"]) * d
        total += L * (24 * d * d + 8 * d * f)
        total += L * 12 * t * d
        frac, prev = 1.0, blocks[0]
        for i, c in enumerate(blocks):
            per = 2 * prev * c + 2 * c * c + (2 * prev * c if prev != c else 0)
            total += per * frac
            if i % 2 == 0:
                frac /= 2
            prev = c
    fin, fout = cfg["in_fusion_dim"], cfg["out_fusion_dim"]
    tp = t
    for i in range(len(blocks)):
        if i % 2 == 0:
            tp //= 2
    pooled_frac = tp / t
    fusion = 3 * 2 * fin * fout + 2 * fout * fout + 2 * fout * fout + 2 * 2 * fout * 3 * fout + 4 * tp * fout
    total += fusion * pooled_frac
    if vocab:
        total += (3 * 2 * blocks[-1] * vocab + 2 * fout * vocab) * pooled_frac
    return total
