"""``KeypointModule``, ``CoordinateAttention``, ``CoordinatesMerge`` and
``SeparativeCoordinateAttention`` with the interface of the reference
``model/keypoint_module.py``.

Data flow of one stream (reference ``:22-31,153-198``):

    keypoints[B,T,K_s,2] --K1 front end--> self / causal embeddings (LN'd)
    4 x self layer      : QKV GEMM -> attention -> out_proj+res+LN -> fc1+GELU -> fc2+res+LN
    1 x cross K/V GEMM  : K,V of all 4 merge layers from the final self map (N = 4*2*D)
    4 x (causal layer   : QKV GEMM -> causal attention -> out_proj+res+LN
         merge layer    : Q GEMM -> cross attention -> out_proj+res+LN -> fc1+GELU -> fc2+res+LN)
    residual network    : see residual.py

Every function here takes *lists* (one entry per anatomical stream with equal
shapes); the three streams of the model run as grouped launches.
"""

from __future__ import annotations

import os

import ctypes as C
from typing import List, Optional, Sequence

import torch
from torch import nn
from torch.nn import functional as F

from . import _lib as L
from . import functional as F_
from .attention import CrossAttention, SelfAttention, SelfCausalAttention, attention_core
from .functional import Act, Precision
from .layers import CoordinateMapping, FeedForward, LearningPositionEmbedding
from .residual import ResidualNetwork, residual_network_forward
from .utils import create_attention_mask, create_causal_attention_mask  # noqa: F401  (interface parity)


def _inference_only(mod: nn.Module, p: float):
    if mod.training and p > 0:
        raise RuntimeError("scattennet_b200 is inference-only: call .eval() before forward (dropout > 0 in training mode)")


def _additive(mask: Optional[torch.Tensor], b: int, tq: int, tk: int) -> Optional[torch.Tensor]:
    if mask is None:
        return None
    return mask.to(torch.float32).expand(b, 1, tq, tk).contiguous()


class CoordinateAttention(nn.Module):
    def __init__(self, cfg, attn_type="self_attn"):
        super().__init__()
        self.attn_type = attn_type
        if attn_type == "self_attn":
            self.attn = SelfAttention(d_model=cfg["d_model"], num_heads=cfg["attention_heads"], dropout=cfg["attention_dropout"])
            self.mlp = FeedForward(cfg["d_model"], cfg["ff_dim"], cfg["dropout"])
            self.last_layer_norm = nn.LayerNorm(cfg["d_model"])
        elif attn_type == "causal_attn":
            self.attn = SelfCausalAttention(d_model=cfg["d_model"], num_heads=cfg["attention_heads"], dropout=cfg["attention_dropout"])
            self.mlp = nn.Identity()
            self.last_layer_norm = nn.Identity()
        else:
            raise ValueError(f"Invalid attention type: {attn_type}")
        self.attn_layer_norm = nn.LayerNorm(cfg["d_model"])
        self.dropout = cfg["dropout"]
        self.activation_fn = nn.GELU()
        self.precision: Optional[str] = None

    def forward(self, coord_embed, attention_mask=None):
        _inference_only(self, self.dropout)
        F_.require_cuda(coord_embed, attention_mask)
        prec = F_.get_precision(self.precision)
        b, t, _ = coord_embed.shape
        out = coordinate_attention_forward(prec, [self], [Act.from_f32(coord_embed)], b, t, None,
                                           _additive(attention_mask, b, t, t))
        return out[0].f32.view_as(coord_embed).to(coord_embed.dtype)


# On the tensor-core engine the residual stream between the layers of a ladder lives in split planes only:
# an activation that is just the next GEMM's operand and the next LayerNorm's residual is written once (planes),
# not twice (planes + fp32) - the outbound store path is what bounds these launches.  `keep_f32` asks for the
# fp32 copy as well (module-level calls, final outputs, anything a non-GEMM consumer reads).


# Measured: +4 % frames/s at B = 256.  At B = 8 it used to cost 1.5 % while the cluster kernels fetched the residual
# with ld.global (the plane conversion sat on the critical path); with the TMA-staged residual the planes arrive
# beside the operands and only the halved output traffic remains.  SCATT_PLANES_ONLY_SMALL=0 restores fp32 + planes
# below the large-batch threshold (A/B).
_PLANES_ONLY_SMALL = os.environ.get("SCATT_PLANES_ONLY_SMALL", "1") != "0"


def _keep_f32(prec, acts: List[Act], keep_f32: bool) -> bool:
    # scatt_linear runs an N = 256 LayerNorm GEMM one CTA per 128-row tile once the row tiles of the group exceed 74
    large = ((acts[0].rows + 127) // 128) * len(acts) > 74
    return keep_f32 or not prec.uses_planes or not (large or _PLANES_ONLY_SMALL)


def _attn_out_ln(prec, attns, ctx, norms, residuals: List[Act], keep_f32: bool = True) -> List[Act]:
    keep_f32 = _keep_f32(prec, ctx, keep_f32)
    return F_.linear(prec, ctx, [F_.pack_of(a, "out", [a.out_proj]) for a in attns],
                     F_.make_epilogue(residual_mode=L.RES_BEFORE_LN, layer_norm=True), residuals=residuals,
                     lns=norms, out_f32=keep_f32 or not prec.uses_planes)


def _ffn_ln(prec, mlps: Sequence[FeedForward], norms, h: List[Act], keep_f32: bool = True) -> List[Act]:
    keep_f32 = _keep_f32(prec, h, keep_f32)
    f = F_.linear(prec, h, [F_.pack_of(m, "fc1", [m.fc1]) for m in mlps], F_.make_epilogue(act_pre=L.ACT_GELU),
                  out_f32=not prec.uses_planes)
    return F_.linear(prec, f, [F_.pack_of(m, "fc2", [m.fc2]) for m in mlps],
                     F_.make_epilogue(residual_mode=L.RES_BEFORE_LN, layer_norm=True), residuals=h, lns=norms,
                     out_f32=keep_f32 or not prec.uses_planes)


def _fused_tail_ok(prec, attns, mlps, norms1, norms2, ctx: List[Act], residuals: List[Act]) -> bool:
    """``scatt_attn_block`` takes the tail of the layer: plane operands, biased linears, both LayerNorms alike."""
    if not F_.attn_block_supported(prec, ctx[0].rows, ctx[0].cols, mlps[0].fc1.out_features, len(ctx)):
        return False
    if any(c.planes is None for c in ctx) or any(r.planes is None and r.f32 is None for r in residuals):
        return False
    if any(l.bias is None for a, m in zip(attns, mlps) for l in (a.out_proj, m.fc1, m.fc2)):
        return False
    return all(n1.eps == norms1[0].eps == n2.eps for n1, n2 in zip(norms1, norms2))


def _attn_tail(prec, attns, mlps, norms1, norms2, ctx: List[Act], residuals: List[Act], keep_f32: bool) -> List[Act]:
    """out_proj + residual + LayerNorm, FeedForward + residual + LayerNorm (reference
    ``model/keypoint_module.py:62-72,98-107``): one fused launch when the shapes allow, three GEMM launches otherwise."""
    if _fused_tail_ok(prec, attns, mlps, norms1, norms2, ctx, residuals):
        return F_.attn_block(prec, ctx, residuals, [F_.pack_of(a, "out", [a.out_proj]) for a in attns], norms1,
                             [F_.pack_of(m, "fc1", [m.fc1]) for m in mlps], [F_.pack_of(m, "fc2", [m.fc2]) for m in mlps],
                             norms2, out_f32=keep_f32)
    h = _attn_out_ln(prec, attns, ctx, norms1, residuals, False)
    return _ffn_ln(prec, mlps, norms2, h, keep_f32)


def coordinate_attention_forward(prec: Precision, mods: Sequence[CoordinateAttention], xs: List[Act], B: int, T: int,
                                 key_mask: Optional[torch.Tensor], additive: Optional[torch.Tensor] = None,
                                 keep_f32: bool = True, q_for: Optional[Sequence["CoordinatesMerge"]] = None):
    """reference ``model/keypoint_module.py:61-80`` for a group of streams.  ``q_for`` (causal layers on the plane
    path): the merge layers that consume the result - their q projection is fused into this layer's tail and
    ``(outputs, q)`` is returned."""
    kind = L.ATTN_SELF if mods[0].attn_type == "self_attn" else L.ATTN_CAUSAL
    is_self = mods[0].attn_type == "self_attn"
    ctx = attention_core(prec, [m.attn for m in mods], xs, None, B, T, T, kind, key_mask, additive)
    if q_for is not None:
        qa = [m.attn for m in q_for]
        return F_.attn_out_q(prec, ctx, xs, [F_.pack_of(m.attn, "out", [m.attn.out_proj]) for m in mods],
                             [m.attn_layer_norm for m in mods], [F_.pack_of(a, "q", [a.q_proj]) for a in qa], qa[0].scaling)
    if is_self:
        return _attn_tail(prec, [m.attn for m in mods], [m.mlp for m in mods], [m.attn_layer_norm for m in mods],
                          [m.last_layer_norm for m in mods], ctx, xs, keep_f32)
    return _attn_out_ln(prec, [m.attn for m in mods], ctx, [m.attn_layer_norm for m in mods], xs, keep_f32)


class CoordinatesMerge(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.attn = CrossAttention(d_model=cfg["d_model"], num_heads=cfg["attention_heads"], dropout=cfg["attention_dropout"])
        self.mlp = FeedForward(cfg["d_model"], cfg["ff_dim"], cfg["dropout"])
        self.attn_layer_norm = nn.LayerNorm(cfg["d_model"])
        self.last_layer_norm = nn.LayerNorm(cfg["d_model"])
        self.dropout = cfg["dropout"]
        self.precision: Optional[str] = None

    def forward(self, y_embed, x_embed, cross_attn_mask=None):
        _inference_only(self, self.dropout)
        F_.require_cuda(y_embed, x_embed, cross_attn_mask)
        prec = F_.get_precision(self.precision)
        b, tq, _ = y_embed.shape
        tk = x_embed.shape[1]
        out = coordinates_merge_forward(prec, [self], [Act.from_f32(y_embed)], [Act.from_f32(x_embed)], None, b, tq, tk, None,
                                        _additive(cross_attn_mask, b, tq, tk))
        return out[0].f32.view_as(y_embed).to(y_embed.dtype)


def coordinates_merge_forward(prec: Precision, mods: Sequence[CoordinatesMerge], ys: List[Act], xs: Optional[List[Act]],
                              kv_views, B: int, Tq: int, Tk: int, key_mask: Optional[torch.Tensor],
                              additive: Optional[torch.Tensor] = None, keep_f32: bool = True,
                              q_planes: Optional[List[Act]] = None) -> List[Act]:
    """reference ``model/keypoint_module.py:97-115`` for a group of streams."""
    ctx = attention_core(prec, [m.attn for m in mods], ys, xs, B, Tq, Tk, L.ATTN_CROSS, key_mask, additive, kv_views, q_planes)
    return _attn_tail(prec, [m.attn for m in mods], [m.mlp for m in mods], [m.attn_layer_norm for m in mods],
                      [m.last_layer_norm for m in mods], ctx, ys, keep_f32)


class SeparativeCoordinateAttention(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        self.dropout = cfg["dropout"]
        n = cfg["attn_layers"]
        self.self_attn_layers = nn.ModuleList([CoordinateAttention(cfg, attn_type="self_attn") for _ in range(n)])
        self.causal_attn_layers = nn.ModuleList([CoordinateAttention(cfg, attn_type="causal_attn") for _ in range(n)])
        self.coordinates_merge = nn.ModuleList([CoordinatesMerge(cfg) for _ in range(n)])
        self.first_self_norm = nn.LayerNorm(cfg["d_model"])
        self.first_causal_norm = nn.LayerNorm(cfg["d_model"])
        self.self_pos_embed = LearningPositionEmbedding(cfg["max_position_embeddings"], cfg["d_model"])
        self.causal_pos_embed = LearningPositionEmbedding(cfg["max_position_embeddings"], cfg["d_model"])
        self.x_self = cfg.get("self_attn_x", True)
        self.precision: Optional[str] = None

    def forward(self, x_embed, y_embed, attention_mask=None, return_attn_map=False):
        _inference_only(self, self.dropout)
        F_.require_cuda(x_embed, y_embed, attention_mask)
        prec = F_.get_precision(self.precision)
        b, t, d = x_embed.shape
        s_in, c_in = (x_embed, y_embed) if self.x_self else (y_embed, x_embed)
        s = F_.posembed_layernorm(prec, s_in, self.self_pos_embed.weight, self.first_self_norm, b, t)
        c = F_.posembed_layernorm(prec, c_in, self.causal_pos_embed.weight, self.first_causal_norm, b, t)
        km = F_.key_mask_u8(attention_mask)  # the reference requires a [B,T] mask here too (model/utils.py:5)
        outs, selfs = sca_forward(prec, [self], [s], [c], km, b, t, need_self_f32=return_attn_map)
        outputs = outs[0].f32.view(b, t, d).to(x_embed.dtype)
        if return_attn_map:
            return {"outputs": outputs, "self_attn_map": selfs[0].f32.view(b, t, d).to(x_embed.dtype),
                    "causal_attn_map": outputs}
        return outputs


OVERLAP_BRANCHES = True  # run the first causal layer concurrently with the self branch
_side_stream = F_.side_stream


def sca_forward(prec: Precision, mods: Sequence[SeparativeCoordinateAttention], s: List[Act], c: List[Act],
                key_mask: torch.Tensor, B: int, T: int, need_self_f32: bool = False):
    """Layer loops of reference ``model/keypoint_module.py:176-187`` on already
    position-embedded + normalised inputs, for a group of streams."""
    n = len(mods[0].self_attn_layers)
    d = s[0].cols
    # The first causal layer only needs the causal-branch embedding: run it on a side stream while the
    # self branch works (inside a captured forward this becomes a parallel graph branch).
    c_first, q_first, join = None, None, None
    # causal layer i feeds merge layer i only: its out_proj + LayerNorm and the merge layer's q projection run as one launch
    # (scatt_attn_out_q) where the fused-tail kernel is in use
    fuse_q = (n > 0 and T <= F_.ATTN_PLANES_MAX_T and d // mods[0].causal_attn_layers[0].attn.num_heads == 16
              and all(c_.planes is not None for c_ in c)
              and all(l.bias is not None for m in mods for i in range(n) for l in (m.causal_attn_layers[i].attn.out_proj, m.coordinates_merge[i].attn.q_proj))
              and F_.attn_out_q_supported(prec, c[0].rows, d, d, len(mods)))
    if OVERLAP_BRANCHES and n > 0:
        main = torch.cuda.current_stream()
        side = _side_stream(main.device)
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        with torch.cuda.stream(side):
            c_first = coordinate_attention_forward(prec, [m.causal_attn_layers[0] for m in mods], c, B, T, key_mask, keep_f32=False,
                                                   q_for=[m.coordinates_merge[0] for m in mods] if fuse_q else None)
            if fuse_q:
                c_first, q_first = c_first
            join = torch.cuda.Event()
            join.record(side)
        for a in list(c) + list(c_first) + list(q_first or []):  # tensors that cross streams: keep the allocator honest in eager mode
            for t in (a.f32, a.planes):
                if t is not None:
                    t.record_stream(side)
                    t.record_stream(main)
    for i in range(n):
        s = coordinate_attention_forward(prec, [m.self_attn_layers[i] for m in mods], s, B, T, key_mask,
                                         keep_f32=need_self_f32 and i == n - 1)
    # K / V of every merge layer read the same final self map: the first layer's pair (N = 2 D) on this stream,
    # the pairs of the remaining layers as one N = (n - 1) * 2 D GEMM per stream on a side branch that runs
    # beside the first merge layer
    def kv_pack(m, tag, layers):
        lins, scales = [], []
        for i in layers:
            a = m.coordinates_merge[i].attn
            lins += [a.k_proj, a.v_proj]
            scales += [1.0, 0.5]
        return F_.pack_of(m, tag, lins, scales)

    kv_planes = prec.uses_planes and T <= F_.ATTN_PLANES_MAX_T  # TMA-fed attention takes the planes as they are
    kv_rest, kv_branch = None, None
    split_kv = OVERLAP_BRANCHES and n > 1
    if split_kv:
        for a in s:
            a.with_planes(prec)  # made here, not inside the branch: both GEMMs read them
        with F_.SideBranch(s) as kv_branch:
            kv_rest = F_.linear(prec, s, [kv_pack(m, "merge_kv_rest", range(1, n)) for m in mods], F_.make_epilogue(),
                                out_f32=not kv_planes, out_planes=kv_planes)
    kv_first = F_.linear(prec, s, [kv_pack(m, "merge_kv_first" if split_kv else "merge_kv", range(1 if split_kv else n)) for m in mods],
                         F_.make_epilogue(), out_f32=not kv_planes, out_planes=kv_planes)
    for i in range(n):
        q_pre = None
        if i == 0 and c_first is not None:
            torch.cuda.current_stream().wait_event(join)
            c, q_pre = c_first, q_first
        else:
            c = coordinate_attention_forward(prec, [m.causal_attn_layers[i] for m in mods], c, B, T, key_mask, keep_f32=False,
                                             q_for=[m.coordinates_merge[i] for m in mods] if fuse_q else None)
            if fuse_q:
                c, q_pre = c
        if split_kv and i == 1:
            kv_branch.join(kv_rest)
        kv_src, j = (kv_rest, i - 1) if (split_kv and i >= 1) else (kv_first, i)
        if kv_planes:
            kv_views = [((kv.planes, 2 * j * d), (kv.planes, (2 * j + 1) * d)) for kv in kv_src]
        else:
            kv_views = [(kv.f32[:, 2 * j * d : (2 * j + 1) * d], kv.f32[:, (2 * j + 1) * d : (2 * j + 2) * d]) for kv in kv_src]
        c = coordinates_merge_forward(prec, [m.coordinates_merge[i] for m in mods], c, None, kv_views, B, T, T, key_mask,
                                      keep_f32=i == n - 1, q_planes=q_pre if kv_planes else None)  # the ladder's result feeds the residual network in fp32
    return c, s


class KeypointModule(nn.Module):
    def __init__(self, joint_idx, num_frame, cfg=None):
        super().__init__()
        self.joint_idx = joint_idx
        self.num_frame = num_frame
        self.coordinate_mapping = CoordinateMapping(len(joint_idx), cfg["d_model"])
        self.sca = SeparativeCoordinateAttention(cfg)
        self.residual = ResidualNetwork(cfg["residual_blocks"])
        self.precision: Optional[str] = None

    def forward(self, keypoints, attention_mask=None):
        """``keypoints``: the already gathered ``[B,T,K_s,2]`` region, like the reference."""
        _inference_only(self, self.sca.dropout)
        F_.require_cuda(keypoints, attention_mask)
        prec = F_.get_precision(self.precision)
        b, t, k, _ = keypoints.shape
        kp = keypoints.float().contiguous()
        idx = torch.arange(k, dtype=torch.int32, device=kp.device)
        outs = streams_forward(prec, [self], kp, [idx], F_.key_mask_u8(attention_mask), b, t)
        (acts, t_out) = outs[-1]
        return acts[0].f32.view(b, t_out, -1).to(keypoints.dtype)


def _transposed_weight(lin: nn.Linear) -> torch.Tensor:
    """``lin.weight^T`` ([K_s, D], contiguous) cached on the module; rebuilt when the parameter changes."""
    key = (lin.weight.data_ptr(), lin.weight._version)
    ent = lin.__dict__.get("_scatt_wt")
    if ent is None or ent[0] != key:
        ent = (key, lin.weight.detach().float().t().contiguous())
        lin.__dict__["_scatt_wt"] = ent
    return ent[1]


def frontend_forward(prec: Precision, mods: Sequence[KeypointModule], keypoints: torch.Tensor,
                     joint_idx: Sequence[torch.Tensor], B: int, T: int, want_gathered: bool = False, want_f32: bool = False):
    """K1: region gather + x/y split + CoordinateMapping + position embedding +
    first LayerNorm for all streams in one launch.  Returns ``(self_acts,
    causal_acts, gathered)``."""
    dev = keypoints.device
    K = keypoints.shape[2]
    d = mods[0].coordinate_mapping.mapping_x.out_features
    max_pos = mods[0].sca.self_pos_embed.weight.shape[0] - 2
    if T > max_pos:
        raise IndexError("index out of range in self")
    G = len(mods)
    arr = (L.FrontendStream * G)()
    s_acts, c_acts, gathered = [], [], []
    for g, m in enumerate(mods):
        sca, cm = m.sca, m.coordinate_mapping
        maps = (cm.mapping_x, cm.mapping_y) if sca.x_self else (cm.mapping_y, cm.mapping_x)
        coords = (0, 1) if sca.x_self else (1, 0)
        tables = (sca.self_pos_embed.weight, sca.causal_pos_embed.weight)
        norms = (sca.first_self_norm, sca.first_causal_norm)
        st = arr[g]
        st.joint_idx = joint_idx[g].data_ptr()
        st.n_joints = int(joint_idx[g].numel())
        acts = []
        for br in range(2):
            st.coord[br] = coords[br]
            st.map_wt[br] = _transposed_weight(maps[br]).data_ptr()
            st.map_b[br] = maps[br].bias.data_ptr()
            st.pos[br] = tables[br].data_ptr()
            st.ln_g[br] = norms[br].weight.data_ptr()
            st.ln_b[br] = norms[br].bias.data_ptr()
            # tensor-core engine: the embeddings are GEMM operands and LayerNorm residuals - both read the split planes,
            # so the fp32 rows are not written at all (12.3 -> 6.1 KB of output per frame and stream pair)
            o = torch.empty(B * T, d, dtype=torch.float32, device=dev) if (want_f32 or not prec.uses_planes) else None
            op = torch.empty(2, B * T, d, dtype=prec.plane_dtype, device=dev) if prec.uses_planes else None
            st.out[br] = F_._ptr(o)
            st.out_planes[br] = F_._ptr(op)
            acts.append(Act(o, op))
        if want_gathered:
            gt = torch.empty(B, T, st.n_joints, 2, dtype=torch.float32, device=dev)
            st.gathered = gt.data_ptr()
            gathered.append(gt)
        s_acts.append(acts[0])
        c_acts.append(acts[1])
    n_used = sum(int(j.numel()) for j in joint_idx)
    out_bytes = 2 * G * d * ((4 if (want_f32 or not prec.uses_planes) else 0) + (4 if prec.uses_planes else 0))
    with F_._timed("frontend_kernel", 4.0 * n_used * d * B * T, float(B * T) * (n_used * 8 + out_bytes)):
        L.check(L.load().scatt_frontend(keypoints.data_ptr(), B, T, K, d, arr, G, max_pos, prec.plane_fmt, F_._stream()),
                "scatt_frontend")
    return s_acts, c_acts, gathered


def streams_forward(prec: Precision, mods: Sequence[KeypointModule], keypoints: torch.Tensor,
                    joint_idx: Sequence[torch.Tensor], key_mask: torch.Tensor, B: int, T: int):
    """Front end + SCA + residual network for a group of streams reading the
    same ``keypoints[B,T,K,2]``; returns the residual network's block outputs."""
    s, c, _ = frontend_forward(prec, mods, keypoints, joint_idx, B, T)
    h, _ = sca_forward(prec, [m.sca for m in mods], s, c, key_mask, B, T)
    return residual_network_forward(prec, [m.residual for m in mods], h, B, T)
