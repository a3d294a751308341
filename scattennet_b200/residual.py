"""``ResidualBlock`` / ``ResidualNetwork`` with the interface of the reference
``model/residual.py`` (Linear-LN-ReLU-Linear-LN, + residual, ReLU, optional
``MaxPool1d(2,2)`` over time).  Each block is two fused GEMM launches (LayerNorm,
residual add and ReLU in the epilogue) plus the temporal-pool kernel.

The ``shortcuts`` keep their parameters (state-dict parity: keys
``shortcuts.{1,2}.projection.*``) but are only evaluated when their output
shape would match the block output - which never happens for T >= 2 (reference
``model/residual.py:110-113``), so the reference's dead
projection flops are not spent.
"""

from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from . import _lib as L
from . import functional as F_
from .functional import Act, Precision


class PermuteLayer(nn.Module):
    def __init__(self, *dims):
        super().__init__()
        self.dims = dims

    def forward(self, x):
        return x.permute(*self.dims)


class ResidualBlock(nn.Module):
    def __init__(self, in_dim, out_dim, downsample=False):
        super().__init__()
        self.downsample = downsample
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.need_projection = in_dim != out_dim
        if self.need_projection:
            self.projection = nn.Linear(in_dim, out_dim)
        self.linear1 = nn.Linear(in_dim, out_dim)
        self.norm1 = nn.LayerNorm(out_dim)
        self.relu = nn.ReLU()
        self.linear2 = nn.Linear(out_dim, out_dim)
        self.norm2 = nn.LayerNorm(out_dim)
        if self.downsample:
            self.pool = nn.MaxPool1d(kernel_size=2, stride=2)
        self.precision: Optional[str] = None

    def forward(self, x):
        F_.require_cuda(x)
        prec = F_.get_precision(self.precision)
        b, t, _ = x.shape
        out, t2 = residual_blocks_forward(prec, [self], [Act.from_f32(x)], b, t)
        return out[0].f32.view(b, t2, self.out_dim).to(x.dtype)


def residual_blocks_forward(prec: Precision, blocks: Sequence[ResidualBlock], xs: List[Act], B: int, T: int):
    """One ``ResidualBlock`` per anatomical stream (same shapes), grouped launches."""
    blk = blocks[0]
    proj_branch = None
    if blk.need_projection:  # reads the same input as linear1: a side branch beside linear1 / norm1
        for x in xs:
            x.with_planes(prec)  # made here, not inside the branch: both GEMMs read them
        with F_.SideBranch(xs) as proj_branch:
            res_acts = F_.linear(prec, xs, [F_.pack_of(m, "projection", [m.projection]) for m in blocks], F_.make_epilogue(),
                                 out_planes=False)
        res = [r.f32 for r in res_acts]
    else:
        res = [x.f32 for x in xs]
    h = F_.linear(prec, xs, [F_.pack_of(m, "linear1", [m.linear1]) for m in blocks],
                  F_.make_epilogue(layer_norm=True, act_post=L.ACT_RELU), lns=[m.norm1 for m in blocks],
                  out_f32=not prec.uses_planes)
    if proj_branch is not None:
        proj_branch.join(res_acts)
    out = F_.linear(prec, h, [F_.pack_of(m, "linear2", [m.linear2]) for m in blocks],
                    F_.make_epilogue(layer_norm=True, residual_mode=L.RES_AFTER_LN, act_post=L.ACT_RELU), residuals=res,
                    lns=[m.norm2 for m in blocks], out_planes=not blk.downsample)
    if blk.downsample:
        out = F_.pool_pairs_group(prec, [o.f32 for o in out], B, T)
        T = T // 2
    return out, T


class ResidualNetwork(nn.Module):
    def __init__(self, residual_blocks):
        super().__init__()
        self.residual_blocks = residual_blocks
        self.blocks = nn.ModuleList()
        self.shortcuts = nn.ModuleList()
        for i in range(len(residual_blocks)):
            in_dim = residual_blocks[i - 1] if i > 0 else residual_blocks[0]
            out_dim = residual_blocks[i]
            self.blocks.append(ResidualBlock(in_dim, out_dim, downsample=(i % 2 == 0)))
            if i > 0:
                # same registration rule as the reference (model/residual.py:63-90)
                prev = residual_blocks[i - 2] if i > 1 else residual_blocks[0]
                need_projection = prev != residual_blocks[i]
                need_downsample = (i % 2 == 0) and ((i - 1) % 2 == 1)
                if need_projection or need_downsample:
                    shortcut = nn.Sequential()
                    if need_projection:
                        shortcut.add_module("projection", nn.Linear(prev, residual_blocks[i]))
                    if need_downsample:
                        shortcut.add_module("permute1", PermuteLayer(0, 2, 1))
                        shortcut.add_module("pool", nn.MaxPool1d(kernel_size=2, stride=2))
                        shortcut.add_module("permute2", PermuteLayer(0, 2, 1))
                    self.shortcuts.append(shortcut)
                else:
                    self.shortcuts.append(None)
        self.precision: Optional[str] = None

    def forward(self, x):
        F_.require_cuda(x)
        prec = F_.get_precision(self.precision)
        b, t, _ = x.shape
        outs = residual_network_forward(prec, [self], [Act.from_f32(x)], b, t)
        views = [o[0].f32.view(b, tt, -1).to(x.dtype) for o, tt in outs]
        return views[-1], views


def residual_network_forward(prec: Precision, nets: Sequence[ResidualNetwork], xs: List[Act], B: int, T: int):
    """Returns ``[(acts_per_stream, T_i)]`` for every block output."""
    net = nets[0]
    history = [(xs, T)]
    results = []
    cur, t = xs, T
    for i in range(len(net.blocks)):
        out, t_out = residual_blocks_forward(prec, [n.blocks[i] for n in nets], cur, B, t)
        if i > 0:
            src, t_src = history[i - 2 if i > 1 else 0]
            sc = net.shortcuts[i - 1]
            has_proj = sc is not None and hasattr(sc, "projection")
            has_pool = sc is not None and hasattr(sc, "pool")
            sc_t = t_src // 2 if has_pool else t_src
            sc_c = sc.projection.out_features if has_proj else src[0].cols
            if (sc_t, sc_c) == (t_out, out[0].cols):  # never true for T >= 2; kept for fidelity
                out = _add_shortcut(prec, nets, i, src, t_src, out, B, has_proj, has_pool)
        cur, t = out, t_out
        results.append((out, t_out))
        history.append((out, t_out))
    return results


def _add_shortcut(prec, nets, i, src, t_src, out, B, has_proj, has_pool):
    if has_proj:
        sc = F_.linear(prec, src, [F_.pack_of(n.shortcuts[i - 1], "projection", [n.shortcuts[i - 1].projection]) for n in nets],
                       F_.make_epilogue(), out_planes=False)
        sc = [s.f32 for s in sc]
    else:
        sc = [s.f32 for s in src]
    if has_pool:
        sc = [F_.pool_pairs(prec, s, B, t_src).f32 for s in sc]
    return [F_.rowwise(prec, o.f32, F_.make_epilogue(residual_mode=L.RES_AFTER_LN), residual=s) for o, s in zip(out, sc)]
