"""Host-side glue between the PyTorch modules and the C ABI of ``libscatt.so``.

Everything here is plumbing: tensors are allocated by torch (so CUDA-graph
capture and the caching allocator work), their raw device pointers are handed
to the kernels on torch's current stream.  No arithmetic is done in torch.

An activation travelling between kernels is an :class:`Act`: an optional fp32
``[rows, cols]`` matrix (residual stream, attention operands, final outputs)
and optional 16-bit hi/lo split planes ``[2, rows, cols]`` (tensor-core GEMM
operands).  Grouped calls take one entry per anatomical stream and become one
kernel launch (``grid.z`` = stream).
"""

from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib as L

# ----------------------------------------------------------------------------- precision


@dataclass(frozen=True)
class Precision:
    """How contractions are evaluated.

    ``fp32``    fp32 FMA on the CUDA cores (exact-order engine, fp32 tier).
    ``fp16xN``  tcgen05 tensor cores on fp16 hi/lo split planes, N in 1..3
                product terms, fp32 accumulation in TMEM.  ``fp16x3`` has
                fp32-grade products (max-abs ~1e-5 vs the fp32 reference on the
                C1 parity run); ``fp16x1`` is the plain 16-bit product (~7e-3).
    ``bf16xN``  same with bf16 planes (wider range, 3 fewer mantissa bits).
    """

    name: str
    engine: int
    plane_fmt: int
    terms: int

    @property
    def uses_planes(self) -> bool:
        return self.engine == L.ENGINE_TCGEN05

    @property
    def plane_dtype(self):
        return torch.float16 if self.plane_fmt == L.PLANE_F16 else torch.bfloat16


PRECISIONS = {
    "fp32": Precision("fp32", L.ENGINE_SIMT, L.PLANE_F16, 0),
    **{f"fp16x{n}": Precision(f"fp16x{n}", L.ENGINE_TCGEN05, L.PLANE_F16, n) for n in (1, 2, 3)},
    **{f"bf16x{n}": Precision(f"bf16x{n}", L.ENGINE_TCGEN05, L.PLANE_BF16, n) for n in (1, 2, 3)},
}
_default_precision = "fp16x3"


def set_default_precision(name: str) -> None:
    global _default_precision
    if name not in PRECISIONS:
        raise KeyError(f"unknown precision {name!r}; have {sorted(PRECISIONS)}")
    _default_precision = name


def get_precision(name: Optional[str] = None) -> Precision:
    return PRECISIONS[name or _default_precision]


# ----------------------------------------------------------------------------- per-op profiling hook


class OpProfiler:
    """CUDA-event brackets around every C-ABI call made while active (eager
    passes only - events cannot be timed inside graph capture).  Used by
    ``bench.py`` to find the dominant kernel and its achieved rate."""

    def __init__(self):
        self.records = []  # (kernel symbol, start event, end event, algorithmic flops, algorithmic bytes)

    def summary(self):
        """Per kernel symbol (``scatt_last_kernel``: name + template arguments): calls, event-bracketed ms, flops, bytes."""
        torch.cuda.synchronize()
        agg = {}
        for name, e0, e1, flops, nbytes in self.records:
            a = agg.setdefault(name, {"calls": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
            a["calls"] += 1
            a["ms"] += e0.elapsed_time(e1)
            a["flops"] += flops
            a["bytes"] += nbytes
        return agg


_profiler: Optional[OpProfiler] = None


class profile_ops:
    def __enter__(self):
        global _profiler
        _profiler = OpProfiler()
        return _profiler

    def __exit__(self, *exc):
        global _profiler
        _profiler = None
        return False


class _timed:
    __slots__ = ("name", "flops", "nbytes", "e0")

    def __init__(self, name: str, flops: float, nbytes: float):
        self.name, self.flops, self.nbytes = name, flops, nbytes

    def __enter__(self):
        if _profiler is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _profiler is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            sym = L.load().scatt_last_kernel().decode() or self.name  # the symbol the C ABI call actually launched
            _profiler.records.append((sym, self.e0, e1, self.flops, self.nbytes))
        return False


# ----------------------------------------------------------------------------- helpers


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    """The product path has no CPU fallback: fail loudly on host tensors."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "scattennet_b200 runs on a CUDA device (sm_100a) only; got a CPU tensor and there is no CPU fallback"
            )
        dev = t.device
    return dev


def as_f32_2d(x: torch.Tensor) -> torch.Tensor:
    x = x.reshape(-1, x.shape[-1])
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous()


class Act:
    """fp32 matrix and / or split planes of the same ``[rows, cols]`` activation."""

    __slots__ = ("f32", "planes", "rows", "cols")

    def __init__(self, f32: Optional[torch.Tensor] = None, planes: Optional[torch.Tensor] = None):
        ref = f32 if f32 is not None else planes
        self.f32, self.planes = f32, planes
        self.rows, self.cols = (ref.shape[-2], ref.shape[-1])

    @staticmethod
    def from_f32(x: torch.Tensor) -> "Act":
        return Act(as_f32_2d(x))

    def with_planes(self, prec: Precision) -> "Act":
        if prec.uses_planes and self.planes is None:
            self.planes = split_planes(self.f32, prec)
        return self


def split_planes(x: torch.Tensor, prec: Precision, scale: float = 1.0) -> torch.Tensor:
    rows, cols = x.shape
    planes = torch.empty(2, rows, cols, dtype=prec.plane_dtype, device=x.device)
    L.check(L.load().scatt_split_planes(x.data_ptr(), rows, cols, x.stride(0), scale, planes.data_ptr(), prec.plane_fmt,
                                        _stream()), "scatt_split_planes")
    return planes


# ----------------------------------------------------------------------------- side branches

_side_streams = {}
SIDE_BRANCHES = os.environ.get("SCATT_SIDE_BRANCHES", "1") != "0"  # False: every SideBranch body runs in stream order (one chain of launches)


def side_stream(device) -> torch.cuda.Stream:
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


class SideBranch:
    """``with SideBranch(inputs) as br: outs = ...`` runs the body on the device's side stream, forked from
    the current stream; ``br.join(outs)`` makes the current stream wait for it.  Inside a captured forward
    the body becomes a parallel graph branch: at small batches one launch fills a fraction of the 148 SMs,
    so independent launches overlap.  ``inputs`` / ``outs`` are the :class:`Act` that cross streams."""

    def __init__(self, inputs: Sequence["Act"]):
        self.main = torch.cuda.current_stream()
        self.side = side_stream(self.main.device) if SIDE_BRANCHES else self.main
        self.inputs = list(inputs)
        if self.side is self.main:
            return
        fork = torch.cuda.Event()
        fork.record(self.main)
        self.side.wait_event(fork)

    def __enter__(self):
        self._ctx = torch.cuda.stream(self.side)
        self._ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.side is self.main:
            return self._ctx.__exit__(*exc)
        self.done = torch.cuda.Event()
        self.done.record(self.side)
        return self._ctx.__exit__(*exc)

    def join(self, outs: Sequence["Act"] = ()):
        if self.side is self.main:
            return
        torch.cuda.current_stream().wait_event(self.done)
        for a in list(self.inputs) + list(outs):  # keep the caching allocator honest in eager mode
            for t in (a.f32, a.planes):
                if t is not None:
                    t.record_stream(self.side)
                    t.record_stream(self.main)


# ----------------------------------------------------------------------------- packed weights


class PackedLinear:
    """Row-concatenation of one or more ``nn.Linear`` (weight ``[N_i, K]``),
    each optionally scaled by an exact power of two, kept as fp32 for the SIMT
    engine and as split planes for the tcgen05 engine."""

    def __init__(self, linears: Sequence[torch.nn.Linear], scales: Optional[Sequence[float]], key):
        scales = list(scales) if scales is not None else [1.0] * len(linears)
        with torch.no_grad():
            ws = [l.weight.detach().float() * s if s != 1.0 else l.weight.detach().float() for l, s in zip(linears, scales)]
            self.w32 = (torch.cat(ws, 0) if len(ws) > 1 else ws[0]).contiguous()
            if all(l.bias is not None for l in linears):
                bs = [l.bias.detach().float() for l in linears]
                self.b32 = (torch.cat(bs, 0) if len(bs) > 1 else bs[0]).contiguous()
            else:
                self.b32 = None
            # output widths that are not a multiple of 32 (an arbitrary gloss vocabulary) are zero-padded for
            # the kernels; `linear` returns the first n_valid columns
            self.n_valid = self.w32.shape[0]
            pad = (-self.n_valid) % 32
            if pad:
                self.w32 = torch.cat([self.w32, self.w32.new_zeros(pad, self.w32.shape[1])], 0).contiguous()
                if self.b32 is not None:
                    self.b32 = torch.cat([self.b32, self.b32.new_zeros(pad)], 0).contiguous()
        self.N, self.K = self.w32.shape
        self.key = key
        self._planes = {}

    def planes(self, prec: Precision) -> torch.Tensor:
        p = self._planes.get(prec.plane_fmt)
        if p is None:
            p = split_planes(self.w32, prec)
            self._planes[prec.plane_fmt] = p
        return p


def pack_of(owner: torch.nn.Module, tag: str, linears: Sequence[torch.nn.Linear], scales=None) -> PackedLinear:
    """Cached :class:`PackedLinear` of ``linears``; rebuilt when a parameter was
    moved or modified in place (``load_state_dict`` bumps ``_version``)."""
    key = tuple(
        (l.weight.data_ptr(), l.weight._version, 0 if l.bias is None else l.bias.data_ptr(), 0 if l.bias is None else l.bias._version)
        for l in linears
    )
    cache = owner.__dict__.setdefault("_scatt_packs", {})
    ent = cache.get(tag)
    if ent is None or ent.key != key:
        ent = PackedLinear(linears, scales, key)
        cache[tag] = ent
    return ent


L2_PREFETCH = os.environ.get("SCATT_L2_PREFETCH", "1") != "0"  # False: no weight prefetch branch at the top of a small-batch step


def weight_planes_of(root: torch.nn.Module, prec: Precision) -> List[torch.Tensor]:
    """The split-plane copies of every packed weight under ``root`` that exist for this precision (module order)."""
    out = []
    for m in root.modules():
        for pk in m.__dict__.get("_scatt_packs", {}).values():
            p = pk._planes.get(prec.plane_fmt)
            if p is not None:
                out.append(p)
    return out


def l2_prefetch(tensors: Sequence[torch.Tensor]) -> None:
    """``scatt_l2_prefetch``: hint the listed device buffers into L2 (one launch per 1024 buffers, no data dependency)."""
    n = len(tensors)
    if n == 0:
        return
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
    sizes = (C.c_int64 * n)(*[t.numel() * t.element_size() for t in tensors])
    with _timed("l2_prefetch_kernel", 0.0, float(sum(sizes))):
        L.check(L.load().scatt_l2_prefetch(ptrs, sizes, n, _stream()), "scatt_l2_prefetch")


# ----------------------------------------------------------------------------- ops


def make_epilogue(act_pre=L.ACT_NONE, residual_mode=L.RES_NONE, layer_norm=False, act_post=L.ACT_NONE, clamp=0.0,
                  scale_cols=0, scale=1.0, ln_eps=1e-5) -> L.Epilogue:
    return L.Epilogue(act_pre, residual_mode, 1 if layer_norm else 0, act_post, clamp, scale_cols, scale, ln_eps)


def linear(prec: Precision, xs: Sequence[Act], packs: Sequence[PackedLinear], ep: L.Epilogue,
           residuals: Optional[Sequence[torch.Tensor]] = None, lns: Optional[Sequence[torch.nn.LayerNorm]] = None,
           out_f32: bool = True, out_planes: bool = True) -> List[Act]:
    """Grouped ``y_g = epilogue(x_g W_g^T + b_g)`` - one launch for all ``g``."""
    G = len(xs)
    M, K, N = xs[0].rows, xs[0].cols, packs[0].N
    dev = (xs[0].f32 if xs[0].f32 is not None else xs[0].planes).device
    want_planes = out_planes and prec.uses_planes
    # a LayerNorm the engine does not fuse into the GEMM runs as a row-wise tail in place on the fp32 output
    ln_scratch = bool(ep.layer_norm) and not L.load().scatt_linear_ln_fused(M, N, G, prec.engine)
    need_f32 = out_f32 or not want_planes or ln_scratch
    probs = (L.LinearProblem * G)()
    outs: List[Act] = []
    keep = []
    for g in range(G):
        x, pk = xs[g], packs[g]
        if pk.K != K or pk.N != N or x.rows != M:
            raise ValueError("grouped linear: all problems must share M, N, K")
        y = torch.empty(M, N, dtype=torch.float32, device=dev) if need_f32 else None
        yp = torch.empty(2, M, N, dtype=prec.plane_dtype, device=dev) if want_planes else None
        p = probs[g]
        if prec.uses_planes:
            x.with_planes(prec)
            p.x_planes, p.w_planes = x.planes.data_ptr(), pk.planes(prec).data_ptr()
        else:
            p.x, p.w = x.f32.data_ptr(), pk.w32.data_ptr()
        p.bias = _ptr(pk.b32)
        if residuals is not None:
            r = residuals[g]
            if isinstance(r, Act):  # fp32 if the activation has it, else its split planes (tcgen05 engine only)
                if r.f32 is not None:
                    p.residual = r.f32.data_ptr()
                elif prec.uses_planes:
                    p.residual_planes = r.planes.data_ptr()
                else:
                    raise ValueError("the fp32 engine needs an fp32 residual")
            else:
                p.residual = r.data_ptr()
        if lns is not None:
            p.ln_g, p.ln_b = lns[g].weight.data_ptr(), lns[g].bias.data_ptr()
        p.y, p.y_planes = _ptr(y), _ptr(yp)
        outs.append(Act(y, yp))
        keep.append((x, pk))
    ldx = xs[0].f32.stride(0) if xs[0].f32 is not None else K
    r0 = residuals[0] if residuals is not None else None
    if isinstance(r0, Act):
        r0 = r0.f32
    ldres = r0.stride(0) if r0 is not None else N
    name = "linear_tc_kernel" if prec.uses_planes else "linear_simt_kernel"
    if prec.uses_planes:  # operand planes actually read (hi, + lo of x from 2 terms, + lo of W from 3), outputs written
        in_bytes = M * K * 2.0 * (2 if prec.terms >= 2 else 1) + N * K * 2.0 * (2 if prec.terms >= 3 else 1)
    else:
        in_bytes = (M * K + N * K) * 4.0
    out_bytes = M * N * 4.0 * ((1 if need_f32 else 0) + (1 if want_planes else 0)) + (M * N * 4.0 if residuals is not None else 0.0)
    lib = L.load()
    # few row tiles + a long K loop: the launch is split along K through a scratch buffer (scatt_linear_ws)
    ws_bytes = int(lib.scatt_linear_workspace_bytes(G, M, N, K, prec.engine))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if ws_bytes else None
    with _timed(name, 2.0 * G * M * N * K, G * (in_bytes + out_bytes)):
        L.check(lib.scatt_linear_ws(probs, G, M, N, K, ldx, ldres, N, C.byref(ep), prec.engine, prec.plane_fmt,
                                    max(prec.terms, 1), _ptr(ws), ws_bytes, _stream()), "scatt_linear")
    for g in range(G):
        nv = getattr(packs[g], "n_valid", N)
        if nv != N:  # drop the zero-padded columns (plumbing copy; only for widths that are not a multiple of 32)
            if want_planes or ep.layer_norm:
                raise ValueError("grouped linear: a padded output width supports plain fp32 outputs only")
            outs[g] = Act(outs[g].f32[:, :nv].contiguous())
    return outs


FUSED_BLOCK = os.environ.get("SCATT_FUSED_BLOCK", "1") != "0"  # False: out_proj+LN, fc1, fc2+LN as three scatt_linear launches
# Row tiles (128 rows, all streams of the group) from which the fused layer tail replaces the three GEMM launches.
# One CTA (a 2-CTA cluster up to 74 tiles) runs a whole row tile through all three GEMMs, so a handful of tiles leaves
# most SMs idle where the separate launches spread each GEMM over ~100 CTAs.  Measured cross-over at T = 200
# (profiles/r02_sweep_batch_1gpu.md): B = 4 (21 tiles) -4 %, B = 8 (39 tiles) +0.5..1 %, B = 16 +5 %, B = 24 +8 %.
FUSED_BLOCK_MIN_TILES = int(os.environ.get("SCATT_FUSED_BLOCK_MIN_TILES", "36"))


def attn_block_supported(prec: Precision, M: int, D: int, F: int, group: int = 1) -> bool:
    if not (FUSED_BLOCK and prec.uses_planes) or ((M + 127) // 128) * group < FUSED_BLOCK_MIN_TILES:
        return False
    return bool(L.load().scatt_attn_block_supported(M, D, F))


def attn_block(prec: Precision, ctx: Sequence[Act], residuals: Sequence[Act], out_packs: Sequence[PackedLinear],
               norms1: Sequence[torch.nn.LayerNorm], fc1_packs: Sequence[PackedLinear], fc2_packs: Sequence[PackedLinear],
               norms2: Sequence[torch.nn.LayerNorm], out_f32: bool = False, out_planes: bool = True) -> List[Act]:
    """Grouped ``LN2(h + fc2(GELU(fc1(h))))`` with ``h = LN1(x + ctx Wo^T + bo)`` - everything of a self / merge layer
    behind the attention core as ONE launch (``scatt_attn_block``); ``h`` and the hidden activation stay on the SM."""
    G = len(ctx)
    M, D, Fh = ctx[0].rows, ctx[0].cols, fc1_packs[0].N
    dev = ctx[0].planes.device
    probs = (L.BlockProblem * G)()
    outs: List[Act] = []
    for g in range(G):
        residuals[g].with_planes(prec)
        y = torch.empty(M, D, dtype=torch.float32, device=dev) if out_f32 else None
        yp = torch.empty(2, M, D, dtype=prec.plane_dtype, device=dev) if (out_planes or not out_f32) else None
        p = probs[g]
        p.ctx_planes, p.residual_planes = ctx[g].planes.data_ptr(), residuals[g].planes.data_ptr()
        p.wo_planes, p.bo = out_packs[g].planes(prec).data_ptr(), out_packs[g].b32.data_ptr()
        p.ln1_g, p.ln1_b = norms1[g].weight.data_ptr(), norms1[g].bias.data_ptr()
        p.w1_planes, p.b1 = fc1_packs[g].planes(prec).data_ptr(), fc1_packs[g].b32.data_ptr()
        p.w2_planes, p.b2 = fc2_packs[g].planes(prec).data_ptr(), fc2_packs[g].b32.data_ptr()
        p.ln2_g, p.ln2_b = norms2[g].weight.data_ptr(), norms2[g].bias.data_ptr()
        p.y, p.y_planes = _ptr(y), _ptr(yp)
        outs.append(Act(y, yp))
    flops = 2.0 * G * M * (D * D + 2.0 * D * Fh)
    nbytes = G * (3.0 * M * D * 4 + (D * D + 2.0 * D * Fh) * 4)  # ctx + x in, y out (planes), weights once
    with _timed("attn_block_kernel", flops, nbytes):
        L.check(L.load().scatt_attn_block(probs, G, M, D, Fh, norms1[0].eps, prec.plane_fmt, max(prec.terms, 1), _stream()),
                "scatt_attn_block")
    return outs


FUSED_OUT_Q = os.environ.get("SCATT_FUSED_OUT_Q", "1") != "0"  # False: the causal layer's out_proj+LN and the merge layer's q_proj as two launches


def attn_out_q_supported(prec: Precision, M: int, D: int, N: int, group: int = 1) -> bool:
    """Same small-batch / shape envelope as the fused layer tail (it is the same kernel in another mode)."""
    if not (FUSED_BLOCK and FUSED_OUT_Q and prec.uses_planes) or ((M + 127) // 128) * group < FUSED_BLOCK_MIN_TILES:
        return False
    return bool(L.load().scatt_attn_out_q_supported(M, D, N))


def attn_out_q(prec: Precision, ctx: Sequence[Act], residuals: Sequence[Act], out_packs: Sequence[PackedLinear],
               norms: Sequence[torch.nn.LayerNorm], q_packs: Sequence[PackedLinear], q_scale: float):
    """Grouped ``h = LN(x + ctx Wo^T + bo)`` and ``q = (h Wq^T + bq) * q_scale`` as ONE launch (``scatt_attn_out_q``):
    the tail of a causal layer and the q projection of the merge layer that consumes it.  Returns ``(h, q)`` as
    planes-only :class:`Act` lists."""
    G = len(ctx)
    M, D, N = ctx[0].rows, ctx[0].cols, q_packs[0].N
    dev = ctx[0].planes.device
    probs = (L.OutQProblem * G)()
    hs: List[Act] = []
    qs: List[Act] = []
    for g in range(G):
        residuals[g].with_planes(prec)
        hp = torch.empty(2, M, D, dtype=prec.plane_dtype, device=dev)
        qp = torch.empty(2, M, N, dtype=prec.plane_dtype, device=dev)
        p = probs[g]
        p.ctx_planes, p.residual_planes = ctx[g].planes.data_ptr(), residuals[g].planes.data_ptr()
        p.wo_planes, p.bo = out_packs[g].planes(prec).data_ptr(), out_packs[g].b32.data_ptr()
        p.ln_g, p.ln_b = norms[g].weight.data_ptr(), norms[g].bias.data_ptr()
        p.wq_planes, p.bq = q_packs[g].planes(prec).data_ptr(), q_packs[g].b32.data_ptr()
        p.h_planes, p.q_planes = hp.data_ptr(), qp.data_ptr()
        hs.append(Act(None, hp))
        qs.append(Act(None, qp))
    flops = 2.0 * G * M * (D * D + D * N)
    nbytes = G * ((3.0 * M * D + M * N) * 4 + (D * D + D * N) * 4)  # ctx + x in, h + q out (planes), weights once
    with _timed("attn_block_kernel", flops, nbytes):
        L.check(L.load().scatt_attn_out_q(probs, G, M, D, N, norms[0].eps, q_scale, prec.plane_fmt, max(prec.terms, 1), _stream()),
                "scatt_attn_out_q")
    return hs, qs


ATTN_TC_MAX_T = 256  # longest key sequence the tcgen05 attention kernel takes (longer ones run on the fp32 kernel)


def stream_attention(prec: Precision, qs, ks, vs, B: int, Tq: int, Tk: int, H: int, kind: int,
                     key_mask: Optional[torch.Tensor] = None, additive: Optional[torch.Tensor] = None) -> List[Act]:
    """Grouped flash-style attention; ``qs/ks/vs`` are fp32 2-D views (row stride = leading dim)."""
    G = len(qs)
    D = qs[0].shape[1]
    dev = qs[0].device
    probs = (L.AttentionProblem * G)()
    outs = []
    for g in range(G):
        o = torch.empty(B * Tq, D, dtype=torch.float32, device=dev) if not prec.uses_planes else None
        op = torch.empty(2, B * Tq, D, dtype=prec.plane_dtype, device=dev) if prec.uses_planes else None
        p = probs[g]
        p.q, p.k, p.v = qs[g].data_ptr(), ks[g].data_ptr(), vs[g].data_ptr()
        p.key_mask, p.additive = _ptr(key_mask), _ptr(additive)
        p.out, p.out_planes = _ptr(o), _ptr(op)
        outs.append(Act(o, op))
    flops = (2.0 * B * Tq * (Tq + 1) * D if kind == L.ATTN_CAUSAL else 4.0 * B * Tq * Tk * D) * G
    tc = prec.uses_planes and Tk <= ATTN_TC_MAX_T and additive is None
    with _timed("stream_attention_tc_kernel" if tc else "stream_attention_kernel", flops, G * 4.0 * B * max(Tq, Tk) * D * 4):
        L.check(L.load().scatt_attention(probs, G, B, Tq, Tk, H, D // H, qs[0].stride(0), ks[0].stride(0), vs[0].stride(0),
                                         kind, prec.engine, prec.plane_fmt, max(prec.terms, 1), _stream()), "scatt_attention")
    return outs


ATTN_PLANES_MAX_T = 1568  # longest key sequence of the TMA-fed tcgen05 attention kernel (7 blocks of 224 keys)


def stream_attention_planes(prec: Precision, qs, ks, vs, B: int, Tq: int, Tk: int, H: int, kind: int,
                            key_mask: Optional[torch.Tensor] = None) -> List[Act]:
    """Grouped attention whose operands are plane outputs of projection GEMMs.
    ``qs / ks / vs``: per stream ``(planes [2, rows, ld], first column of head 0)``."""
    G = len(qs)
    D = H * 16
    dev = qs[0][0].device
    probs = (L.AttentionPlanesProblem * G)()
    outs = []
    for g in range(G):
        op = torch.empty(2, B * Tq, D, dtype=prec.plane_dtype, device=dev)
        p = probs[g]
        for field, (pl, col) in (("q", qs[g]), ("k", ks[g]), ("v", vs[g])):
            o = getattr(p, field)
            o.planes, o.rows, o.ld, o.col = pl.data_ptr(), pl.shape[1], pl.shape[2], col
        p.key_mask = _ptr(key_mask)
        p.out, p.out_planes = None, op.data_ptr()
        outs.append(Act(None, op))
    flops = (2.0 * B * Tq * (Tq + 1) * D if kind == L.ATTN_CAUSAL else 4.0 * B * Tq * Tk * D) * G
    with _timed("stream_attention_fa_kernel", flops, G * 4.0 * B * max(Tq, Tk) * D * 4):
        L.check(L.load().scatt_attention_planes(probs, G, B, Tq, Tk, H, 16, kind, prec.plane_fmt, max(prec.terms, 1), _stream()),
                "scatt_attention_planes")
    return outs


def fusion_attention(prec: Precision, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, T: int) -> Act:
    D = q.shape[1]
    o = torch.empty(B * T, D, dtype=torch.float32, device=q.device) if not prec.uses_planes else None
    op = torch.empty(2, B * T, D, dtype=prec.plane_dtype, device=q.device) if prec.uses_planes else None
    with _timed("fusion_attention_kernel", 4.0 * B * T * T * D, 4.0 * B * T * D * 4):
        L.check(L.load().scatt_fusion_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), B, T, D, _ptr(o), _ptr(op),
                                                prec.plane_fmt, _stream()), "scatt_fusion_attention")
    return Act(o, op)


def fusion_attention_planes_supported(prec: Precision, T: int, D: int) -> bool:
    return prec.uses_planes and bool(L.load().scatt_fusion_attention_planes_supported(T, D))


def fusion_attention_planes(prec: Precision, q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, T: int,
                            out_f32: bool = False) -> Act:
    """Fusion attention on the tensor cores; ``q / k / v`` are split planes ``[2, B*T, D]`` (outputs of the squeeze GEMMs)."""
    D = q.shape[2]
    o = torch.empty(B * T, D, dtype=torch.float32, device=q.device) if out_f32 else None
    op = torch.empty(2, B * T, D, dtype=prec.plane_dtype, device=q.device)
    with _timed("fusion_attention_tc_kernel", 4.0 * B * T * T * D, 4.0 * B * T * D * 4):
        L.check(L.load().scatt_fusion_attention_planes(q.data_ptr(), k.data_ptr(), v.data_ptr(), B, T, D, _ptr(o), op.data_ptr(),
                                                       prec.plane_fmt, max(prec.terms, 1), _stream()), "scatt_fusion_attention_planes")
    return Act(o, op)


def pool_pairs(prec: Precision, x: torch.Tensor, B: int, T: int) -> Act:
    Cc = x.shape[1]
    if T < 2:
        # MaxPool1d(2, 2) on a single frame: the reference raises (model/residual.py:42)
        raise RuntimeError("max_pool1d() Invalid computed output size: 0")
    y = torch.empty(B * (T // 2), Cc, dtype=torch.float32, device=x.device)
    yp = torch.empty(2, B * (T // 2), Cc, dtype=prec.plane_dtype, device=x.device) if prec.uses_planes else None
    with _timed("pool_pairs_kernel", 0.0, 1.5 * B * T * Cc * 4):
        L.check(L.load().scatt_pool_pairs(x.data_ptr(), B, T, Cc, y.data_ptr(), _ptr(yp), prec.plane_fmt, _stream()),
                "scatt_pool_pairs")
    return Act(y, yp)


def pool_pairs_group(prec: Precision, xs: Sequence[torch.Tensor], B: int, T: int) -> List[Act]:
    """:func:`pool_pairs` for the anatomical streams in one launch."""
    G, Cc = len(xs), xs[0].shape[1]
    if T < 2:
        raise RuntimeError("max_pool1d() Invalid computed output size: 0")
    dev = xs[0].device
    ys = [torch.empty(B * (T // 2), Cc, dtype=torch.float32, device=dev) for _ in range(G)]
    yps = [torch.empty(2, B * (T // 2), Cc, dtype=prec.plane_dtype, device=dev) if prec.uses_planes else None for _ in range(G)]
    arr = lambda ts: (C.c_void_p * G)(*[_ptr(t) for t in ts])
    with _timed("pool_pairs_kernel", 0.0, G * 1.5 * B * T * Cc * 4):
        L.check(L.load().scatt_pool_pairs_group(arr(xs), arr(ys), arr(yps), G, B, T, Cc, prec.plane_fmt, _stream()),
                "scatt_pool_pairs_group")
    return [Act(y, yp) for y, yp in zip(ys, yps)]


def posembed_layernorm(prec: Precision, x: torch.Tensor, table: torch.Tensor, ln: torch.nn.LayerNorm, B: int, T: int) -> Act:
    D = x.shape[-1]
    max_pos = table.shape[0] - 2
    if T > max_pos:
        raise IndexError("index out of range in self")  # what nn.Embedding raises in the reference (model/layers.py:28)
    x2 = as_f32_2d(x)
    out = torch.empty(B * T, D, dtype=torch.float32, device=x.device)
    planes = torch.empty(2, B * T, D, dtype=prec.plane_dtype, device=x.device) if prec.uses_planes else None
    L.check(L.load().scatt_posembed_layernorm(x2.data_ptr(), table.data_ptr(), ln.weight.data_ptr(), ln.bias.data_ptr(),
                                              out.data_ptr(), _ptr(planes), B, T, D, max_pos, prec.plane_fmt, _stream()),
            "scatt_posembed_layernorm")
    return Act(out, planes)


def rowwise(prec: Precision, z: torch.Tensor, ep: L.Epilogue, ln: Optional[torch.nn.LayerNorm] = None,
            residual: Optional[torch.Tensor] = None) -> Act:
    M, N = z.shape
    y = torch.empty(M, N, dtype=torch.float32, device=z.device)
    yp = torch.empty(2, M, N, dtype=prec.plane_dtype, device=z.device) if prec.uses_planes else None
    L.check(L.load().scatt_rowwise(z.data_ptr(), M, N, z.stride(0), _ptr(residual), N if residual is None else residual.stride(0),
                                   _ptr(ln.weight) if ln is not None else None, _ptr(ln.bias) if ln is not None else None,
                                   C.byref(ep), y.data_ptr(), N, _ptr(yp), prec.plane_fmt, _stream()), "scatt_rowwise")
    return Act(y, yp)


def key_mask_u8(mask: torch.Tensor) -> torch.Tensor:
    """``[B, T]`` 0/1 mask of any integer / bool / float dtype -> uint8 (1 = valid key)."""
    return (mask != 0).to(torch.uint8).contiguous()


# ----------------------------------------------------------------------------- consumers of the path


def lstm_bidir(prec: Precision, gates_x: torch.Tensor, w_hh: torch.Tensor, B: int, T: int, H: int,
               out_f32: bool = True) -> Act:
    """Recurrent part of one bidirectional LSTM layer.  ``gates_x [B*T, 8H]`` fp32 (input projections of both
    directions, biases added), ``w_hh [2, 4H, H]`` fp32; returns ``[B*T, 2H]`` (forward | reverse)."""
    dev = gates_x.device
    want_planes = prec.uses_planes
    y = torch.empty(B * T, 2 * H, dtype=torch.float32, device=dev) if (out_f32 or not want_planes) else None
    yp = torch.empty(2, B * T, 2 * H, dtype=prec.plane_dtype, device=dev) if want_planes else None
    lib = L.load()
    ws = torch.empty(max(int(lib.scatt_lstm_workspace_bytes(B, H)), 16), dtype=torch.uint8, device=dev)
    # per step and direction: 2 * B * 4H * H recurrent flops; bytes: gates in, states out (W_hh stays on chip)
    with _timed("lstm_bidir_kernel", 2.0 * 2 * B * T * 4 * H * H, B * T * (8 * H + 2 * H) * 4.0):
        L.check(lib.scatt_lstm_bidir(gates_x.data_ptr(), gates_x.stride(0), w_hh.data_ptr(), _ptr(y), _ptr(yp), ws.data_ptr(),
                                     B, T, H, prec.plane_fmt, _stream()), "scatt_lstm_bidir")
    return Act(y, yp)


def log_softmax_clamp(logits: torch.Tensor, time_major: bool = False, clamp_min: float = -100.0,
                      clamp_max: float = 0.0) -> torch.Tensor:
    """``clamp(log_softmax(logits [B,T,V], -1), clamp_min, clamp_max)`` as ``[B,T,V]`` or ``[T,B,V]``."""
    require_cuda(logits)
    B, T, V = logits.shape
    x = logits if (logits.dtype == torch.float32 and logits.is_contiguous()) else logits.float().contiguous()
    out = torch.empty((T, B, V) if time_major else (B, T, V), dtype=torch.float32, device=x.device)
    with _timed("log_softmax_kernel", 0.0, 2.0 * B * T * V * 4):
        L.check(L.load().scatt_log_softmax(x.data_ptr(), V, V, B, T, 1 if time_major else 0, clamp_min, clamp_max,
                                           out.data_ptr(), _stream()), "scatt_log_softmax")
    return out


def finite_flags(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """Device int32 scalar with bit ``i`` set when ``tensors[i]`` holds a NaN or an infinity (no host sync)."""
    n = len(tensors)
    ts = [t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous() for t in tensors]
    flags = torch.empty(1, dtype=torch.int32, device=ts[0].device)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() for t in ts])
    sizes = (C.c_int64 * n)(*[t.numel() for t in ts])
    L.check(L.load().scatt_finite_check(ptrs, sizes, n, flags.data_ptr(), _stream()), "scatt_finite_check")
    return flags


def ctc_beam_decode(logits: torch.Tensor, lengths: Optional[torch.Tensor] = None, beam: int = 5):
    """CTC prefix beam search (top path) on ``logits [B,T,V]`` (class 0 = blank).  Returns device tensors
    ``(ids [B,T] int32 padded with -1, n_ids [B] int32, log-probability [B] fp32)`` - no host sync."""
    require_cuda(logits)
    B, T, V = logits.shape
    x = logits if (logits.dtype == torch.float32 and logits.is_contiguous()) else logits.float().contiguous()
    dev = x.device
    ids = torch.empty(B, T, dtype=torch.int32, device=dev)
    n_ids = torch.empty(B, dtype=torch.int32, device=dev)
    score = torch.empty(B, dtype=torch.float32, device=dev)
    lens = None if lengths is None else lengths.to(device=dev, dtype=torch.int32).contiguous()
    with _timed("ctc_beam_kernel", 0.0, float(B) * T * V * 4):
        L.check(L.load().scatt_ctc_beam_decode(x.data_ptr(), B, T, V, _ptr(lens), beam, ids.data_ptr(), n_ids.data_ptr(),
                                               score.data_ptr(), _stream()), "scatt_ctc_beam_decode")
    return ids, n_ids, score
