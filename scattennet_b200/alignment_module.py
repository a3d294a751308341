"""``AlignmentModule`` of the reference (``model/alignment_module.py:5-72``): a
2-layer bidirectional LSTM over the fused features followed by the gloss
classifier - the producer of ``alignment_gloss_logits``, the first consumer of
the encoder path.

Same constructor arguments, ``forward`` signature and state-dict layout
(``rnn.weight_ih_l0`` ... ``rnn.bias_hh_l1_reverse``, ``gloss_layer.*``).  Per
layer: one tensor-core GEMM computes the input projections of every time step
and both directions, then one persistent kernel (``scatt_lstm_bidir``) runs the
recurrence with ``W_hh`` resident in shared memory.  Like the reference, padded
frames are NOT skipped (it feeds the padded batch to ``nn.LSTM`` unpacked).
"""

from __future__ import annotations

import torch
from torch import nn

from . import _lib as L
from . import functional as F_
from .functional import Act


class _PackedRNNInput:
    """``W_ih`` of both directions stacked ``[8H, in]`` with ``b_ih + b_hh`` as the GEMM bias, and ``W_hh``
    stacked ``[2, 4H, H]``; rebuilt when a parameter changed (same rule as ``functional.pack_of``)."""

    def __init__(self, rnn: nn.LSTM, layer: int):
        sfx = [f"_l{layer}", f"_l{layer}_reverse"]
        g = lambda n: getattr(rnn, n).detach().float()
        with torch.no_grad():
            self.w32 = torch.cat([g("weight_ih" + s) for s in sfx], 0).contiguous()
            self.b32 = torch.cat([g("bias_ih" + s) + g("bias_hh" + s) for s in sfx], 0).contiguous()
            self.w_hh = torch.stack([g("weight_hh" + s) for s in sfx], 0).contiguous()
        self.N, self.K = self.w32.shape
        self._planes = {}

    def planes(self, prec):
        p = self._planes.get(prec.plane_fmt)
        if p is None:
            p = F_.split_planes(self.w32, prec)
            self._planes[prec.plane_fmt] = p
        return p


def _rnn_pack(owner: nn.Module, rnn: nn.LSTM, layer: int) -> _PackedRNNInput:
    names = [f"{k}_l{layer}{r}" for r in ("", "_reverse") for k in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    key = tuple((getattr(rnn, n).data_ptr(), getattr(rnn, n)._version) for n in names)
    cache = owner.__dict__.setdefault("_scatt_rnn_packs", {})
    ent = cache.get(layer)
    if ent is None or ent[0] != key:
        ent = (key, _PackedRNNInput(rnn, layer))
        cache[layer] = ent
    return ent[1]


def alignment_forward(prec, m: "AlignmentModule", x: Act, B: int, T: int, clamp: float = 0.0) -> Act:
    """``x``: fused features, rows ``b*T + t``.  Returns the ``[B*T, V]`` logits (batch-major like the
    reference's ``permute(1, 0, 2)`` result), clamped to ``+-clamp`` when ``clamp > 0``."""
    H = m.lstm_hidden_size
    for layer in range(m.num_layers):
        pk = _rnn_pack(m, m.rnn, layer)
        gates = F_.linear(prec, [x], [pk], F_.make_epilogue(), out_planes=False)[0]
        x = F_.lstm_bidir(prec, gates.f32, pk.w_hh, B, T, H, out_f32=not prec.uses_planes)
    return F_.linear(prec, [x], [F_.pack_of(m, "gloss", [m.gloss_layer])], F_.make_epilogue(clamp=clamp), out_planes=False)[0]


class AlignmentModule(nn.Module):
    def __init__(self, cls_num, input_size, hidden_size, num_layers=2, dropout=0.3, bidirectional=True):
        super().__init__()
        self.hidden_size = hidden_size
        self.num_layers = num_layers
        self.input_size = input_size
        self.bidirectional = bidirectional
        self.num_directions = 2 if bidirectional else 1
        self.lstm_hidden_size = int(hidden_size / self.num_directions)
        self.dropout = dropout
        if not bidirectional or self.lstm_hidden_size != 512:
            raise NotImplementedError(
                "scattennet_b200.AlignmentModule is built for the reference configuration: bidirectional, "
                f"512 hidden units per direction (got bidirectional={bidirectional}, hidden_size={hidden_size})")
        # parameter container only (identical state-dict keys); its cuDNN forward is never called
        self.rnn = nn.LSTM(input_size=input_size, hidden_size=self.lstm_hidden_size, num_layers=num_layers,
                           dropout=dropout, bidirectional=bidirectional)
        self.gloss_layer = nn.Linear(hidden_size, cls_num)
        self.precision = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """``x [T, B, input_size]`` (time-major, as the reference passes ``fuse_output.permute(1, 0, 2)``)
        -> logits ``[B, T, cls_num]``."""
        if self.training and self.dropout > 0 and self.num_layers > 1:
            raise RuntimeError("scattennet_b200 is inference-only: call .eval() before forward")
        F_.require_cuda(x)
        t, b, _ = x.shape
        prec = F_.get_precision(self.precision)
        rows = x.permute(1, 0, 2).reshape(b * t, -1)  # batch-major rows b*T + t
        out = alignment_forward(prec, self, Act.from_f32(rows), b, t)
        return out.f32.view(b, t, -1)
