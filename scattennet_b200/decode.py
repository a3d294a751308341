"""CTC decode of the per-frame logits on the GPU.

``ctc_decode`` keeps the call of the reference's ``utils.ctc_decode(gloss_logits, beam_size, input_lengths)``
(``utils.py:164-189``) and its return type - one list of gloss ids per sequence - but the beam search runs in a
CUDA kernel (``scatt_ctc_beam_decode``) instead of TensorFlow on the host, so ``B * T' * V`` logits stay on the
device and a few hundred token ids come back.
"""

from __future__ import annotations

from typing import List

import torch

from . import functional as F_


def ctc_decode(gloss_logits: torch.Tensor, beam_size: int, input_lengths: torch.Tensor) -> List[List[int]]:
    """``gloss_logits [B,T',V]`` (class 0 = blank), ``input_lengths [B]`` -> decoded gloss-id sequences."""
    ids, n_ids, _ = F_.ctc_beam_decode(gloss_logits, input_lengths, beam_size)
    ids, n_ids = ids.cpu(), n_ids.cpu()  # the one synchronising read-back: token ids only
    return [ids[b, : int(n_ids[b])].tolist() for b in range(ids.shape[0])]
