"""Mask helpers with the interface of the reference ``model/utils.py``.

They exist for callers that drive the low-level attention classes directly;
the containers in this package never call them on the hot path (the kernels
apply the same semantics from the ``[B,T]`` key mask without materialising a
``[B,1,T,T]`` tensor).  Pure tensor construction - no kernels needed.
"""

import torch


def create_attention_mask(mask, dtype, tgt_len=None):
    """``[B,1,tgt,src]`` additive mask: 0 for a valid key, ``finfo(dtype).min``
    for a padded key (reference ``model/utils.py:3-12``)."""
    bsz, src_len = mask.size()
    tgt_len = tgt_len if tgt_len is not None else src_len
    padded = (mask[:, None, None, :].expand(bsz, 1, tgt_len, src_len) == 0)
    return torch.zeros(bsz, 1, tgt_len, src_len, dtype=dtype, device=mask.device).masked_fill(padded, torch.finfo(dtype).min)


def create_causal_attention_mask(attention_mask, input_shape, inputs_embeds):
    """Key-padding mask plus +1.0 on the lower triangle, as the reference builds
    it (``model/utils.py:15-28``); the real causal cut is made inside
    ``SelfCausalAttention``."""
    bsz, q_len = input_shape[0], input_shape[1]
    out = create_attention_mask(attention_mask[:, :q_len], inputs_embeds.dtype, tgt_len=q_len)
    tri = torch.tril(torch.ones((q_len, q_len), device=inputs_embeds.device, dtype=inputs_embeds.dtype))
    return out + tri[None, None, :, :]
