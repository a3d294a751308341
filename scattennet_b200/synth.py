"""Deterministic synthetic weights and keypoint batches.

There is no pretrained SCAttenNet checkpoint (reference ``README.md:20-21``
links are placeholders) and no dataset in the sandbox, so parity tests, golden
fixtures and the bench all use weights and inputs generated here.  Every
tensor is a pure function of ``(seed, parameter name, shape)`` so the reference
model (in the container that has ``/root/reference``) and this package (on the
GPU box that does not) can be given bit-identical parameters without shipping
a 200 MB state dict.

The value distributions follow the reference's own initialisation
(``model/__init__.py:108-117``: Xavier-uniform ``nn.Linear`` weights,
``nn.Embedding`` N(0,1) position tables) except that biases and LayerNorm
affine parameters are perturbed away from (0, 1, 0) when ``style="perturbed"``
so that every bias / affine code path is exercised by parity tests.
"""

from __future__ import annotations

import math
import zlib
from typing import Iterable, Mapping, Sequence

import torch

from .config import NUM_KEYPOINTS

_NORM_TOKENS = ("norm", ".bn1.")


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 0x9E3779B1 + zlib.crc32(name.encode())) % (1 << 62))
    return g


def synth_tensor(name: str, shape: Sequence[int], seed: int = 0, style: str = "perturbed") -> torch.Tensor:
    """One fp32 parameter tensor, a pure function of (seed, name, shape, style)."""
    if style not in ("perturbed", "reference"):
        raise ValueError(style)
    shape = tuple(int(s) for s in shape)
    g = _gen(seed, name)
    leaf = name.rsplit(".", 1)[-1]
    is_norm = any(tok in "." + name for tok in _NORM_TOKENS) and len(shape) == 1
    if name.endswith("pos_embed.weight") or name.endswith("embed_positions.weight"):
        return torch.randn(shape, generator=g, dtype=torch.float32)
    if is_norm:
        if style == "reference":
            return torch.ones(shape) if leaf == "weight" else torch.zeros(shape)
        u = torch.rand(shape, generator=g, dtype=torch.float32)
        return 1.0 + 0.2 * (u - 0.5) if leaf == "weight" else 0.1 * (u - 0.5)
    if len(shape) == 2:
        fan_out, fan_in = shape
        bound = math.sqrt(6.0 / (fan_in + fan_out))
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound
    if len(shape) == 1:
        if style == "reference":
            return torch.zeros(shape)
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * 0.05
    raise ValueError(f"no synthetic rule for {name} {shape}")


def synth_state_dict(shapes: Mapping[str, Sequence[int]], seed: int = 0, style: str = "perturbed") -> dict:
    """Synthetic state dict for any ``{name: shape}`` map (e.g. ``{k: v.shape for k, v in m.state_dict().items()}``)."""
    return {name: synth_tensor(name, shape, seed, style) for name, shape in shapes.items()}


def load_synth_(module: torch.nn.Module, seed: int = 0, style: str = "perturbed", strict: bool = True):
    """Overwrite every parameter of ``module`` with its synthetic value (in place)."""
    sd = synth_state_dict({k: tuple(v.shape) for k, v in module.state_dict().items()}, seed, style)
    module.load_state_dict(sd, strict=strict)
    return module


def parity_lengths(batch: int, t: int) -> list:
    """Ragged right-padded lengths of the parity runs:
    ``[200,187,160,200,133,96,200,64] * T/200`` cycled over the batch."""
    base = [200, 187, 160, 200, 133, 96, 200, 64]
    return [max(1, (base[i % len(base)] * t) // 200) for i in range(batch)]


def synth_batch(batch: int, t: int, seed: int = 1, lengths: Iterable[int] | None = None, num_keypoints: int = NUM_KEYPOINTS):
    """``(keypoints [B,T,K,2] fp32 in [0,1), mask [B,T] int64)``.

    Padded frames are zeroed like the reference collator does
    (``dataset.py:80-96``); ``lengths=None`` gives full-length masks
    (throughput runs)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    kp = torch.rand(batch, t, num_keypoints, 2, generator=g, dtype=torch.float32)
    if lengths is None:
        mask = torch.ones(batch, t, dtype=torch.int64)
    else:
        lens = torch.as_tensor(list(lengths), dtype=torch.int64)
        if lens.numel() != batch:
            raise ValueError("one length per sequence")
        mask = (torch.arange(t)[None, :] < lens[:, None]).to(torch.int64)
        kp = kp * mask[:, :, None, None].to(kp.dtype)
    return kp, mask
