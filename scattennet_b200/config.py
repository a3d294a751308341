"""Shape contract of the SCAttenNet encoder (the two shipped YAMLs, restated).

Only the keys the encoder path consumes are kept.  Values follow
``configs/phoenix-2014t.yaml:206-277`` and ``configs/phoenix-2014.yaml:207-278``
of the reference (the two files differ only in ``residual_blocks`` and
``in_fusion_dim``).  ``num_frame`` is stored by ``KeypointModule`` and never
read (reference ``model/keypoint_module.py:17``).
"""

from __future__ import annotations

import copy

NUM_KEYPOINTS = 542  # K = max joint index + 1 of the collated batch (reference main.py:140-154)
VOCAB_STUB = 1120  # classifier width used for synthetic runs

_COMMON = {
    "attention_dropout": 0.0,
    "d_model": 256,
    "dropout": 0.2,
    "attention_heads": 16,
    "ff_dim": 768,
    "attn_layers": 4,
    "num_frame": 180,
    "max_position_embeddings": 256,
    "out_fusion_dim": 1024,
    "body_idx": list(range(11, 17)),
    "right_idx": list(range(54, 75)),
    "left_idx": list(range(33, 54)),
    # the BiLSTM alignment head that consumes fuse_embed (configs/phoenix-2014t.yaml:220-225)
    "alignment_module": {"input_size": 1024, "hidden_size": 1024, "num_layers": 2, "dropout": 0.3, "bidirectional": True},
}

PHOENIX_2014T = dict(_COMMON, residual_blocks=[256, 256, 512, 512], in_fusion_dim=512)
PHOENIX_2014 = dict(_COMMON, residual_blocks=[256, 256], in_fusion_dim=256)

_BY_NAME = {"phoenix-2014t": PHOENIX_2014T, "phoenix-2014": PHOENIX_2014}


def model_config(name: str, **overrides) -> dict:
    """Return a private copy of a named model config with overrides applied.

    ``max_position_embeddings`` must be raised (512) for the T=400 runs: the
    shipped value 256 makes both the reference and this package raise
    ``IndexError`` for T > 256 (reference ``model/layers.py:17-28``).
    """
    if name not in _BY_NAME:
        raise KeyError(f"unknown config {name!r}; have {sorted(_BY_NAME)}")
    cfg = copy.deepcopy(_BY_NAME[name])
    cfg.update(overrides)
    return cfg


def load_yaml_model_config(path: str) -> dict:
    """Read the ``model:`` section of a reference YAML (same keys as above)."""
    import yaml

    with open(path) as fh:
        return yaml.safe_load(fh)["model"]


def pooled_length(cfg: dict, t: int) -> int:
    """Frames left after the residual network (MaxPool1d(2,2) on even blocks,
    reference ``model/residual.py:57-61,40-43``)."""
    for i in range(len(cfg["residual_blocks"])):
        if i % 2 == 0:
            t //= 2
    return t
