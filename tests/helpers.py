"""Shared test helpers (CPU side)."""

import json
import os

import torch

from scattennet_b200 import synth
from scattennet_b200.config import model_config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

ENCODER_CASES = ["enc_2014t_small", "enc_2014t_odd", "enc_2014t_allpad", "enc_2014t_hole", "enc_2014_small",
                 "enc_2014t_c1", "enc_2014_t400"]
FEATURES = ["body_embed", "left_embed", "right_embed", "fuse_embed"]
LOGITS = ["left", "right", "body", "fuse_coord_gloss_logits"]


def reference_shapes(cfg_name):
    """``{key: shape}`` of the reference ``MSCA_Net.state_dict()`` (committed fixture)."""
    with open(os.path.join(GOLDEN, f"state_dict_{cfg_name}.json")) as fh:
        return {k: tuple(v) for k, v in json.load(fh).items()}


def case_inputs(meta):
    """Rebuild ``(cfg, state_dict, keypoints, mask)`` of an encoder golden case from its recorded seeds."""
    m = meta["meta"]
    cfg = model_config(m["cfg"], **m["over"])
    shapes = reference_shapes(m["cfg"])
    if "max_position_embeddings" in m["over"]:
        rows = m["over"]["max_position_embeddings"] + 2
        shapes = {k: ((rows, v[1]) if k.endswith("pos_embed.weight") else v) for k, v in shapes.items()}
    sd = synth.synth_state_dict(shapes, m["seed_w"])
    kp, mask = synth.synth_batch(m["batch"], m["t"], m["seed_in"], m["lengths"])
    if m["mask_override"] is not None:
        mask = torch.tensor(m["mask_override"], dtype=torch.int64)
    return cfg, sd, kp, mask


def subsample(out, meta):
    """Apply the frame / vocab strides the fixture was stored with."""
    m = meta["meta"]
    fs, (lf, lv) = m["frame_step"], m["logit_step"]
    res = {}
    for k, v in out.items():
        res[k] = v[:, ::fs] if k in FEATURES else v[:, ::lf, ::lv]
    return res


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))
