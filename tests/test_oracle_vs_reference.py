"""Live re-check of the oracle against the UNMODIFIED reference (``/root/reference``), wherever that tree exists
(the build container; skipped on the GPU box).  The committed fixtures (``tests/golden``) pin the same thing
offline; this test guards against a stale fixture / oracle pair."""

import copy
import os
import sys

import pytest
import torch

from conftest import REFERENCE
from oracle import scatt_oracle as O
from scattennet_b200 import synth
from scattennet_b200.config import VOCAB_STUB, model_config

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "model")), reason="reference tree not present")


class _StubTokenizer:
    def __len__(self):
        return VOCAB_STUB


@pytest.mark.parametrize("cfg_name,batch,t,lengths", [("phoenix-2014t", 2, 24, [24, 13]), ("phoenix-2014", 2, 18, [18, 7])])
def test_oracle_equals_reference(cfg_name, batch, t, lengths):
    sys.path.insert(0, REFERENCE)
    try:
        from model import MSCA_Net  # the reference, imported - never copied
    finally:
        sys.path.remove(REFERENCE)
    cfg = model_config(cfg_name)
    full = dict(cfg)
    full.update(alignment_module=dict(input_size=1024, hidden_size=1024, num_layers=2, dropout=0.3, bidirectional=True),
                self_distillation=True, distillation_weight={"left": 1.0, "right": 1.0, "body": 1.0})
    torch.set_num_threads(min(8, os.cpu_count() or 1))
    model = MSCA_Net(copy.deepcopy(full), _StubTokenizer(), "cpu").eval()
    synth.load_synth_(model, seed=4)
    kp, mask = synth.synth_batch(batch, t, seed=2, lengths=lengths)
    with torch.no_grad():
        body = model.body_encoder(kp[:, :, cfg["body_idx"], :], mask)
        left = model.left_encoder(kp[:, :, cfg["left_idx"], :], mask)
        right = model.right_encoder(kp[:, :, cfg["right_idx"], :], mask)
        fuse = model.coordinates_fusion(left, right, body)
        logits = torch.clamp(model.recognition_head.fuse_coord_classifier(fuse), min=-50, max=50)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        out = O.encoder_forward(sd, cfg, kp, mask)
    want = {"body_embed": body, "left_embed": left, "right_embed": right, "fuse_embed": fuse, "fuse_coord_gloss_logits": logits}
    for k, v in want.items():
        assert float((out[k] - v).abs().max()) <= 2e-5, k
