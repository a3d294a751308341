"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container (the only place ``/root/reference`` exists):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these are outputs of the
reference's own modules (imported, never copied) on seeded synthetic inputs
with the deterministic synthetic weights of ``scattennet_b200.synth``.  Inputs
and weights are NOT stored (they are pure functions of the recorded seeds);
only outputs and checksums are, so the fixtures stay small.
"""

from __future__ import annotations

import copy
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("SCATT_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from scattennet_b200 import synth  # noqa: E402
from scattennet_b200.config import VOCAB_STUB, model_config  # noqa: E402

from model import MSCA_Net  # noqa: E402  (reference)
from model.attention import CrossAttention, SelfAttention, SelfCausalAttention  # noqa: E402
from model.encoder import Encoder  # noqa: E402
from model.fusion import CoordinatesFusion  # noqa: E402
from model.keypoint_module import SeparativeCoordinateAttention  # noqa: E402
from model.residual import ResidualNetwork  # noqa: E402
from model.utils import create_attention_mask, create_causal_attention_mask  # noqa: E402


class StubTokenizer:
    def __len__(self):
        return VOCAB_STUB


def full_cfg(name, **over):
    cfg = model_config(name, **over)
    # keys MSCA_Net.__init__ needs beyond the encoder path
    cfg.update(
        alignment_module=dict(input_size=1024, hidden_size=1024, num_layers=2, dropout=0.3, bidirectional=True),
        self_distillation=True,
        distillation_weight={"left": 1.0, "right": 1.0, "body": 1.0},
    )
    return cfg


def checksum(sd):
    return float(sum(v.double().abs().sum() for v in sd.values()))


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()})
    print(f"{name}: {os.path.getsize(path) / 1e6:.2f} MB")


@torch.no_grad()
def encoder_case(name, cfg_name, batch, t, lengths, seed_w=0, seed_in=1, frame_step=1, logit_step=(1, 1), mask_override=None, **over):
    cfg = full_cfg(cfg_name, **over)
    model = MSCA_Net(copy.deepcopy(cfg), StubTokenizer(), "cpu").eval()
    synth.load_synth_(model, seed=seed_w)
    kp, mask = synth.synth_batch(batch, t, seed=seed_in, lengths=lengths)
    if mask_override is not None:
        mask = torch.as_tensor(mask_override, dtype=torch.int64)
    body = model.body_encoder(kp[:, :, cfg["body_idx"], :], mask)
    left = model.left_encoder(kp[:, :, cfg["left_idx"], :], mask)
    right = model.right_encoder(kp[:, :, cfg["right_idx"], :], mask)
    fuse = model.coordinates_fusion(left, right, body)
    rh = model.recognition_head
    clamp = lambda z: torch.clamp(z, min=-50, max=50)
    logits = {
        "left": clamp(rh.left_gloss_classifier(left)),
        "right": clamp(rh.right_gloss_classifier(right)),
        "body": clamp(rh.body_gloss_classifier(body)),
        "fuse_coord_gloss_logits": clamp(rh.fuse_coord_classifier(fuse)),
    }
    fs, (lf, lv) = frame_step, logit_step
    save(
        name,
        meta=json.dumps(dict(cfg=cfg_name, over=over, batch=batch, t=t, lengths=lengths, seed_w=seed_w, seed_in=seed_in,
                             frame_step=fs, logit_step=[lf, lv], mask_override=mask_override)),
        weight_checksum=checksum(model.state_dict()),
        input_checksum=float(kp.double().sum()),
        body_embed=body[:, ::fs], left_embed=left[:, ::fs], right_embed=right[:, ::fs], fuse_embed=fuse[:, ::fs],
        **{k: v[:, ::lf, ::lv] for k, v in logits.items()},
    )


@torch.no_grad()
def module_cases():
    d, h, b, t = 256, 16, 2, 24
    g = torch.Generator().manual_seed(7)
    x = torch.randn(b, t, d, generator=g)
    kv = torch.randn(b, t, d, generator=g)
    mask = (torch.arange(t)[None] < torch.tensor([t, 15])[:, None]).long()
    hole = mask.clone()
    hole[0, 3] = 0
    hole[1, 0] = 0  # non-prefix masks
    out = {}
    for cls, nm in ((SelfAttention, "self"), (CrossAttention, "cross"), (SelfCausalAttention, "causal")):
        m = cls(d, h).eval()
        synth.load_synth_(m, seed=11)
        for mk_name, mk in (("prefix", mask), ("hole", hole)):
            if nm == "causal":
                add = create_causal_attention_mask(mk, (b, t), x)
                y = m(x, add)
            elif nm == "cross":
                add = create_attention_mask(mk, x.dtype, tgt_len=t)
                y = m(x, kv, add)
            else:
                add = create_attention_mask(mk, x.dtype)
                y = m(x, add)
            out[f"{nm}_{mk_name}"] = y
    # arbitrary dense additive mask through the low-level interface
    dense = torch.randn(b, 1, t, t, generator=g)
    m = SelfAttention(d, h).eval()
    synth.load_synth_(m, seed=11)
    out["self_dense"] = m(x, dense)
    save("mod_attention", x=x, kv=kv, mask=mask, hole=hole, dense=dense, **out)

    # SeparativeCoordinateAttention with maps, both self_attn_x settings, an all-padded row
    cfg = model_config("phoenix-2014t")
    xe = torch.randn(3, 13, d, generator=g)
    ye = torch.randn(3, 13, d, generator=g)
    m3 = (torch.arange(13)[None] < torch.tensor([13, 6, 0])[:, None]).long()
    res = {}
    for flag in (True, False):
        c = dict(cfg, self_attn_x=flag)
        s = SeparativeCoordinateAttention(c).eval()
        synth.load_synth_(s, seed=12)
        o = s(xe, ye, m3, return_attn_map=True)
        res[f"outputs_x{int(flag)}"] = o["outputs"]
        res[f"self_map_x{int(flag)}"] = o["self_attn_map"]
    save("mod_sca", x_embed=xe, y_embed=ye, mask=m3, **res)

    # ResidualNetwork: both YAML block lists, odd T, and the reference's own __main__ list
    res = {}
    for nm, blocks, tt in (("2014t", [256, 256, 512, 512], 21), ("2014", [256, 256], 21), ("main", [256, 256, 256], 9), ("t5", [256, 256, 512, 512], 5)):
        r = ResidualNetwork(blocks).eval()
        synth.load_synth_(r, seed=13)
        xi = torch.randn(2, tt, 256, generator=torch.Generator().manual_seed(100 + tt))
        y, outs = r(xi)
        res[f"x_{nm}"] = xi
        res[f"y_{nm}"] = y
        for i, o in enumerate(outs):
            res[f"y_{nm}_b{i}"] = o
    save("mod_residual", **res)

    # CoordinatesFusion incl. the reference's __main__ shape family
    f = CoordinatesFusion(512, 1024, 0.1).eval()
    synth.load_synth_(f, seed=14)
    l, r_, bd = (torch.rand(2, 7, 512, generator=g) * 3 for _ in range(3))
    save("mod_fusion", left=l, right=r_, body=bd, out=f(l, r_, bd))

    # generic Encoder (dead in the live model; interface kept)
    ecfg = dict(d_model=256, encoder_attention_heads=16, attention_dropout=0.0, dropout=0.1, activation_dropout=0.0,
                encoder_ffn_dim=768, encoder_layers=2, encoder_layerdrop=0.0, max_position_embeddings=64)
    e = Encoder(ecfg).eval()
    synth.load_synth_(e, seed=15)
    xe = torch.randn(2, 19, 256, generator=g)
    me = (torch.arange(19)[None] < torch.tensor([19, 8])[:, None]).long()
    save("mod_encoder", x=xe, mask=me, out=e(xe, me), cfg=json.dumps(ecfg))


@torch.no_grad()
def alignment_cases():
    """Consumers of the path: the reference's BiLSTM alignment head on the fused features
    of encoder cases above (same seeds), ``AlignmentModule`` alone, and the CTC log-prob front of ``compute_loss``."""
    from model.alignment_module import AlignmentModule  # noqa: E402  (reference)

    for name, cfg_name, batch, t, lengths, lstep, over in (
        ("align_2014t_small", "phoenix-2014t", 2, 16, [16, 11], (1, 1), {}),
        ("align_2014t_odd", "phoenix-2014t", 3, 37, [37, 20, 1], (1, 1), {}),
        ("align_2014_small", "phoenix-2014", 2, 18, [18, 7], (1, 1), {}),
        ("align_2014t_c1", "phoenix-2014t", 8, 200, synth.parity_lengths(8, 200), (1, 8), {}),
    ):
        cfg = full_cfg(cfg_name, **over)
        model = MSCA_Net(copy.deepcopy(cfg), StubTokenizer(), "cpu").eval()
        synth.load_synth_(model, seed=0)
        kp, mask = synth.synth_batch(batch, t, seed=1, lengths=lengths)
        body = model.body_encoder(kp[:, :, cfg["body_idx"], :], mask)
        left = model.left_encoder(kp[:, :, cfg["left_idx"], :], mask)
        right = model.right_encoder(kp[:, :, cfg["right_idx"], :], mask)
        fuse = model.coordinates_fusion(left, right, body)
        heads = model.recognition_head(left, right, fuse, body)  # the reference's own RecognitionHead.forward
        al = heads["alignment_gloss_logits"]
        lf, lv = lstep
        save(name,
             meta=json.dumps(dict(cfg=cfg_name, over=over, batch=batch, t=t, lengths=lengths, seed_w=0, seed_in=1,
                                  frame_step=1, logit_step=[lf, lv], mask_override=None)),
             alignment_gloss_logits=al[:, ::lf, ::lv],
             fuse_coord_gloss_logits=heads["fuse_coord_gloss_logits"][:, ::lf, ::lv])

    # AlignmentModule alone, time-major input as the reference passes it, plus compute_loss's log-prob front
    g = torch.Generator().manual_seed(21)
    m = AlignmentModule(cls_num=97, input_size=1024, hidden_size=1024).eval()
    synth.load_synth_(m, seed=16)
    res = {}
    for nm, (tt, bb) in (("a", (9, 2)), ("b", (23, 11)), ("c", (1, 3))):
        x = torch.randn(tt, bb, 1024, generator=g)
        lg = m(x)
        res[f"x_{nm}"] = x
        res[f"logits_{nm}"] = lg
        res[f"logp_{nm}"] = torch.clamp(torch.nn.functional.log_softmax(lg.permute(1, 0, 2), dim=-1), min=-100, max=0)
    save("mod_alignment", **res)


def state_dict_keys():
    for nm in ("phoenix-2014t", "phoenix-2014"):
        m = MSCA_Net(full_cfg(nm), StubTokenizer(), "cpu")
        keys = {k: list(v.shape) for k, v in m.state_dict().items()}
        with open(os.path.join(HERE, f"state_dict_{nm}.json"), "w") as fh:
            json.dump(keys, fh, indent=0)
        print(nm, len(keys), "keys")


if __name__ == "__main__":
    torch.set_num_threads(8)
    if "--alignment-only" in sys.argv:  # adds the section-8f fixtures without rewriting the encoder ones
        alignment_cases()
        sys.exit(0)
    state_dict_keys()
    module_cases()
    encoder_case("enc_2014t_small", "phoenix-2014t", 2, 16, [16, 11])
    encoder_case("enc_2014t_odd", "phoenix-2014t", 3, 37, [37, 20, 1])
    encoder_case("enc_2014t_allpad", "phoenix-2014t", 2, 12, [12, 0])
    encoder_case("enc_2014t_hole", "phoenix-2014t", 2, 10, None,
                 mask_override=[[1, 1, 0, 1, 1, 1, 0, 0, 1, 1], [0, 1, 1, 1, 1, 1, 1, 1, 0, 0]])
    encoder_case("enc_2014_small", "phoenix-2014", 2, 18, [18, 7])
    encoder_case("enc_2014t_c1", "phoenix-2014t", 8, 200, synth.parity_lengths(8, 200), logit_step=(1, 8))
    encoder_case("enc_2014_t400", "phoenix-2014", 2, 400, synth.parity_lengths(2, 400)[::-1], frame_step=4,
                 logit_step=(4, 8), max_position_embeddings=512)
    alignment_cases()
