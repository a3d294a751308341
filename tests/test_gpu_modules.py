"""GPU parity of the nn.Module mirror against the reference's own outputs
(tests/golden/mod_*.npz): construct our module, load the same synthetic
weights, compare.  Tolerances are the north-star tiers: 1e-4 for the fp32 tier
(`fp32` SIMT engine and `fp16x3` tensor-core engine), 1e-2 for the 16-bit tier
(`fp16x1`)."""

import pytest
import torch

import scattennet_b200 as S
from scattennet_b200 import synth
from scattennet_b200.config import model_config
from scattennet_b200.utils import create_attention_mask, create_causal_attention_mask

pytestmark = pytest.mark.gpu
DEV = "cuda"
TIERS = {"fp32": 1e-4, "fp16x3": 1e-4, "fp16x1": 1e-2}


def build(cls, seed, *args, **kw):
    m = cls(*args, **kw).eval()
    synth.load_synth_(m, seed)
    return m.to(DEV)


def maxerr(a, b):
    return float((a.detach().cpu() - b).abs().max())


@pytest.mark.parametrize("mode", list(TIERS))
def test_attention_modules(golden, mode):
    S.set_default_precision(mode)
    arr, _ = golden("mod_attention")
    x, kv = arr["x"].to(DEV), arr["kv"].to(DEV)
    b, t, _ = x.shape
    for mk in ("prefix", "hole"):
        m = arr["mask" if mk == "prefix" else "hole"].to(DEV)
        pad = create_attention_mask(m, torch.float32)
        assert maxerr(build(S.SelfAttention, 11, 256, 16)(x, pad), arr[f"self_{mk}"]) <= TIERS[mode]
        assert maxerr(build(S.CrossAttention, 11, 256, 16)(x, kv, create_attention_mask(m, torch.float32, tgt_len=t)),
                      arr[f"cross_{mk}"]) <= TIERS[mode]
        cm = create_causal_attention_mask(m, (b, t), x)
        assert maxerr(build(S.SelfCausalAttention, 11, 256, 16)(x, cm), arr[f"causal_{mk}"]) <= TIERS[mode]
    assert maxerr(build(S.SelfAttention, 11, 256, 16)(x, arr["dense"].to(DEV)), arr["self_dense"]) <= TIERS[mode]
    S.set_default_precision("fp16x3")


@pytest.mark.parametrize("mode", list(TIERS))
def test_sca_module(golden, mode):
    S.set_default_precision(mode)
    arr, _ = golden("mod_sca")
    for flag in (True, False):
        m = build(S.SeparativeCoordinateAttention, 12, model_config("phoenix-2014t", self_attn_x=flag))
        o = m(arr["x_embed"].to(DEV), arr["y_embed"].to(DEV), arr["mask"].to(DEV), return_attn_map=True)
        assert set(o) == {"outputs", "self_attn_map", "causal_attn_map"}
        assert maxerr(o["outputs"], arr[f"outputs_x{int(flag)}"]) <= TIERS[mode]
        assert maxerr(o["self_attn_map"], arr[f"self_map_x{int(flag)}"]) <= TIERS[mode]
    S.set_default_precision("fp16x3")


@pytest.mark.parametrize("mode", list(TIERS))
def test_residual_network(golden, mode):
    S.set_default_precision(mode)
    arr, _ = golden("mod_residual")
    for nm, blocks in (("2014t", [256, 256, 512, 512]), ("2014", [256, 256]), ("main", [256, 256, 256]), ("t5", [256, 256, 512, 512])):
        m = build(S.ResidualNetwork, 13, blocks)
        y, outs = m(arr[f"x_{nm}"].to(DEV))
        assert y.shape == arr[f"y_{nm}"].shape
        assert maxerr(y, arr[f"y_{nm}"]) <= TIERS[mode]
        assert len(outs) == len(blocks)
        for i, o in enumerate(outs):
            assert maxerr(o, arr[f"y_{nm}_b{i}"]) <= TIERS[mode]
    # a single frame cannot be pooled: the reference raises, so do we
    with pytest.raises(RuntimeError):
        build(S.ResidualNetwork, 13, [256, 256])(torch.zeros(1, 1, 256, device=DEV))
    S.set_default_precision("fp16x3")


@pytest.mark.parametrize("mode", list(TIERS))
def test_fusion_and_encoder_modules(golden, mode):
    S.set_default_precision(mode)
    arr, _ = golden("mod_fusion")
    m = build(S.CoordinatesFusion, 14, 512, 1024, 0.1)
    out = m(arr["left"].to(DEV), arr["right"].to(DEV), arr["body"].to(DEV))
    assert maxerr(out, arr["out"]) <= TIERS[mode] * (3 if mode == "fp16x1" else 1)  # |logit| ~ 2e2 in this fixture
    arr, meta = golden("mod_encoder")
    e = build(S.Encoder, 15, meta["cfg"])
    assert maxerr(e(arr["x"].to(DEV), arr["mask"].to(DEV)), arr["out"]) <= TIERS[mode]
    with pytest.raises(IndexError):
        e(torch.zeros(1, 65, 256, device=DEV), torch.ones(1, 65, device=DEV))
    S.set_default_precision("fp16x3")


def test_standalone_layers_match_torch():
    S.set_default_precision("fp32")
    ff = build(S.FeedForward, 21, 256, 768, 0.2)
    x = torch.randn(2, 9, 256, generator=torch.Generator().manual_seed(1)).to(DEV)
    ref = torch.nn.functional.linear(torch.nn.functional.gelu(torch.nn.functional.linear(x, ff.fc1.weight, ff.fc1.bias)), ff.fc2.weight, ff.fc2.bias)
    assert maxerr(ff(x), ref.cpu()) <= 1e-5
    cm = build(S.CoordinateMapping, 22, 21, 256)
    xc, yc = torch.rand(2, 9, 21, device=DEV), torch.rand(2, 9, 21, device=DEV)
    xe, ye = cm(xc, yc)
    assert maxerr(xe, torch.nn.functional.linear(xc, cm.mapping_x.weight, cm.mapping_x.bias).cpu()) <= 1e-5
    assert maxerr(ye, torch.nn.functional.linear(yc, cm.mapping_y.weight, cm.mapping_y.bias).cpu()) <= 1e-5
    pe = build(S.LearningPositionEmbedding, 23, 64, 256)
    assert maxerr(pe(x), (x + pe.weight[2:11][None]).cpu()) == 0.0
    S.set_default_precision("fp16x3")
