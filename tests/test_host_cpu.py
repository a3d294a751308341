"""CPU-side checks: the C ABI library builds, loads and exports every symbol
include/scatt.h declares; state-dict layout equals the reference's; host logic
fails loudly without a GPU (no CPU fallback)."""

import json
import os
import re

import pytest
import torch

import scattennet_b200 as S
from scattennet_b200 import _lib, synth
from scattennet_b200.config import VOCAB_STUB, model_config, pooled_length

from helpers import reference_shapes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    header = open(os.path.join(ROOT, "include", "scatt.h")).read()
    declared = set(re.findall(r"\b(scatt_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.scatt_abi_version() == 3
    assert b"sm_100a" in lib.scatt_version()


def test_linear_ln_fused_query_is_host_logic():
    """scatt_linear_ln_fused mirrors the kernel choice of scatt_linear (no device needed): LayerNorm is fused for
    N = 256 always and for N = 512 / 1024 while the 4- / 8-CTA clusters of all problems fit one wave of 148 SMs."""
    lib = _lib.load()
    tc, simt = _lib.ENGINE_TCGEN05, _lib.ENGINE_SIMT
    assert lib.scatt_linear_ln_fused(1600, 256, 3, tc) == 1
    assert lib.scatt_linear_ln_fused(51200, 256, 3, tc) == 1
    assert lib.scatt_linear_ln_fused(800, 512, 3, tc) == 1      # 7 row tiles x 3 x 4 CTAs = 84
    assert lib.scatt_linear_ln_fused(400, 1024, 1, tc) == 1     # 4 x 8 = 32
    assert lib.scatt_linear_ln_fused(128 * 13, 512, 3, tc) == 0  # 156 CTAs > 148: GEMM + row-wise tail
    assert lib.scatt_linear_ln_fused(2400, 1024, 1, tc) == 0    # 19 x 8 = 152
    assert lib.scatt_linear_ln_fused(400, 384, 1, tc) == 0
    assert lib.scatt_linear_ln_fused(400, 256, 1, simt) == 0


def test_debug_setters_validate_their_argument():
    """Host logic of the schedule overrides (no device needed): out-of-range modes are refused with an error string."""
    lib = _lib.load()
    assert lib.scatt_debug_set_attn_persist(0) == 0 and lib.scatt_debug_set_attn_persist(2) == 0
    assert lib.scatt_debug_set_attn_persist(3) != 0 and b"debug_set_attn_persist" in lib.scatt_last_error()
    assert lib.scatt_debug_set_attn_persist(0) == 0
    assert lib.scatt_debug_set_block_cluster(3) != 0
    assert lib.scatt_debug_set_block_cluster(0) == 0


def test_ctypes_structs_match_header_sizes():
    import ctypes as C

    assert C.sizeof(_lib.Epilogue) == 32
    assert C.sizeof(_lib.LinearProblem) == 88
    assert C.sizeof(_lib.AttentionProblem) == 56
    assert C.sizeof(_lib.FrontendStream) == 8 + 4 + 8 + 4 + 7 * 16 + 8  # idx, n, coord[2], pad, 7 ptr pairs, gathered


@pytest.mark.parametrize("name", ["phoenix-2014t", "phoenix-2014"])
def test_state_dict_layout_equals_reference(name):
    ref = {k: v for k, v in reference_shapes(name).items() if "fuse_alignment_head" not in k}
    own = {k: tuple(v.shape) for k, v in S.MSCAEncoder(model_config(name), VOCAB_STUB).state_dict().items()}
    assert list(own.keys()) == list(ref.keys())  # same keys, same order
    assert own == ref


def test_load_reference_state_dict_ignores_out_of_path_keys():
    shapes = reference_shapes("phoenix-2014t")
    sd = synth.synth_state_dict(shapes, 0)
    m = S.MSCAEncoder(model_config("phoenix-2014t"), VOCAB_STUB)
    m.load_reference_state_dict(sd)
    assert torch.equal(m.body_encoder.sca.self_pos_embed.weight, sd["body_encoder.sca.self_pos_embed.weight"])
    del sd["left_encoder.residual.shortcuts.1.projection.weight"]
    with pytest.raises(KeyError):
        m.load_reference_state_dict(sd)


def test_attention_module_key_order_and_errors():
    a = S.SelfAttention(256, 16)
    assert [k.split(".")[0] for k in a.state_dict()][::2] == ["k_proj", "v_proj", "q_proj", "out_proj"]
    with pytest.raises(ValueError):
        S.SelfAttention(250, 16)
    with pytest.raises(ValueError):
        S.CoordinateAttention(model_config("phoenix-2014t"), attn_type="nope")


def test_no_cpu_fallback():
    m = S.MSCAEncoder(model_config("phoenix-2014t"), VOCAB_STUB).eval()
    kp, mask = synth.synth_batch(1, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(kp, mask)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        S.SelfAttention(256, 16).eval()(torch.zeros(1, 4, 256), torch.zeros(1, 1, 4, 4))


def test_training_mode_raises():
    m = S.MSCAEncoder(model_config("phoenix-2014t"), VOCAB_STUB)
    kp, mask = synth.synth_batch(1, 8)
    with pytest.raises(RuntimeError, match="inference-only"):
        m(kp, mask)


def test_synth_is_deterministic_and_config_helpers():
    a = synth.synth_tensor("x.attn.q_proj.weight", (4, 8), 3)
    b = synth.synth_tensor("x.attn.q_proj.weight", (4, 8), 3)
    assert torch.equal(a, b)
    assert not torch.equal(a, synth.synth_tensor("x.attn.k_proj.weight", (4, 8), 3))
    assert pooled_length(model_config("phoenix-2014t"), 200) == 50
    assert pooled_length(model_config("phoenix-2014"), 400) == 200
    assert pooled_length(model_config("phoenix-2014t"), 37) == 9
    kp, mask = synth.synth_batch(2, 6, lengths=[6, 2])
    assert kp.shape == (2, 6, 542, 2) and float(kp[1, 2:].abs().sum()) == 0.0
    assert mask.tolist() == [[1] * 6, [1, 1, 0, 0, 0, 0]]


def test_mask_helpers_match_reference_semantics():
    from scattennet_b200.utils import create_attention_mask, create_causal_attention_mask

    mask = torch.tensor([[1, 1, 0], [1, 0, 0]])
    m = create_attention_mask(mask, torch.float32)
    assert m.shape == (2, 1, 3, 3)
    assert float(m[0, 0, 1, 2]) == torch.finfo(torch.float32).min and float(m[0, 0, 2, 0]) == 0.0
    c = create_causal_attention_mask(mask, (2, 3), torch.zeros(2, 3, 4))
    assert float(c[0, 0, 1, 0]) == 1.0 and float(c[0, 0, 0, 1]) == 0.0
    assert float(c[0, 0, 2, 2]) == torch.finfo(torch.float32).min  # min + 1 rounds to min
