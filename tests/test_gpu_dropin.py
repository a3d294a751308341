"""GPU tests of the drop-in surface around the kernels:

* the ``MSCA_Net.forward`` call pattern of INTEGRATION.md section 1 - three ``KeypointModule.forward(keypoints[:, :, idx, :],
  mask)`` calls and ``CoordinatesFusion.forward`` (reference ``model/__init__.py:86-97,133-154``) - against the
  reference's own outputs (``tests/golden/enc_2014t_c1.npz``);
* the whole encoder with the fused layer tail (``scatt_attn_block``) forced on at small batches;
* ``forward_host`` called back to back with different batches (double-buffered staging);
* captured graphs dropped when the weights change; input validation; the fp16-plane overflow policy.
"""

import pytest
import torch

import scattennet_b200 as S
from scattennet_b200 import functional as F_
from scattennet_b200 import synth
from scattennet_b200.config import VOCAB_STUB, model_config
from scattennet_b200.fusion import CoordinatesFusion
from scattennet_b200.keypoint_module import KeypointModule

from helpers import FEATURES, LOGITS, case_inputs, subsample

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sub(sd, prefix):
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


@pytest.mark.parametrize("mode,tol", [("fp16x3", 1e-4), ("fp16x1", 1e-2)])
def test_msca_net_call_pattern(golden, mode, tol):
    """What a maintainer gets after the import swap of INTEGRATION.md: the reference's constructor calls, its
    ``load_state_dict`` keys and its forward calls, module by module."""
    arr, meta = golden("enc_2014t_c1")
    cfg, sd, kp, mask = case_inputs(meta)
    F_.set_default_precision(mode)
    try:
        body_encoder = KeypointModule(cfg["body_idx"], num_frame=cfg["num_frame"], cfg=cfg).eval()
        left_encoder = KeypointModule(cfg["left_idx"], num_frame=cfg["num_frame"], cfg=cfg).eval()
        right_encoder = KeypointModule(cfg["right_idx"], num_frame=cfg["num_frame"], cfg=cfg).eval()
        coordinates_fusion = CoordinatesFusion(cfg["in_fusion_dim"], cfg["out_fusion_dim"], 0.2).eval()
        body_encoder.load_state_dict(_sub(sd, "body_encoder."), strict=True)
        left_encoder.load_state_dict(_sub(sd, "left_encoder."), strict=True)
        right_encoder.load_state_dict(_sub(sd, "right_encoder."), strict=True)
        coordinates_fusion.load_state_dict(_sub(sd, "coordinates_fusion."), strict=True)
        for m in (body_encoder, left_encoder, right_encoder, coordinates_fusion):
            m.to(DEV)
        keypoints, m_dev = kp.to(DEV), mask.to(DEV)
        with torch.no_grad():
            body = body_encoder(keypoints[:, :, cfg["body_idx"], :], m_dev)
            left = left_encoder(keypoints[:, :, cfg["left_idx"], :], m_dev)
            right = right_encoder(keypoints[:, :, cfg["right_idx"], :], m_dev)
            fuse = coordinates_fusion(left, right, body)
            torch.cuda.synchronize()
    finally:
        F_.set_default_precision("fp16x3")
    out = subsample({"body_embed": body.cpu(), "left_embed": left.cpu(), "right_embed": right.cpu(), "fuse_embed": fuse.cpu()}, meta)
    for k in FEATURES:
        assert out[k].shape == arr[k].shape, k
        assert float((out[k] - arr[k]).abs().max()) <= tol, (k, float((out[k] - arr[k]).abs().max()))


@pytest.mark.parametrize("mode,tol", [("fp16x3", 1e-4), ("fp16x1", 1e-2), ("bf16x3", 4e-4)])
@pytest.mark.parametrize("name", ["enc_2014t_c1", "enc_2014t_hole", "enc_2014t_allpad", "enc_2014_t400"])
def test_encoder_golden_with_fused_tail(golden, monkeypatch, name, mode, tol):
    """The small-batch goldens normally take the three-launch layer tail (few row tiles); force the fused kernel."""
    monkeypatch.setattr(F_, "FUSED_BLOCK_MIN_TILES", 0)
    arr, meta = golden(name)
    cfg, sd, kp, mask = case_inputs(meta)
    m = S.MSCAEncoder(cfg, VOCAB_STUB, precision=mode).eval()
    m.load_reference_state_dict(sd)
    m = m.to(DEV)
    with torch.no_grad():
        m(kp.to(DEV), mask.to(DEV))  # packs the weights (one split-plane launch per parameter)
        n0 = S._lib.launch_count()
        out = m(kp.to(DEV), mask.to(DEV))
        torch.cuda.synchronize()
    launches = S._lib.launch_count() - n0
    assert launches <= 62, launches  # 74+ without the fused tail: 8 layers x 2 launches fewer
    out = subsample({k: v.cpu() for k, v in out.items()}, meta)
    for k in FEATURES + LOGITS:
        assert torch.isfinite(out[k]).all(), k
        assert float((out[k] - arr[k]).abs().max()) <= tol, (name, mode, k, float((out[k] - arr[k]).abs().max()))


@pytest.mark.parametrize("use_graph", [True, False])
def test_forward_host_back_to_back(use_graph):
    """Two (then four) calls with different batches and no synchronisation in between: the pinned staging is
    double-buffered with an event per slot, so every call's result is its own batch's."""
    cfg = model_config("phoenix-2014t")
    m = S.MSCAEncoder(cfg, VOCAB_STUB, precision="fp16x3", use_graph=use_graph).eval()
    synth.load_synth_(m, 0)
    m = m.to(DEV)
    batches = [synth.synth_batch(3, 48, seed=100 + i, lengths=[48, 40 - i, 17 + i]) for i in range(4)]
    with torch.no_grad():
        want = [m(kp.to(DEV), mask.to(DEV))["fuse_coord_gloss_logits"].clone().cpu() for kp, mask in batches]
        torch.cuda.synchronize()
        r0 = m.forward_host(*batches[0])
        r1 = m.forward_host(*batches[1])
        torch.cuda.synchronize()
        assert torch.equal(r0["fuse_coord_gloss_logits"], want[0])
        assert torch.equal(r1["fuse_coord_gloss_logits"], want[1])
        got = []
        for kp, mask in batches:  # results are valid until the call two steps later: copy each before that
            r = m.forward_host(kp, mask)
            if len(got) >= 1:
                torch.cuda.current_stream().synchronize()
            got.append(r)
        torch.cuda.synchronize()
        assert torch.equal(got[2]["fuse_coord_gloss_logits"], want[2])
        assert torch.equal(got[3]["fuse_coord_gloss_logits"], want[3])


def test_graphs_dropped_when_weights_change():
    cfg = model_config("phoenix-2014t")
    kp, mask = synth.synth_batch(2, 32, seed=5, lengths=[32, 20])
    kp, mask = kp.to(DEV), mask.to(DEV)
    m = S.MSCAEncoder(cfg, VOCAB_STUB, precision="fp16x3", use_graph=True).eval()
    synth.load_synth_(m, 0)
    m = m.to(DEV)
    with torch.no_grad():
        a = {k: v.clone() for k, v in m(kp, mask).items()}
        assert len(m._graphs) == 1
        other = S.MSCAEncoder(cfg, VOCAB_STUB, precision="fp16x3").eval()
        synth.load_synth_(other, 3)
        m.load_state_dict(other.state_dict())
        assert len(m._graphs) == 0  # stale packed-weight pointers must not be replayed
        b = m(kp, mask)
        ref = other.to(DEV)(kp, mask)
        torch.cuda.synchronize()
    for k in b:
        assert torch.equal(b[k], ref[k]), k
    assert not torch.equal(a["fuse_embed"], b["fuse_embed"])
    for i in range(12):  # the cache is bounded (variable-length inference)
        kp_i, mask_i = synth.synth_batch(1, 16 + 4 * i, seed=1)
        with torch.no_grad():
            m(kp_i.to(DEV), mask_i.to(DEV))
    assert len(m._graphs) <= m.max_cached_shapes


def test_graph_reads_repeated_inputs_in_place():
    """The same input tensors on consecutive calls: from the second call on a graph bound to them replays without the
    device-to-device input copies; refilling them in place is seen by the next call, other tensors take the copying
    graph, and every route returns what the eager path returns."""
    cfg = model_config("phoenix-2014t")
    m = S.MSCAEncoder(cfg, VOCAB_STUB, precision="fp16x3", use_graph=True).eval()
    synth.load_synth_(m, 0)
    m = m.to(DEV)
    eager = S.MSCAEncoder(cfg, VOCAB_STUB, precision="fp16x3", use_graph=False).eval()
    synth.load_synth_(eager, 0)
    eager = eager.to(DEV)
    batches = [synth.synth_batch(2, 48, seed=s, lengths=[48, 30 + s]) for s in (1, 2, 3)]
    kp, mask = batches[0][0].to(DEV), batches[0][1].to(DEV)
    with torch.no_grad():
        for step, (kp_h, mask_h) in enumerate(batches + batches[:1]):
            kp.copy_(kp_h.to(DEV))      # refill the SAME device tensors
            mask.copy_(mask_h.to(DEV))
            got = {k: v.clone() for k, v in m(kp, mask).items()}
            want = eager(kp, mask)
            torch.cuda.synchronize()
            for k in want:
                assert torch.equal(got[k], want[k]), (step, k)
            bound = next(iter(m._graphs.values()))[6]
            assert (bound is not None) == (step >= 1)
        other_kp, other_mask = batches[1][0].to(DEV), batches[1][1].to(DEV)  # different storage: the copying graph
        got = {k: v.clone() for k, v in m(other_kp, other_mask).items()}
        want = eager(other_kp, other_mask)
        torch.cuda.synchronize()
        for k in want:
            assert torch.equal(got[k], want[k]), k


def test_input_validation():
    cfg = model_config("phoenix-2014t")
    m = S.MSCAEncoder(cfg, VOCAB_STUB).eval().to(DEV)
    kp, mask = synth.synth_batch(2, 16, seed=1)
    with pytest.raises(IndexError):  # a 133-joint tensor with the 542-joint config (the reference's gather raises too)
        m(kp[:, :, :70].contiguous().to(DEV), mask.to(DEV))
    with pytest.raises(ValueError):
        m(kp.to(DEV), mask[:, :8].to(DEV))
    with pytest.raises(ValueError):
        m(kp[..., 0].to(DEV), mask.to(DEV))
    with pytest.raises(IndexError):
        m.forward_host(kp[:, :, :70].contiguous(), mask)


def test_fp16_plane_overflow_policy():
    """fp16 planes hold |x| <= 65504.  Activations beyond that (possible with a trained checkpoint's un-normalised
    fusion activations; never with the synthetic Xavier weights) become inf / NaN - they are REPORTED, not hidden:
    ``check_finite=True`` raises, and the ``bf16x3`` mode (8 exponent bits) computes the same batch finitely."""
    cfg = model_config("phoenix-2014t")
    kp, mask = synth.synth_batch(2, 16, seed=1)
    outs = {}
    for mode in ("fp16x3", "bf16x3"):
        m = S.MSCAEncoder(cfg, VOCAB_STUB, precision=mode).eval()
        synth.load_synth_(m, 0)
        with torch.no_grad():
            m.coordinates_fusion.body_se.weight.mul_(3e5)  # GELU(body_se(x)) ~ 1e5..1e6 feeds the fusion attention as V
        m = m.to(DEV)
        with torch.no_grad():
            if mode == "fp16x3":
                with pytest.raises(ValueError, match="NaN or inf"):
                    m(kp.to(DEV), mask.to(DEV), check_finite=True)
            else:
                outs[mode] = m(kp.to(DEV), mask.to(DEV), check_finite=True)
    assert torch.isfinite(outs["bf16x3"]["fuse_embed"]).all()
