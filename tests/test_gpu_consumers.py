"""GPU parity of the path's consumers: the persistent BiLSTM recurrence, the
AlignmentModule mirror, the CTC log-prob front and the fused finiteness flag - against the oracle on seeded
inputs and against the reference's own outputs (tests/golden/align_*.npz, mod_alignment.npz)."""

import pytest
import torch

import scattennet_b200 as S
from oracle import scatt_oracle as O
from scattennet_b200 import functional as F_
from scattennet_b200 import synth
from scattennet_b200.config import VOCAB_STUB

from helpers import case_inputs, subsample

pytestmark = pytest.mark.gpu
DEV = "cuda"
TIERS = {"fp32": 1e-4, "fp16x3": 1e-4, "fp16x1": 1e-2}


def recurrence_cpu(gates_x, w_hh, B, T, H):
    """gates_x [B*T, 8H] -> [B*T, 2H]: the oracle's cell applied to precomputed input projections."""
    g = gates_x.view(B, T, 2, 4 * H)
    out = torch.empty(B, T, 2 * H)
    for d in range(2):
        h = torch.zeros(B, H)
        c = torch.zeros(B, H)
        for t in (range(T - 1, -1, -1) if d else range(T)):
            a = g[:, t, d] + h @ w_hh[d].T
            i, f, gg, o = a.split(H, dim=1)
            c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
            h = torch.sigmoid(o) * torch.tanh(c)
            out[:, t, d * H:(d + 1) * H] = h
    return out.view(B * T, 2 * H)


@pytest.mark.parametrize("B,T", [(1, 1), (1, 7), (3, 5), (8, 50), (11, 13), (20, 9), (8, 200)])
def test_lstm_bidir_op(B, T):
    H = 512
    g = torch.Generator().manual_seed(B * 1000 + T)
    gates = torch.randn(B * T, 8 * H, generator=g)
    w_hh = (torch.rand(2, 4 * H, H, generator=g) * 2 - 1) * 0.06
    want = recurrence_cpu(gates, w_hh, B, T, H)
    for mode in ("fp32", "fp16x3"):
        prec = F_.get_precision(mode)
        act = F_.lstm_bidir(prec, gates.to(DEV), w_hh.to(DEV), B, T, H)
        torch.cuda.synchronize()
        err = float((act.f32.cpu() - want).abs().max())
        assert err <= 2e-5, (mode, err)
        if prec.uses_planes:
            rebuilt = act.planes[0].float() + act.planes[1].float()
            assert float((rebuilt.cpu() - act.f32.cpu()).abs().max()) <= 1e-6


def test_lstm_bidir_rejects_unsupported():
    prec = F_.get_precision("fp32")
    with pytest.raises(Exception):
        F_.lstm_bidir(prec, torch.zeros(4, 8 * 256, device=DEV), torch.zeros(2, 1024, 256, device=DEV), 2, 2, 256)


@pytest.mark.parametrize("mode", list(TIERS))
def test_alignment_module_golden(golden, mode):
    arr, _ = golden("mod_alignment")
    m = S.AlignmentModule(cls_num=97, input_size=1024, hidden_size=1024).eval()
    synth.load_synth_(m, seed=16)
    m = m.to(DEV)
    m.precision = mode
    with torch.no_grad():
        for nm in "abc":
            out = m(arr[f"x_{nm}"].to(DEV)).cpu()
            assert out.shape == arr[f"logits_{nm}"].shape
            err = float((out - arr[f"logits_{nm}"]).abs().max())
            assert err <= TIERS[mode], (nm, mode, err)


def test_alignment_module_interface():
    with pytest.raises(NotImplementedError):
        S.AlignmentModule(cls_num=10, input_size=64, hidden_size=64)
    m = S.AlignmentModule(cls_num=10, input_size=1024, hidden_size=1024)
    ref = torch.nn.LSTM(1024, 512, num_layers=2, bidirectional=True)
    assert [k for k in m.state_dict()] == ["rnn." + k for k in ref.state_dict()] + ["gloss_layer.weight", "gloss_layer.bias"]
    with pytest.raises(RuntimeError):
        m.train()(torch.zeros(2, 1, 1024, device=DEV))
    with pytest.raises(RuntimeError):
        m.eval()(torch.zeros(2, 1, 1024))


@pytest.mark.parametrize("B,T,V", [(2, 9, 97), (8, 50, 1120), (3, 1, 1), (1, 4, 5000), (4, 7, 1536), (2, 3, 1540), (300, 50, 1120)])
def test_log_softmax_clamp(B, T, V):
    g = torch.Generator().manual_seed(V)
    x = torch.randn(B, T, V, generator=g) * 20
    x[0, 0, 0] = 200.0  # drives the other classes of that row below the -100 clamp
    want_tm = O.ctc_log_probs(x)
    got_tm = F_.log_softmax_clamp(x.to(DEV), time_major=True).cpu()
    got_bm = F_.log_softmax_clamp(x.to(DEV), time_major=False).cpu()
    assert got_tm.shape == (T, B, V) and got_bm.shape == (B, T, V)
    assert float((got_tm - want_tm).abs().max()) <= 2e-5
    assert torch.equal(got_bm.permute(1, 0, 2), got_tm)
    assert float(got_tm.min()) >= -100.0 and float(got_tm.max()) <= 0.0
    if V > 1:
        assert float(got_tm.min()) == -100.0


def test_finite_flags():
    a = torch.randn(1000, device=DEV)
    b = torch.randn(37, 5, device=DEV)
    c = torch.randn(3, device=DEV)
    assert int(F_.finite_flags([a, b, c]).item()) == 0
    b[36, 4] = float("nan")
    c[0] = float("-inf")
    flags = int(F_.finite_flags([a, b, c]).item())
    assert flags == O.non_finite_mask([a.cpu(), b.cpu(), c.cpu()]) == 0b110
    big = torch.zeros(3_000_001, device=DEV)
    big[-1] = float("inf")
    assert int(F_.finite_flags([big]).item()) == 1


@pytest.mark.parametrize("mode", list(TIERS))
@pytest.mark.parametrize("name", ["align_2014t_small", "align_2014t_odd", "align_2014_small", "align_2014t_c1"])
def test_encoder_alignment_golden(golden, name, mode):
    arr, meta = golden(name)
    cfg, sd, kp, mask = case_inputs(meta)
    m = S.MSCAEncoder(cfg, VOCAB_STUB, precision=mode, alignment=True).eval()
    m.load_reference_state_dict(sd)
    m = m.to(DEV)
    with torch.no_grad():
        out = m(kp.to(DEV), mask.to(DEV))
        torch.cuda.synchronize()
    out = subsample({k: v.cpu() for k, v in out.items() if k in arr}, meta)
    for k in ("alignment_gloss_logits", "fuse_coord_gloss_logits"):
        assert out[k].shape == arr[k].shape, k
        err = float((out[k] - arr[k]).abs().max())
        assert err <= TIERS[mode], (name, mode, k, err)


def test_encoder_alignment_graph_replay(golden):
    arr, meta = golden("align_2014t_odd")
    cfg, sd, kp, mask = case_inputs(meta)
    m = S.MSCAEncoder(cfg, VOCAB_STUB, precision="fp16x3", alignment=True).eval()
    m.load_reference_state_dict(sd)
    m = m.to(DEV)
    with torch.no_grad():
        eager = m(kp.to(DEV), mask.to(DEV))["alignment_gloss_logits"].clone()
        m.use_graph = True
        first = m(kp.to(DEV), mask.to(DEV))["alignment_gloss_logits"].clone()
        again = m(kp.to(DEV), mask.to(DEV))["alignment_gloss_logits"].clone()
    assert torch.equal(eager, first) and torch.equal(eager, again)
    assert float((eager.cpu() - arr["alignment_gloss_logits"]).abs().max()) <= 1e-4


def test_encoder_odd_vocabulary():
    """A gloss vocabulary that is not a multiple of 8 (the kernels' output-tile granularity) is zero-padded
    inside the packed weights; logits keep the exact [B, T', V] shape."""
    from scattennet_b200.config import model_config

    cfg = model_config("phoenix-2014t")
    V = 1115
    m = S.MSCAEncoder(cfg, V, precision="fp16x3", alignment=True).eval()
    synth.load_synth_(m, seed=5)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    kp, mask = synth.synth_batch(2, 24, seed=9, lengths=[24, 10])
    with torch.no_grad():
        want = O.encoder_forward(sd, cfg, kp, mask, alignment=True)
        got = m.to(DEV)(kp.to(DEV), mask.to(DEV))
    for k in ("left", "right", "body", "fuse_coord_gloss_logits", "alignment_gloss_logits"):
        assert got[k].shape == want[k].shape == (2, 6, V), k
        assert got[k].is_contiguous()
        assert float((got[k].cpu() - want[k]).abs().max()) <= 1e-4, k


def test_forward_check_finite():
    """The reference's NaN / inf guards (model/__init__.py:130-167) as one flag read-back."""
    from scattennet_b200.config import model_config

    cfg = model_config("phoenix-2014t")
    m = S.MSCAEncoder(cfg, 64, precision="fp16x3").eval()
    synth.load_synth_(m, seed=5)
    m = m.to(DEV)
    kp, mask = synth.synth_batch(2, 16, seed=3, lengths=[16, 9])
    with torch.no_grad():
        out = m(kp.to(DEV), mask.to(DEV), check_finite=True)
        assert set(out) >= {"body_embed", "fuse_embed", "fuse_coord_gloss_logits"}
        bad = kp.clone()
        bad[1, 2, 40, 0] = float("nan")  # a left-hand joint of a valid frame
        with pytest.raises(ValueError, match="input keypoints"):
            m(bad.to(DEV), mask.to(DEV), check_finite=True)
        m.coordinates_fusion.out_proj.bias[3] = float("inf")  # in-place under no_grad: bumps the version, weights repack
        with pytest.raises(ValueError, match="fuse_embed"):
            m(kp.to(DEV), mask.to(DEV), check_finite=True)


def test_peer_gather_two_gpus():
    """NVLink peer-memory logits gather (scatt_peer_allgather) against NCCL on two GPUs of one box: identical
    values over several calls (both buffer parities).  Skipped on single-GPU boxes."""
    import os
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29547", os.path.join(root, "tools", "test_peer_gather.py")],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "equal=True" in res.stdout
