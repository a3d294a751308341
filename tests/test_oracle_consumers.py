"""Pin the oracle's restatement of the path's consumers (BiLSTM alignment head, CTC
log-prob front) against outputs of the reference itself (tests/golden/align_*.npz, mod_alignment.npz)."""

import pytest
import torch

from oracle import scatt_oracle as O
from scattennet_b200 import synth

from helpers import case_inputs, subsample

TOL = 2e-5


def alignment_sd(cls_num=97, seed=16):
    from scattennet_b200.alignment_module import AlignmentModule

    shapes = {k: tuple(v.shape) for k, v in AlignmentModule(cls_num, 1024, 1024).state_dict().items()}
    return {"a." + k: v for k, v in synth.synth_state_dict(shapes, seed).items()}


def test_alignment_module_and_log_probs(golden):
    arr, _ = golden("mod_alignment")
    sd = alignment_sd()
    with torch.no_grad():
        for nm in "abc":
            lg = O.alignment_module(sd, "a", arr[f"x_{nm}"])
            assert lg.shape == arr[f"logits_{nm}"].shape
            assert float((lg - arr[f"logits_{nm}"]).abs().max()) <= TOL, nm
            lp = O.ctc_log_probs(arr[f"logits_{nm}"])
            assert lp.shape == arr[f"logp_{nm}"].shape
            assert float((lp - arr[f"logp_{nm}"]).abs().max()) <= TOL, nm


def test_lstm_direction_equals_torch_lstm():
    """The written-out recurrence against ATen's own nn.LSTM (what the reference calls)."""
    torch.manual_seed(3)
    rnn = torch.nn.LSTM(24, 16, num_layers=1, bidirectional=True).eval()
    x = torch.randn(7, 3, 24)
    with torch.no_grad():
        want, _ = rnn(x)
        fwd = O.lstm_direction(rnn.weight_ih_l0, rnn.weight_hh_l0, rnn.bias_ih_l0, rnn.bias_hh_l0, x, False)
        bwd = O.lstm_direction(rnn.weight_ih_l0_reverse, rnn.weight_hh_l0_reverse, rnn.bias_ih_l0_reverse,
                               rnn.bias_hh_l0_reverse, x, True)
    assert float((torch.cat([fwd, bwd], 2) - want).abs().max()) <= 1e-6


@pytest.mark.parametrize("name", ["align_2014t_small", "align_2014t_odd", "align_2014_small"])
def test_encoder_alignment_cases(golden, name):
    arr, meta = golden(name)
    cfg, sd, kp, mask = case_inputs(meta)
    with torch.no_grad():
        out = subsample(O.encoder_forward(sd, cfg, kp, mask, alignment=True), meta)
    for k in ("alignment_gloss_logits", "fuse_coord_gloss_logits"):
        assert out[k].shape == arr[k].shape, k
        assert float((out[k] - arr[k]).abs().max()) <= 5e-5, k


def test_non_finite_mask():
    a, b, c = torch.zeros(5), torch.tensor([1.0, float("nan")]), torch.tensor([float("-inf")])
    assert O.non_finite_mask([a, b, c]) == 0b110
    assert O.non_finite_mask([a]) == 0
