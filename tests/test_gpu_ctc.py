"""GPU parity of the CTC beam-search kernel (scatt_ctc_beam_decode) against oracle/ctc_oracle.py, which restates
the reference's utils.ctc_decode (TensorFlow's ctc_beam_search_decoder + the reference's pre / post-processing) and
is itself pinned against brute-force enumeration (tests/test_oracle_ctc.py).  Token ids must match exactly."""

import numpy as np
import pytest
import torch

import scattennet_b200 as S
from oracle import ctc_oracle as C
from scattennet_b200 import functional as F_

pytestmark = pytest.mark.gpu
DEV = "cuda"


def peaky_logits(B, T, V, seed, sharp=6.0):
    """Random logits with a dominant class per frame (blank most of the time) - what a trained CTC head emits."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, T, V)).astype(np.float32)
    hot = np.where(rng.random((B, T)) < 0.6, 0, rng.integers(1, V, size=(B, T)))
    for b in range(B):
        for t in range(T):
            x[b, t, hot[b, t]] += sharp * rng.random()
    return x


@pytest.mark.parametrize("B,T,V,beam,seed", [(3, 6, 5, 5, 0), (4, 12, 8, 5, 1), (2, 50, 1120, 5, 2), (8, 50, 1120, 5, 3),
                                             (2, 200, 1120, 5, 4), (3, 20, 97, 1, 5), (3, 20, 97, 16, 6), (2, 9, 2, 5, 7)])
def test_beam_decode_equals_oracle(B, T, V, beam, seed):
    x = peaky_logits(B, T, V, seed)
    rng = np.random.default_rng(100 + seed)
    lens = [int(v) for v in rng.integers(0, T + 1, size=B)]
    lens[0] = T
    want = C.ctc_decode(x, beam, lens)
    got = S.ctc_decode(torch.from_numpy(x).to(DEV), beam, torch.tensor(lens))
    assert got == want
    ids, n_ids, score = F_.ctc_beam_decode(torch.from_numpy(x).to(DEV), torch.tensor(lens), beam)
    ids, n_ids, score = ids.cpu(), n_ids.cpu(), score.cpu()
    for b in range(B):
        assert ids[b, int(n_ids[b]):].eq(-1).all()
        tf_logits = np.concatenate([x[b, : lens[b], 1:], x[b, : lens[b], 0:1]], -1)
        _, lp = C.beam_search(tf_logits, beam) if lens[b] else ([], 0.0)
        assert abs(float(score[b]) - lp) <= 1e-3 * max(1.0, abs(lp))


def test_flat_logits_and_exhaustive_beam():
    """Nearly uniform frames (many near-ties) on a problem small enough for the exact answer."""
    rng = np.random.default_rng(9)
    x = (rng.standard_normal((6, 5, 4)) * 0.7).astype(np.float32)
    got = S.ctc_decode(torch.from_numpy(x).to(DEV), 16, torch.full((6,), 5))
    for b in range(6):
        tf_logits = np.concatenate([x[b, :, 1:], x[b, :, 0:1]], -1)
        lab16, lp16 = C.beam_search(tf_logits, 16)
        assert got[b] == [k for i, k in enumerate(l + 1 for l in lab16) if i == 0 or k != lab16[i - 1] + 1]


def test_decode_rejects_bad_arguments():
    x = torch.zeros(1, 4, 8, device=DEV)
    with pytest.raises(Exception):
        F_.ctc_beam_decode(x, None, 17)
    with pytest.raises(RuntimeError):
        F_.ctc_beam_decode(x.cpu(), None, 5)


def test_encoder_logits_decode():
    """End of the chain: encoder logits -> device decode == oracle decode of the same logits."""
    from scattennet_b200 import synth
    from scattennet_b200.config import model_config

    cfg = model_config("phoenix-2014t")
    m = S.MSCAEncoder(cfg, 64, precision="fp16x3").eval()
    synth.load_synth_(m, seed=5)
    m = m.to(DEV)
    kp, mask = synth.synth_batch(3, 48, seed=3, lengths=[48, 30, 9])
    with torch.no_grad():
        logits = m(kp.to(DEV), mask.to(DEV))["fuse_coord_gloss_logits"] * 8.0  # random-init heads are flat: sharpen
    lens = torch.tensor([12, 7, 2])  # valid_len_in of the pooled sequence (T/4)
    assert S.ctc_decode(logits, 5, lens) == C.ctc_decode(logits.cpu().numpy(), 5, lens.tolist())


def test_forward_host_decode():
    """Host batch in, gloss ids out: the decoded ids equal decoding the logits the same call returns."""
    from scattennet_b200 import synth
    from scattennet_b200.config import model_config

    cfg = model_config("phoenix-2014t")
    m = S.MSCAEncoder(cfg, 64, precision="fp16x3", use_graph=True).eval()
    synth.load_synth_(m, seed=5)
    m = m.to(DEV)
    kp, mask = synth.synth_batch(3, 48, seed=3, lengths=[48, 30, 9])
    lens = torch.tensor([12, 7, 2])
    with torch.no_grad():
        logits = m.forward_host(kp, mask)["fuse_coord_gloss_logits"]
        torch.cuda.synchronize()
        logits = logits.clone()
        res = m.forward_host(kp, mask, decode_beam=5, input_lengths=lens)
        torch.cuda.synchronize()
    want = C.ctc_decode(logits.numpy(), 5, lens.tolist())
    got = [res["gloss_ids"][b, : int(res["gloss_len"][b])].tolist() for b in range(3)]
    assert got == want


@pytest.mark.parametrize("kind", ["peaky", "mild", "normal", "flat"])
@pytest.mark.parametrize("beam", [1, 2, 3, 5, 8, 16])
def test_beam_decode_equals_oracle_every_sharpness(kind, beam):
    """48 sequences per case, with the steps where the sequential original deactivates a popped prefix (and so never grows
    it): the kernel reproduces them (grown-branch pass), the top path and its score equal the oracle's.  On 'mild' /
    'normal' logits a plain best-W-of-everything selection gets 1-2 % of the top paths wrong."""
    rng = np.random.default_rng(77 * beam + len(kind))
    B, T, V = 48, 20, 40
    x = rng.standard_normal((B, T, V)).astype(np.float32)
    if kind in ("peaky", "mild"):
        hot = np.where(rng.random((B, T)) < 0.6, 0, rng.integers(1, V, size=(B, T)))  # blank = class 0 here
        boost = (6.0 + 8.0 * rng.random((B, T))) if kind == "peaky" else 6.0 * rng.random((B, T))
        np.put_along_axis(x, hot[..., None], np.take_along_axis(x, hot[..., None], -1) + boost[..., None].astype(np.float32), -1)
    elif kind == "flat":
        x *= np.float32(0.05)
    lens = [T] * B
    want = C.ctc_decode(x, beam, lens)
    got = S.ctc_decode(torch.from_numpy(x).to(DEV), beam, torch.tensor(lens))
    assert got == want
    _, _, score = F_.ctc_beam_decode(torch.from_numpy(x).to(DEV), torch.tensor(lens), beam)
    for b in range(0, B, 6):
        tf_logits = np.concatenate([x[b, :, 1:], x[b, :, 0:1]], -1)
        _, lp = C.beam_search(tf_logits, beam)
        assert abs(float(score[b]) - lp) <= 1e-3 * max(1.0, abs(lp))
