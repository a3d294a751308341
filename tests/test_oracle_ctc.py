"""Known-answer checks of the CTC beam-search restatement (oracle/ctc_oracle.py).  TensorFlow - where the
reference's decode lives (utils.py:173-178) - is absent here, so the pin is the mathematics: with a beam that
holds every prefix the search is exact and must reproduce a brute-force enumeration of all alignments."""

import itertools
import math

import numpy as np
import pytest

from oracle import ctc_oracle as C


@pytest.mark.parametrize("t_len,v,seed", [(1, 3, 0), (2, 3, 1), (3, 3, 2), (4, 3, 3), (5, 3, 4), (4, 4, 5), (5, 4, 6), (6, 3, 7)])
def test_exhaustive_beam_equals_brute_force(t_len, v, seed):
    rng = np.random.default_rng(seed)
    logits = (rng.standard_normal((t_len, v)) * 2.0).astype(np.float32)
    want_lab, want_lp = C.brute_force_best(logits)
    got_lab, got_lp = C.beam_search(logits, beam_width=4096)
    assert got_lab == want_lab
    assert abs(got_lp - want_lp) <= 1e-5


def test_hand_computed_two_steps():
    # V = 2 (label 0, blank 1), T = 2, uniform: P("") = 1/4 (bb), P("0") = 3/4 (0b, b0, 00)
    logits = np.zeros((2, 2), dtype=np.float32)
    lab, lp = C.beam_search(logits, beam_width=5)
    assert lab == [0] and abs(lp - math.log(0.75)) <= 1e-6
    # strongly peaked on "label, blank, label": the repeated label survives as two tokens in the raw path
    logits = np.array([[9.0, 0.0], [0.0, 9.0], [9.0, 0.0]], dtype=np.float32)
    lab, _ = C.beam_search(logits, beam_width=5)
    assert lab == [0, 0]


def test_narrow_beam_is_a_lower_bound_and_usually_exact():
    rng = np.random.default_rng(11)
    exact = 0
    for _ in range(20):
        logits = (rng.standard_normal((5, 4)) * 3.0).astype(np.float32)
        want_lab, want_lp = C.brute_force_best(logits)
        lab, lp = C.beam_search(logits, beam_width=5)
        assert lp <= want_lp + 1e-5
        exact += lab == want_lab
    assert exact >= 18


def test_ctc_decode_wrapper_matches_reference_post_processing():
    """utils.py:166-188: blank 0 rotated to last, +1 afterwards, consecutive duplicates collapsed, per-sequence
    lengths respected."""
    rng = np.random.default_rng(5)
    x = (rng.standard_normal((3, 6, 5)) * 3.0).astype(np.float32)
    lens = [6, 4, 0]
    out = C.ctc_decode(x, 5, lens)
    assert out[2] == []
    for b in range(2):
        tf_logits = np.concatenate([x[b, : lens[b], 1:], x[b, : lens[b], 0:1]], -1)
        lab, _ = C.beam_search(tf_logits, 5)
        assert out[b] == [k for k, _ in itertools.groupby([l + 1 for l in lab])]
        assert all(1 <= g < 5 for g in out[b])
    # a sequence that repeats one gloss across a blank comes back as a single gloss (the reference's groupby quirk)
    peaked = np.full((1, 3, 3), -9.0, dtype=np.float32)
    peaked[0, 0, 2] = peaked[0, 1, 0] = peaked[0, 2, 2] = 9.0
    assert C.ctc_decode(peaked, 5, [3]) == [[2]]


def test_tensorflow_op_test_vector_labels():
    """Input and expected decodes of TensorFlow's ``ctc_decoder_ops_test.py::testCTCDecoderBeamSearch`` (depth 6, blank = 5,
    sequence length 5, beam_width = 2, top_paths = 2, merge_repeated = False): beam 0 decodes to [1, 0], beam 1 to
    [0, 1, 0].  The vector is written down FROM MEMORY of the public test (no network, no TensorFlow here), and only its label
    sequences are asserted - weak evidence next to a TensorFlow-produced fixture, which is why DESIGN.md keeps the decode
    'parity unpinned'."""
    p = np.asarray([[0.30999, 0.309938, 0.0679938, 0.0673362, 0.0708352, 0.173908],
                    [0.215136, 0.439699, 0.0370931, 0.0393967, 0.0381581, 0.230517],
                    [0.199959, 0.489485, 0.0233221, 0.0251417, 0.0233289, 0.238763],
                    [0.279611, 0.452966, 0.0204795, 0.0209126, 0.0194803, 0.20655],
                    [0.51286, 0.288951, 0.0243026, 0.0220788, 0.0219297, 0.129878]], dtype=np.float32)
    logits = np.log(p) + 2.0  # "arbitrary offset - this is fine": the search normalises every frame
    (lab0, lp0), (lab1, lp1) = C.beam_search_top_paths(logits, beam_width=2, top_paths=2)
    assert lab0 == [1, 0] and lab1 == [0, 1, 0]
    assert lp0 > lp1


def _logits(kind, rng, t_len, v):
    x = rng.standard_normal((t_len, v)).astype(np.float32)
    if kind in ("peaky", "mild"):
        hot = np.where(rng.random(t_len) < 0.6, v - 1, rng.integers(0, v - 1, size=t_len))  # blank = last class here
        x[np.arange(t_len), hot] += ((6.0 + 8.0 * rng.random(t_len)) if kind == "peaky" else 6.0 * rng.random(t_len)).astype(np.float32)
    elif kind == "flat":
        x *= np.float32(0.05)
    return x


@pytest.mark.parametrize("kind", ["peaky", "mild", "normal", "flat"])
@pytest.mark.parametrize("beam", [1, 2, 3, 5, 8])
def test_set_form_equals_sequential_search(kind, beam):
    """The heap-free formulation the device kernel computes (best W of live prefixes + extensions of the branches the
    sequential original actually grows, from each frame's 2W best labels) returns the same W paths and scores as the
    sequential restatement - including the steps where the original deactivates a popped prefix and so never
    generates its extensions (about every tenth step)."""
    rng = np.random.default_rng(1000 * beam + len(kind))
    stats = {}
    for _ in range(10):
        x = _logits(kind, rng, 20, 40)
        want = C.beam_search_top_paths(x, beam, beam)
        for lpf in (None, 2 * beam):
            got = C.beam_search_set_form(x, beam, labels_per_frame=lpf, stats=stats)
            assert [p[0] for p in got] == [p[0] for p in want], (kind, beam, lpf)
            assert all(abs(g[1] - w[1]) <= 1e-4 * max(1.0, abs(w[1])) for g, w in zip(got, want))
    if kind == "peaky" and beam >= 2:
        assert stats["deactivated"] > 0  # the order-dependent case does occur in these inputs


def test_fewer_than_2w_labels_are_not_enough():
    """Beam 1 with one label per frame: the frame's best label repeats the prefix's last label right after it was appended
    (its extension is scored from the blank-ending mass, -inf), and the second-best label is the one that extends."""
    rng = np.random.default_rng(106)
    differs = 0
    for _ in range(6):
        x = rng.standard_normal((14, 40)).astype(np.float32)
        differs += C.beam_search_set_form(x, 1, labels_per_frame=1) != C.beam_search_set_form(x, 1, labels_per_frame=2)
        assert [p[0] for p in C.beam_search_set_form(x, 1, labels_per_frame=2)] == [p[0] for p in C.beam_search_top_paths(x, 1, 1)]
    assert differs > 0
