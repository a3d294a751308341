"""TEST INFRASTRUCTURE - not part of the product path.

CPU restatement of the CTC decode the reference runs on the encoder's logits
(``utils.py:164-189``: ``ctc_decode(gloss_logits, beam_size, input_lengths)``):

    1. ``[B,T,V] -> [T,B,V]``, class 0 (the CTC blank of ``nn.CTCLoss(blank=0)``) rotated to the LAST
       class, which is where TensorFlow expects the blank (``utils.py:166-172``);
    2. ``tf.nn.ctc_beam_search_decoder(inputs, sequence_length, beam_width=beam_size, top_paths=1)``
       (``utils.py:173-178``);
    3. ``+ 1`` to undo the rotation and ``itertools.groupby`` to collapse consecutive duplicates
       (``utils.py:180-188``).

Step 2 lives in a third-party dependency that is ABSENT from ``/root/reference`` and from this
image: TensorFlow (no version is pinned anywhere in the reference - it has no requirements file;
``README.md:25-30``).  ``beam_search`` below restates the published algorithm of
``tensorflow/core/util/ctc/ctc_beam_search.h`` (``CTCBeamSearchDecoder<>::Step`` / ``TopPaths``, the
kernel behind ``tf.nn.ctc_beam_search_decoder``; v2 API => ``merge_repeated=False``, no label
selection, default ``BaseBeamScorer`` whose expansion / end scores are 0).

PARITY UNPINNED against TensorFlow itself (it cannot be run here, and the reference holds no golden
vectors for this path).  What pins this file instead (``tests/test_oracle_ctc.py``):
  * with a beam wide enough to hold every prefix, prefix beam search is exact: its best labelling and
    score must equal a brute-force enumeration of all alignments (known-answer check of the recursion);
  * hand-computed 2-step cases; invariance of the result under the reference's rotation + ``+1``;
  * the input of TensorFlow's own op test (``ctc_decoder_ops_test.py::testCTCDecoderBeamSearch``: 5 x 6 probability
    matrix, beam_width 2, top_paths 2) with the two decoded label sequences that test expects - REPRODUCED FROM
    MEMORY (the file cannot be fetched offline), labels only, so it is supporting evidence, not a pin.
"""

from __future__ import annotations

import itertools
import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

LOG_ZERO = -math.inf


def log_sum_exp(a: float, b: float) -> float:
    """ctc_loss_util.h LogSumExp: log(exp(a) + exp(b)) with the log-zero conventions."""
    if a == LOG_ZERO:
        return b
    if b == LOG_ZERO:
        return a
    return (a if a > b else b) + math.log1p(math.exp(-abs(a - b)))


class _Entry:
    """BeamEntry of ctc_beam_entry.h: a prefix in the trie with (blank, label, total) log-probabilities
    for the previous (oldp) and current (newp) time step."""

    __slots__ = ("parent", "label", "children", "oldp", "newp")

    def __init__(self, parent: Optional["_Entry"], label: int):
        self.parent, self.label, self.children = parent, label, {}
        self.oldp = [LOG_ZERO, LOG_ZERO, LOG_ZERO]  # total, blank, label
        self.newp = [LOG_ZERO, LOG_ZERO, LOG_ZERO]

    def active(self) -> bool:
        return self.newp[0] != LOG_ZERO

    def child(self, label: int) -> "_Entry":
        c = self.children.get(label)
        if c is None:
            c = _Entry(self, label)
            self.children[label] = c
        return c

    def labels(self) -> List[int]:
        out, e = [], self
        while e.parent is not None:
            out.append(e.label)
            e = e.parent
        return out[::-1]


def beam_search(logits: np.ndarray, beam_width: int) -> Tuple[List[int], float]:
    """``logits [T, V]`` (raw, un-normalised; blank = class V-1, TensorFlow's convention) ->
    ``(best labelling, its log-probability)``.  One call = one batch element over its own length."""
    return beam_search_top_paths(logits, beam_width, 1)[0]


def beam_search_top_paths(logits: np.ndarray, beam_width: int, top_paths: int) -> List[Tuple[List[int], float]]:
    """Same search, ``TopPaths(n)``: the ``n`` best leaves in descending total probability."""
    t_len, v = logits.shape
    blank = v - 1
    root = _Entry(None, -1)
    root.newp = [0.0, 0.0, LOG_ZERO]  # total = blank = log 1
    leaves: List[_Entry] = [root]
    for t in range(t_len):
        row = logits[t].astype(np.float32)
        mx = np.float32(row.max())
        norm = float(mx + np.log(np.exp(row - mx, dtype=np.float32).sum(dtype=np.float32)))
        logp = (row - np.float32(norm)).astype(np.float32)
        branches = sorted(leaves, key=lambda e: -e.newp[0])  # stable: descending newp.total
        leaves = []
        for b in branches:
            b.oldp = list(b.newp)
        # probabilities of the existing beams at t
        for b in branches:
            if b.parent is not None:
                if b.parent.active():
                    prev = b.parent.oldp[1] if b.label == b.parent.label else b.parent.oldp[0]
                    b.newp[2] = log_sum_exp(b.newp[2], prev)
                b.newp[2] += float(logp[b.label])
            b.newp[1] = b.oldp[0] + float(logp[blank])
            b.newp[0] = log_sum_exp(b.newp[1], b.newp[2])
            leaves.append(b)  # always fits: len(branches) <= beam_width

        def bottom() -> _Entry:
            return min(leaves, key=lambda e: e.newp[0])

        def is_candidate(p) -> bool:
            return p[0] > LOG_ZERO and (len(leaves) < beam_width or p[0] > bottom().newp[0])

        # grow new leaves
        for b in branches:
            if not is_candidate(b.oldp):
                continue
            for c_label in range(blank):
                c = b.child(c_label)
                if c.active():
                    continue
                prev = b.oldp[1] if c_label == b.label else b.oldp[0]
                c.newp = [LOG_ZERO, LOG_ZERO, float(logp[c_label]) + prev]
                c.newp[0] = c.newp[2]
                if is_candidate(c.newp):
                    if len(leaves) == beam_width:
                        bot = bottom()
                        bot.newp = [LOG_ZERO, LOG_ZERO, LOG_ZERO]
                        leaves.remove(bot)
                    leaves.append(c)
                else:
                    c.oldp = [LOG_ZERO, LOG_ZERO, LOG_ZERO]
                    c.newp = [LOG_ZERO, LOG_ZERO, LOG_ZERO]
    ranked = sorted(leaves, key=lambda e: -e.newp[0])  # stable, like the single-best max() over the same order
    return [(e.labels(), e.newp[0]) for e in ranked[:top_paths]]


def ctc_decode(gloss_logits: np.ndarray, beam_size: int, input_lengths: Sequence[int]) -> List[List[int]]:
    """utils.py:164-189 - ``gloss_logits [B,T,V]`` with blank = class 0; returns one gloss-id list per
    sequence (ids in the original numbering, consecutive duplicates collapsed)."""
    out = []
    for b in range(gloss_logits.shape[0]):
        x = np.asarray(gloss_logits[b], dtype=np.float32)[: int(input_lengths[b])]
        tf_logits = np.concatenate([x[:, 1:], x[:, 0:1]], axis=-1)  # blank 0 -> last
        labels, _ = beam_search(tf_logits, beam_size)
        ids = [l + 1 for l in labels]
        out.append([k for k, _ in itertools.groupby(ids)])
    return out


# ----------------------------------------------------------------------------- brute force (known answers)


def brute_force_best(logits: np.ndarray) -> Tuple[List[int], float]:
    """Exact most probable labelling of a tiny problem by enumerating all V^T alignments
    (blank = class V-1): collapse repeats, drop blanks, sum path probabilities per labelling."""
    t_len, v = logits.shape
    blank = v - 1
    x = logits.astype(np.float64)
    logp = x - (x.max(1, keepdims=True) + np.log(np.exp(x - x.max(1, keepdims=True)).sum(1, keepdims=True)))
    table = {}
    for path in itertools.product(range(v), repeat=t_len):
        lab = tuple(k for k, _ in itertools.groupby(path) if k != blank)
        lp = float(sum(logp[t, c] for t, c in enumerate(path)))
        table[lab] = np.logaddexp(table.get(lab, -np.inf), lp)
    best = max(table, key=lambda k: table[k])
    return list(best), float(table[best])
