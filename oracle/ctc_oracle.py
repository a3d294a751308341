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
  * ``beam_search_set_form`` - the heap-free restatement the device kernel implements - returns the same ``W`` paths
    and scores as the sequential restatement on random logits of every sharpness;
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


# ----------------------------------------------------------------------------- set formulation (what csrc/ctc.cu computes)


def _f32_lse(a, b):
    a, b = np.float32(a), np.float32(b)
    if a == LOG_ZERO:
        return b
    if b == LOG_ZERO:
        return a
    return np.float32(max(a, b) + np.float32(np.log1p(np.float32(np.exp(np.float32(-abs(np.float32(a - b))))))))


class _Live:
    __slots__ = ("labels", "par", "lab", "plab", "total", "blank", "label")


def beam_search_set_form(logits: np.ndarray, beam_width: int, labels_per_frame: Optional[int] = None,
                         stats: Optional[dict] = None) -> List[Tuple[List[int], float]]:
    """The same search as ``beam_search_top_paths(logits, W, W)`` restated without the sequential heap, in fp32 - the
    algorithm of the device kernel (``csrc/ctc.cu``), kept here so that it can be checked against the sequential
    restatement on the CPU (``tests/test_oracle_ctc.py``).

    A step of the original is: advance the live prefixes; then, branches in beam order and labels in ascending
    order, push every extension that beats the current worst leaf (popping that leaf).  Two facts reduce it to sets:
      * the leaves at the end are the best ``W`` of (live prefixes + generated extensions), ties to the earlier insertion;
      * an extension is generated unless its branch is skipped, and a branch ``p`` is skipped when ``is_candidate(oldp)``
        fails - harmless, none of its extensions (nor those of its later children) could enter - or when ``p`` was DEACTIVATED: ``p`` had been popped before
        the label loop of its (live, earlier) parent reached ``label(p)``; the parent then finds the child inactive,
        scores it as a fresh extension, fails, and resets its probabilities including ``oldp``.
    So only the set of grown branches has to be reproduced, by counting, in branch order, how many items beat a total.
    Counts over a whole branch need only the frame's ``2 W`` best labels (``labels_per_frame``; ``None`` = all): a prefix
    places at most ``W`` extensions, at most ``W - 1`` labels are live children of it, one label (its own last label) is
    scored from the blank-ending mass and may sink - and a count that covers every listed label is >= ``W`` anyway.  The
    one count that depends on label ORDER (extensions ``(m, k)`` with ``k < label(p)``) scans all labels."""
    t_len, v = logits.shape
    n_lab = v - 1
    root = _Live()
    root.labels, root.par, root.lab, root.plab = [], None, -1, -1
    root.total, root.blank, root.label = np.float32(0), np.float32(0), np.float32(LOG_ZERO)
    live = [root]
    w = beam_width
    for t in range(t_len):
        row = logits[t].astype(np.float32)
        mx = np.float32(row.max())
        norm = np.float32(mx + np.log(np.exp(row - mx, dtype=np.float32).sum(dtype=np.float32)))
        logp = (row - norm).astype(np.float32)
        n = len(live)
        ot, ob = [p.total for p in live], [p.blank for p in live]
        fj = [-1] * n
        for i, p in enumerate(live):
            if p.par is not None:
                for j, q in enumerate(live):
                    if q is p.par:
                        fj[i] = j
                nl = p.label
                if fj[i] >= 0:
                    nl = _f32_lse(nl, ob[fj[i]] if p.lab == p.plab else ot[fj[i]])
                p.label = np.float32(nl + logp[p.lab])
            p.blank = np.float32(ot[i] + logp[n_lab])
            p.total = _f32_lse(p.blank, p.label)
        a = [p.total for p in live]
        forb = [set() for _ in range(n)]
        for j in range(n):
            if fj[j] >= 0:
                forb[fj[j]].add(live[j].lab)
        order = sorted(range(n_lab), key=lambda k: (-float(logp[k]), k))
        top = order if labels_per_frame is None else order[:labels_per_frame]

        def score(i, k):
            return np.float32(logp[k] + (ob[i] if k == live[i].lab else ot[i]))

        def count_listed(i, tau, ge):
            c = 0
            for k in top:
                if k not in forb[i]:
                    s = score(i, k)
                    c += s != LOG_ZERO and ((s >= tau) if ge else (s > tau))
            return c

        grown, dead = [False] * n, [False] * n
        for m in range(n):
            # (the original also skips a branch whose old total does not beat the current worst leaf; none of its extensions
            # could enter, and neither could those of its later children, so that skip needs no reproduction)
            grown[m] = (not dead[m]) and ot[m] != LOG_ZERO
            if not grown[m]:
                continue
            for j in range(m + 1, n):
                if fj[j] != m:
                    continue
                base = sum(1 for l in range(n) if l != j and (a[l] > a[j] or (a[l] == a[j] and l > j)))
                base += sum(count_listed(i, a[j], False) for i in range(m) if grown[i])
                if base >= w:
                    dead[j] = True
                elif base + count_listed(m, a[j], False) >= w:
                    if stats is not None:
                        stats["scans"] = stats.get("scans", 0) + 1
                    part = sum(1 for k in range(live[j].lab) if k not in forb[m] and score(m, k) != LOG_ZERO and score(m, k) > a[j])
                    dead[j] = base + part >= w
        if stats is not None:
            stats["steps"] = stats.get("steps", 0) + 1
            stats["deactivated"] = stats.get("deactivated", 0) + sum(dead)
        items = [(-float(a[i]), 0, i, -1) for i in range(n)]
        for i in range(n):
            if grown[i]:
                for k in top:
                    if k not in forb[i] and score(i, k) != LOG_ZERO:
                        items.append((-float(score(i, k)), 1, i, k))
        items.sort()
        nxt = []
        for neg, kind, i, k in items[:w]:
            if kind == 0:
                nxt.append(live[i])
            else:
                q = _Live()
                q.labels, q.par, q.lab, q.plab = live[i].labels + [k], live[i], k, live[i].lab
                q.total, q.blank, q.label = np.float32(-neg), np.float32(LOG_ZERO), np.float32(-neg)
                nxt.append(q)
        live = nxt
    return [(p.labels, float(p.total)) for p in live]


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
