"""The encoder path of ``MSCA_Net`` as one module (reference
``model/__init__.py:72-159``): region split, the three ``KeypointModule``
streams (``body_encoder``, ``left_encoder``, ``right_encoder``),
``coordinates_fusion`` and the four linear classifiers of ``recognition_head``
with their +-50 clamp.  Sub-module names equal the reference's, so
``MSCA_Net.state_dict()`` loads with ``strict=False``.  ``alignment=True`` adds the
first consumer of the path, the BiLSTM ``fuse_alignment_head`` (``alignment_gloss_logits``,
reference ``model/__init__.py:27-29,51``); the losses and the
tokenizer stay outside.

``forward(keypoints[B,T,K,2], mask[B,T])`` runs ~80 grouped kernel launches;
``use_graph=True`` replays them from a CUDA graph captured per ``(B, T)``.
"""

from __future__ import annotations

import os
from typing import Dict, Optional

import torch
from torch import nn

from . import _lib as L
from . import functional as F_
from .alignment_module import AlignmentModule, alignment_forward
from .fusion import CoordinatesFusion, coordinates_fusion_forward
from .keypoint_module import KeypointModule, streams_forward

PARTS = ("body", "left", "right")  # order of model/__init__.py:133-142
HOST_GRAPH = os.environ.get("SCATT_HOST_GRAPH", "1") != "0"  # forward_host as one graph (copies included)


class LinearHeads(nn.Module):
    """The four ``nn.Linear`` classifiers of ``RecognitionHead`` (reference
    ``model/__init__.py:14-25``), same attribute names."""

    def __init__(self, cfg, vocab_size: int, alignment: bool = False):
        super().__init__()
        c = cfg["residual_blocks"][-1]
        self.left_gloss_classifier = nn.Linear(c, vocab_size)
        self.right_gloss_classifier = nn.Linear(c, vocab_size)
        self.body_gloss_classifier = nn.Linear(c, vocab_size)
        self.fuse_coord_classifier = nn.Linear(cfg["out_fusion_dim"], vocab_size)
        if alignment:
            self.fuse_alignment_head = AlignmentModule(**cfg["alignment_module"], cls_num=vocab_size)


BIND_INPUTS = os.environ.get("SCATT_BIND_INPUTS", "1") != "0"  # False: always copy the inputs into the captured graph's static buffers


class MSCAEncoder(nn.Module):
    def __init__(self, cfg, vocab_size: int, precision: Optional[str] = None, use_graph: bool = False, micro_batches: int = 1,
                 alignment: bool = False):
        super().__init__()
        self.cfg = dict(cfg)
        self.body_encoder = KeypointModule(cfg["body_idx"], num_frame=cfg["num_frame"], cfg=cfg)
        self.left_encoder = KeypointModule(cfg["left_idx"], num_frame=cfg["num_frame"], cfg=cfg)
        self.right_encoder = KeypointModule(cfg["right_idx"], num_frame=cfg["num_frame"], cfg=cfg)
        self.coordinates_fusion = CoordinatesFusion(cfg["in_fusion_dim"], cfg["out_fusion_dim"], 0.2)
        self.recognition_head = LinearHeads(cfg, vocab_size, alignment)
        self.alignment = alignment
        self.precision = precision
        self.use_graph = use_graph
        # > 1: inside the captured graph the batch is cut into this many independent sub-batches that
        # run as parallel graph branches (sequences are independent); their latency-bound kernels overlap
        self.micro_batches = micro_batches
        self._graphs: Dict = {}        # (shape, device, precision, heads, compact) -> captured graph; LRU, see _cache_put
        self._host_staging: Dict = {}  # (B, T, device) -> pinned staging slots + their graphs (forward_host)
        self._idx_cache: Dict = {}
        self.max_cached_shapes = 8     # captured graphs / staging sets kept per cache (variable-length inference)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_graphs())

    # ------------------------------------------------------------------ plumbing
    def invalidate_graphs(self) -> None:
        """Drop every captured graph and staging set: they hold pointers to the packed (split-plane / transposed)
        copies of the weights, which are rebuilt when a parameter changes."""
        self._graphs.clear()
        self._host_staging.clear()

    def _apply(self, fn, recurse=True):  # .to() / .cuda() / .float(): parameters move, captured pointers go stale
        self.invalidate_graphs()
        return super()._apply(fn, recurse)

    def _cache_put(self, cache: Dict, key, value):
        cache[key] = value
        while len(cache) > self.max_cached_shapes:  # dicts keep insertion order: evict the least recently used
            cache.pop(next(iter(cache)))

    @staticmethod
    def _cache_get(cache: Dict, key):
        ent = cache.pop(key, None)
        if ent is not None:
            cache[key] = ent  # most recently used goes last
        return ent

    def _joint_idx(self, device):
        key = str(device)
        if key not in self._idx_cache:
            self._idx_cache[key] = [torch.tensor(self.cfg[p + "_idx"], dtype=torch.int32, device=device) for p in PARTS]
        return self._idx_cache[key]

    def _n_used(self) -> int:
        return len(sorted({j for p in PARTS for j in self.cfg[p + "_idx"]}))

    def _compact_idx(self, device):
        """``(used joints, per-stream indices into the used-joint list)`` for the host-gather path."""
        key = "compact:" + str(device)
        if key not in self._idx_cache:
            used = sorted({j for p in PARTS for j in self.cfg[p + "_idx"]})
            pos = {j: i for i, j in enumerate(used)}
            remap = [torch.tensor([pos[j] for j in self.cfg[p + "_idx"]], dtype=torch.int32, device=device) for p in PARTS]
            self._idx_cache[key] = (torch.tensor(used, dtype=torch.int64), remap)
        return self._idx_cache[key]

    def load_reference_state_dict(self, state_dict, strict_path: bool = True):
        """Load an ``MSCA_Net`` state dict: every key of the encoder path must be
        present and match; keys outside the path (``recognition_head.fuse_alignment_head.*``)
        are ignored."""
        own = self.state_dict()
        picked = {k: v for k, v in state_dict.items() if k in own}
        missing = [k for k in own if k not in picked]
        if strict_path and missing:
            raise KeyError(f"state dict lacks encoder-path keys: {missing[:5]}{'...' if len(missing) > 5 else ''}")
        bad = [k for k in picked if tuple(picked[k].shape) != tuple(own[k].shape)]
        if bad:
            raise ValueError(f"shape mismatch for {bad[:5]}")
        return self.load_state_dict(picked, strict=strict_path)

    # ------------------------------------------------------------------ forward
    def _validate(self, keypoints: torch.Tensor, mask: torch.Tensor, compact: bool = False):
        """Shape / index checks the reference gets from ATen (``keypoints[:, :, idx, :]`` raises ``IndexError`` for a
        joint index past ``K``, ``model/__init__.py:133-142``); the kernels would read out of bounds instead."""
        if keypoints.ndim != 4 or keypoints.shape[-1] != 2:
            raise ValueError(f"keypoints must be [B, T, K, 2]; got {tuple(keypoints.shape)}")
        b, t, k = keypoints.shape[:3]
        if tuple(mask.shape) != (b, t):
            raise ValueError(f"mask must be [B, T] = {(b, t)}; got {tuple(mask.shape)}")
        need = self._n_used() if compact else self._max_joint() + 1
        if k < need:
            raise IndexError(f"index {need - 1} is out of bounds for dimension 2 with size {k}")

    def _max_joint(self) -> int:
        if "max_joint" not in self._idx_cache:
            self._idx_cache["max_joint"] = max(j for p in PARTS for j in self.cfg[p + "_idx"])
        return self._idx_cache["max_joint"]

    def _prefetch_begin(self, prec, b: int, t: int):
        """The weight prefetch branch (``scatt_l2_prefetch`` on the side stream); returns the branch to ``join()`` or None."""
        if not (F_.L2_PREFETCH and prec.uses_planes and b * t <= 8192):
            return None
        planes = F_.weight_planes_of(self, prec)
        if not planes:
            return None
        with F_.SideBranch([]) as pf:
            F_.l2_prefetch(planes)
        return pf

    def _run(self, keypoints: torch.Tensor, key_mask: torch.Tensor, with_heads: bool = True, compact: bool = False,
             prefetch: bool = True) -> Dict[str, torch.Tensor]:
        prec = F_.get_precision(self.precision)
        b, t = keypoints.shape[:2]
        mods = [self.body_encoder, self.left_encoder, self.right_encoder]
        # a compact tensor (only the joints the three streams use, see forward_host) carries remapped indices
        idx = self._compact_idx(keypoints.device)[1] if compact else self._joint_idx(keypoints.device)
        # small batches: ~60 dependent launches each pay DRAM latency on their first weight tile when the step starts with
        # a cold L2; one launch on a parallel branch hints all weight planes (~70 MB) into L2 (csrc/prefetch.cu).  Large
        # batches stream more activations through L2 than it holds - the hint would be evicted before use.
        pf = self._prefetch_begin(prec, b, t) if prefetch else None
        blocks = streams_forward(prec, mods, keypoints, idx, key_mask, b, t)
        (body, left, right), tp = blocks[-1]
        lg, heads_branch = None, None
        if with_heads:
            # the three stream classifiers only need the stream features: they run as a side branch beside
            # the fusion block
            rh = self.recognition_head
            ep = F_.make_epilogue(clamp=50.0)
            for a in (left, right, body):
                a.with_planes(prec)  # made here, not inside the branch: both consumers read them
            with F_.SideBranch([left, right, body]) as heads_branch:
                lg = F_.linear(prec, [left, right, body],
                               [F_.pack_of(rh, "left", [rh.left_gloss_classifier]), F_.pack_of(rh, "right", [rh.right_gloss_classifier]),
                                F_.pack_of(rh, "body", [rh.body_gloss_classifier])], ep, out_planes=False)
        fuse = coordinates_fusion_forward(prec, self.coordinates_fusion, left, right, body, b, tp, out_planes=with_heads)
        out = {"body_embed": body.f32.view(b, tp, -1), "left_embed": left.f32.view(b, tp, -1),
               "right_embed": right.f32.view(b, tp, -1), "fuse_embed": fuse.f32.view(b, tp, -1)}
        if with_heads:
            fl = F_.linear(prec, [fuse], [F_.pack_of(rh, "fuse", [rh.fuse_coord_classifier])], ep, out_planes=False)[0]
            heads_branch.join(lg)
            out.update(left=lg[0].f32.view(b, tp, -1), right=lg[1].f32.view(b, tp, -1), body=lg[2].f32.view(b, tp, -1),
                       fuse_coord_gloss_logits=fl.f32.view(b, tp, -1))
            if self.alignment:
                al = alignment_forward(prec, rh.fuse_alignment_head, fuse, b, tp, clamp=50.0)
                out["alignment_gloss_logits"] = al.f32.view(b, tp, -1)
        if pf is not None:
            pf.join()
        return out

    def forward(self, keypoints: torch.Tensor, mask: torch.Tensor, with_heads: bool = True,
                check_finite: bool = False, compact: bool = False) -> Dict[str, torch.Tensor]:
        """``keypoints [B,T,K,2]`` (the full collated tensor - the region split is
        part of the path), ``mask [B,T]`` 0/1.  Returns the stream / fusion
        features and the clamped per-frame logits (fp32).

        ``check_finite=True`` reproduces the NaN / inf guards of ``MSCA_Net.forward`` (reference
        ``model/__init__.py:130-167``: input, the three stream outputs, the fused features, every head) and
        raises ``ValueError`` naming the first offender - with one fused kernel and one 4-byte read-back
        instead of 16 host synchronisations.  ``compact=True``: ``keypoints`` holds only the joints the three
        streams use, in ascending joint order (what ``forward_host`` ships).

        With ``use_graph=True`` the returned tensors are the captured graph's static outputs: the next call with
        the same shape overwrites them (clone what must survive).  When the same input tensors (same storage) are
        passed on consecutive calls, a graph that reads them in place is captured and replayed from then on - refill
        them in place on the calling stream; the model keeps a reference to them until the shape's cache entry is
        evicted or ``invalidate_graphs()`` is called (``SCATT_BIND_INPUTS=0`` always copies instead).  Captured graphs are dropped by
        ``load_state_dict`` / ``.to()`` / ``.cuda()`` (``invalidate_graphs``); after modifying parameters in place
        call ``invalidate_graphs()`` yourself."""
        if self.training:
            raise RuntimeError("scattennet_b200 is inference-only: call .eval() before forward")
        F_.require_cuda(keypoints, mask)
        self._validate(keypoints, mask, compact)
        if keypoints.dtype != torch.float32 or not keypoints.is_contiguous():
            keypoints = keypoints.float().contiguous()
        if not self.use_graph:
            out = self._run(keypoints, F_.key_mask_u8(mask), with_heads, compact)
        else:
            out = self._run_graph(keypoints, mask, with_heads, compact)
        if check_finite:
            names = ["input keypoints"] + list(out)
            bits = int(F_.finite_flags([keypoints] + [out[k] for k in out]).item())
            if bits:
                raise ValueError("NaN or inf in " + names[(bits & -bits).bit_length() - 1])
        return out

    def forward_host(self, keypoints: torch.Tensor, mask: torch.Tensor, heads=("fuse_coord_gloss_logits",), device=None,
                     gather: bool = False, decode_beam: int = 0, input_lengths: Optional[torch.Tensor] = None,
                     gather_to_host: str = "rank0"):
        """End-to-end call for host-resident batches (the collator -> device path).

        ``keypoints [B,T,K,2]`` / ``mask [B,T]`` are CPU tensors.  Only the joints the three streams read
        (48 of 542 for the Phoenix configs: 384 of 4336 bytes per frame) are gathered - an exact copy - into a
        pinned staging buffer and sent to the device; the requested ``heads`` come back in pinned host
        tensors.  Returns ``{name: host tensor}``; call ``torch.cuda.current_stream().synchronize()`` (or use
        the tensors after any sync) before reading them.

        **Buffering.**  Staging and result buffers are double-buffered per ``(B, T)``: a call may be issued while
        the previous one is still running on the device (the host gather of batch ``i + 1`` overlaps the encoder
        of batch ``i``); a third call first waits, on the host, for the call two steps back to finish.  The
        tensors returned by call ``i`` are therefore valid until call ``i + 2`` of the same shape is issued.

        ``gather=True`` (inside an initialised ``torch.distributed`` job): the first head is all-gathered over
        NVLink on the device.  ``gather_to_host="rank0"``: rank 0 reads the gathered ``[world*B, T', V]`` logits
        back, the other ranks their own shard (one decoder process); ``"shard"``: every rank reads only its own
        shard (one decoder per rank; the gathered tensor stays on the device under ``"<head>/gathered_dev"``).
        ``decode_beam > 0``: the first head is CTC-decoded on the device (prefix beam search, the reference's
        ``utils.ctc_decode``; ``input_lengths [B]`` = valid pooled frames per sequence) and only
        ``gloss_ids [B,T'] int32`` (padded with -1) and ``gloss_len [B]`` come back instead of logits."""
        if self.training:
            raise RuntimeError("scattennet_b200 is inference-only: call .eval() before forward")
        dev = device or next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("scattennet_b200 runs on a CUDA device (sm_100a) only; there is no CPU fallback")
        if keypoints.is_cuda or mask.is_cuda:
            raise RuntimeError("forward_host takes host tensors; call forward() for device-resident batches")
        self._validate(keypoints, mask)
        used, _ = self._compact_idx(dev)
        b, t = keypoints.shape[:2]
        key = ("host", b, t, str(dev))
        st = self._cache_get(self._host_staging, key)
        if st is None:
            # keypoints and mask share ONE staging buffer per slot (and one on the device): a single H2D copy per step
            n_kp = b * t * used.numel() * 2 * 4

            def views(buf):
                return buf[:n_kp].view(torch.float32).view(b, t, used.numel(), 2), buf[n_kp:].view(b, t)

            def make_slot():
                pin = torch.empty(n_kp + b * t, dtype=torch.uint8).pin_memory()
                kp_pin, mask_pin = views(pin)
                return {"pin": pin, "kp_pin": kp_pin, "mask_pin": mask_pin, "out_pin": {}, "busy": None}
            in_dev = torch.empty(n_kp + b * t, dtype=torch.uint8, device=dev)
            kp_dev, mask_dev = views(in_dev)
            st = {"slots": [make_slot(), make_slot()], "next": 0, "in_dev": in_dev, "kp_dev": kp_dev, "mask_dev": mask_dev}
            self._cache_put(self._host_staging, key, st)
        slot = st["slots"][st["next"]]
        st["next"] ^= 1
        if slot["busy"] is not None:
            slot["busy"].synchronize()  # the step that last used this slot has read its staging and written its results
        torch.index_select(keypoints, 2, used, out=slot["kp_pin"])  # exact gather on the host
        slot["mask_pin"].copy_(mask != 0)
        try:
            if self.use_graph and decode_beam <= 0 and HOST_GRAPH and (not gather or gather_to_host == "shard"):
                # the whole step - both H2D copies, the encoder, the D2H copies of the requested heads - is ONE
                # captured graph: a single launch instead of eight stream operations with the host in between.
                # Sharded jobs add one launch behind it: the NVLink push of this rank's logits (the read-back of the own
                # shard is already inside the graph and overlaps it).
                res, out_dev = self._host_graph_step(st, slot, heads, dev, d2h_in_graph=not gather)
                if gather:
                    # the read-back of the own shard (PCIe) and the push of the logits to the peers (NVLink) both start
                    # when the captured encoder has finished: the copy runs on a side stream beside the push kernel
                    from .distributed import gather_logits_peer

                    main = torch.cuda.current_stream(dev)
                    side = F_.side_stream(dev)
                    side.wait_stream(main)
                    with torch.cuda.stream(side):
                        for k in heads:
                            res[k].copy_(out_dev[k], non_blocking=True)
                    res[heads[0] + "/gathered_dev"] = gather_logits_peer(out_dev[heads[0]])
                    main.wait_stream(side)
                return res
            st["in_dev"].copy_(slot["pin"], non_blocking=True)
            out = self.forward(st["kp_dev"], st["mask_dev"], compact=True)
            extra = {}
            if gather:
                import torch.distributed as dist

                from .distributed import gather_logits_peer

                full = gather_logits_peer(out[heads[0]])
                if gather_to_host == "rank0":
                    if dist.get_rank() == 0:
                        out = dict(out)
                        out[heads[0]] = full
                elif gather_to_host == "shard":
                    extra[heads[0] + "/gathered_dev"] = full
                else:
                    raise ValueError("gather_to_host must be 'rank0' or 'shard'")
            if decode_beam > 0:
                lens = st.get("len_dev")
                if input_lengths is not None:
                    if lens is None:
                        lens = st["len_dev"] = torch.empty(b, dtype=torch.int32, device=dev)
                    len_pin = slot.get("len_pin")
                    if len_pin is None:
                        len_pin = slot["len_pin"] = torch.empty(b, dtype=torch.int32).pin_memory()
                    len_pin.copy_(input_lengths.to(torch.int32))
                    lens.copy_(len_pin, non_blocking=True)
                ids, n_ids, _ = F_.ctc_beam_decode(out[heads[0]], lens if input_lengths is not None else None, decode_beam)
                out = {"gloss_ids": ids, "gloss_len": n_ids}
                heads = ("gloss_ids", "gloss_len")
            res = dict(extra)
            for k in heads:
                pin = slot["out_pin"].get(k)
                if pin is None or pin.shape != out[k].shape:
                    pin = torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory()
                    slot["out_pin"][k] = pin
                pin.copy_(out[k], non_blocking=True)
                res[k] = pin
            return res
        finally:
            busy = slot["busy"] or torch.cuda.Event()
            busy.record(torch.cuda.current_stream(dev))
            slot["busy"] = busy

    def _host_graph_step(self, st, slot, heads, dev, d2h_in_graph: bool = True):
        key = ("hostgraph", tuple(heads), F_.get_precision(self.precision).name, d2h_in_graph)
        ent = slot.get(key)
        if ent is None:
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():  # warm-up: packs weights, sets kernel attributes
                st["in_dev"].copy_(slot["pin"], non_blocking=True)
                for _ in range(2):
                    out = self._run(st["kp_dev"], st["mask_dev"], True, True)
                pins = {k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory() for k in heads}
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            # the two slots' graphs replay in stream order, never concurrently: they share one memory pool
            pool = st.get("pool")
            with torch.cuda.graph(graph, pool=pool), torch.no_grad():
                # the weight prefetch does not depend on the batch: it forks before the H2D copies and runs under them
                b, t = st["kp_dev"].shape[:2]
                pf = self._prefetch_begin(F_.get_precision(self.precision), b, t) if min(self.micro_batches, b) <= 1 else None
                st["in_dev"].copy_(slot["pin"], non_blocking=True)
                if pf is not None:
                    out = self._run(st["kp_dev"], st["mask_dev"], True, True, prefetch=False)
                    pf.join()
                else:
                    out = self._run_branches(st["kp_dev"], st["mask_dev"], True, True)
                if d2h_in_graph:
                    for k in heads:
                        pins[k].copy_(out[k], non_blocking=True)
            if pool is None:
                st["pool"] = graph.pool()
            ent = slot[key] = (graph, pins, out)
        graph, pins, out = ent
        graph.replay()
        return dict(pins), out

    # ------------------------------------------------------------------ CUDA graph replay
    def _run_graph(self, keypoints, mask, with_heads, compact=False):
        """Replay the captured graph of this shape.  The first capture reads static copies of the inputs (any tensor may
        be passed from call to call).  When the SAME input tensors come back (same storage, consecutive calls - a serving
        loop that refills its input buffers in place, the benchmark's resident batch) a second graph is captured that
        reads them where they are: no 6.9 MB device-to-device copy of the keypoints and no mask conversion launches in
        front of the replay."""
        key = (tuple(keypoints.shape), str(keypoints.device), F_.get_precision(self.precision).name, with_heads, compact)
        ent = self._cache_get(self._graphs, key)
        if ent is None:
            static_kp = torch.empty_like(keypoints)
            static_mask = torch.empty(mask.shape, dtype=torch.uint8, device=mask.device)
            static_kp.copy_(keypoints)
            static_mask.copy_(F_.key_mask_u8(mask))
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up: packs weights, sets kernel attributes
                for _ in range(2):
                    self._run(static_kp, static_mask, with_heads, compact)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            n0 = L.launch_count()
            with torch.cuda.graph(graph):
                static_out = self._run_branches(static_kp, static_mask, with_heads, compact)
            ent = [graph, static_kp, static_mask, static_out, L.launch_count() - n0, None, None]  # [5] last inputs seen, [6] bound graph
            self._cache_put(self._graphs, key, ent)
        ptrs = (keypoints.data_ptr(), mask.data_ptr(), mask.dtype, tuple(mask.stride()))
        bound = ent[6]
        if bound is not None and bound[0] == ptrs:
            bound[1].replay()
            return bound[2]
        if BIND_INPUTS and ent[5] == ptrs:  # the same tensors twice in a row: capture a graph that reads them in place
            graph2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph2, pool=ent[0].pool()):  # never replayed concurrently with the static graph
                out2 = self._run_branches(keypoints, F_.key_mask_u8(mask), with_heads, compact)
            ent[6] = (ptrs, graph2, out2, keypoints, mask)  # the references keep the storage alive
            graph2.replay()
            return out2
        ent[5] = ptrs
        graph, static_kp, static_mask, static_out = ent[:4]
        static_kp.copy_(keypoints, non_blocking=True)
        static_mask.copy_(F_.key_mask_u8(mask), non_blocking=True)
        graph.replay()
        return static_out

    def _run_branches(self, kp, km, with_heads, compact=False):
        """One `_run` per sub-batch on forked streams (captured as parallel graph branches), joined and
        concatenated on the capturing stream."""
        b = kp.shape[0]
        n = min(self.micro_batches, b)
        if n <= 1:
            return self._run(kp, km, with_heads, compact)
        bounds = [(i * b) // n for i in range(n + 1)]
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        parts, joins = [], []
        for i in range(n):
            st = main if i == 0 else torch.cuda.Stream()
            if i:
                st.wait_event(fork)
            with torch.cuda.stream(st):
                parts.append(self._run(kp[bounds[i]:bounds[i + 1]], km[bounds[i]:bounds[i + 1]], with_heads, compact))
                if i:
                    ev = torch.cuda.Event()
                    ev.record(st)
                    joins.append(ev)
        for ev in joins:
            main.wait_event(ev)
        return {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}

    def graph_launches(self, keypoints_shape, device, with_heads=True) -> int:
        """Kernels inside the captured graph for this shape (0 if not captured)."""
        key = (tuple(keypoints_shape), str(device), F_.get_precision(self.precision).name, with_heads, False)
        ent = self._graphs.get(key)
        return 0 if ent is None else ent[4]
