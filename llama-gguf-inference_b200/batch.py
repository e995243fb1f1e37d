"""Batched decode: several sequences ("slots") advance by one token each in ONE pass over the weights.

What llama-server's continuous batching does for concurrent requests (the reference's benchmark drives it with
`--concurrent N`, /root/reference/scripts/benchmark.py:178-214): the decode tokens of all busy slots form one batch.
Here the batch goes through csrc/gemv_batch.cu (weights streamed once for up to 8 tokens per launch) with per-token
cache addressing for RoPE / KV write / attention; every token's arithmetic is the batch-1 path's, so a sequence
produces bit-identical logits alone or in a batch (tests/test_gpu_engine.py).

The host owns the schedule: each step it hands (slot, token, position) triples; they are staged through pinned
memory, one captured CUDA graph per batch size replays the whole forward pass, and the greedy tokens (and, for
sampled requests, logits rows) are read back.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import cabi


class BatchDecoder:
    def __init__(self, eng, max_batch: int | None = None):
        torch = eng.torch
        self.eng, self.torch, self.lib, self.hp = eng, torch, eng.lib, eng.hp
        hp, dev = eng.hp, eng.dev
        self.nb_max = NB = int(max_batch or max(len(eng.slots), 16))   # >= 16 so that prompts prefill 16 tokens per pass
        self.stream = eng.stream
        self.use_pdl = eng.use_pdl
        # tensor parallel: this rank's heads, FFN rows and vocabulary shard (parallel.py); the row-split projections leave
        # unrounded f64 partials that are all-reduced before the one rounding, the arg-max travels as one sortable key per
        # token -- the batched step is then bit-identical with the single-GPU one, like the single-sequence step
        tp = self.tp = eng.tp_size
        self.nh, self.nkv, self.ffl, self.vl = hp.n_head // tp, hp.n_kv // tp, hp.ff // tp, hp.vocab // tp
        qd, kvd = self.nh * hp.head_dim, self.nkv * hp.head_dim
        f32 = lambda *s: torch.zeros(s, dtype=torch.float32, device=dev)  # noqa: E731
        self.x, self.q, self.k, self.v, self.att = f32(NB, hp.d), f32(NB, qd), f32(NB, kvd), f32(NB, kvd), f32(NB, qd)
        self.h, self.logits = f32(NB, self.ffl), f32(NB, self.vl)
        if tp > 1:
            self.y64 = torch.zeros((NB, hp.d), dtype=torch.float64, device=dev)
            self.keys = torch.zeros(NB, dtype=torch.int64, device=dev)
        kmax = max(hp.d, self.ffl, qd)
        self.act = torch.zeros(max(NB * self.lib.ggb_act_image_bytes(kmax), self.lib.ggb_act_tiled_bytes(kmax, NB)), dtype=torch.uint8, device=dev)
        self.meta = torch.zeros((3, NB), dtype=torch.int32, device=dev)          # token ids | positions | slots
        self.meta_host = torch.zeros((3, NB), dtype=torch.int32).pin_memory()
        self.next_tok = torch.zeros(NB, dtype=torch.int32, device=dev)
        self.next_host = [torch.zeros(NB, dtype=torch.int32).pin_memory() for _ in range(2)]   # two launches may be outstanding
        self.events = [torch.cuda.Event() for _ in range(2)]
        self._seq, self._last, self._warm = 0, None, set()
        self.logits_host = None
        self.slot_stride = hp.n_layer * eng.n_ctx * kvd                         # elements between two slots' caches (this rank's heads)
        self._graphs = {}

    # ------------------------------------------------------------------ one forward pass for nb tokens
    def _enqueue(self, nb: int, s: int, head: bool = True):
        lib, hp, e, pdl = self.lib, self.hp, self.eng, self.use_pdl
        qd = self.nh * hp.head_dim
        ids, pos, slot = (self.meta[i].data_ptr() for i in range(3))
        act = self.act.data_ptr()
        tp = self.tp > 1
        apdl = int(pdl and os.environ.get("GGB_BATCH_PDL", "0") != "0")    # RoPE / KV write and attention launched programmatically too

        def prep(x, norm, k, w, tiled=0):
            if tiled:
                return cabi.check(lib.ggb_act_prep_tiled(x.data_ptr(), norm.data_ptr() if norm is not None else 0, hp.eps, k, nb, act, pdl, s),
                                  "act_prep_tiled")
            cabi.check(lib.ggb_act_prep(x.data_ptr(), norm.data_ptr() if norm is not None else 0, hp.eps, k, nb,
                                        int(w.type == cabi.Q8_0), act, pdl, s), "act_prep")

        def gemv(segs, k, epi, residual=0, tiled=0):
            a = cabi.make_gemv_batch_args(segs, k, act, nb, epilogue=epi, residual=residual, use_pdl=pdl, act_tiled=tiled)
            cabi.check(lib.ggb_gemv_batch(C.byref(a), s), "gemv_batch")

        def row_split(x, w, k):
            """x += W h for a projection whose K is split across the ranks.  More than 8 tokens of a long vector (ffn_down) go
            through tiled activation images: one pass over the weights instead of two 8-token passes, same bits."""
            epi = cabi.EPI_STORE_F64 if tp else cabi.EPI_RESIDUAL
            y = self.y64.data_ptr() if tp else xp
            probe = cabi.make_gemv_batch_args([(w.ptr, w.type, w.rows, y)], k, act, nb, epilogue=epi)
            tiled = int(lib.ggb_gemv_batch_prefers_tiled(C.byref(probe)))
            prep(x, None, k, w, tiled)
            gemv([(w.ptr, w.type, w.rows, y)], k, epi, residual=0 if tp else xp, tiled=tiled)
            if tp:
                e.dist.all_reduce(self.y64[:nb], op=e.dist.ReduceOp.SUM, group=e.pg)
                cabi.check(lib.ggb_residual_add_f64(xp, self.y64.data_ptr(), nb * hp.d, 0, s), "residual_add_f64")

        cabi.check(lib.ggb_embed_rows(e.emb_type, e.emb_canon.data_ptr(), hp.d, ids, nb, self.x.data_ptr(), s), "embed_rows")
        xp = self.x.data_ptr()
        for i, L in enumerate(e.layers):
            prep(self.x, L["attn_norm"], hp.d, L["wq"])
            gemv([(L["wq"].ptr, L["wq"].type, L["wq"].rows, self.q.data_ptr()),
                  (L["wk"].ptr, L["wk"].type, L["wk"].rows, self.k.data_ptr()),
                  (L["wv"].ptr, L["wv"].type, L["wv"].rows, self.v.data_ptr())], hp.d, cabi.EPI_STORE)
            kc, vc = e.k_all[0, i].data_ptr(), e.v_all[0, i].data_ptr()
            cabi.check(lib.ggb_rope_kv_batch(self.q.data_ptr(), self.k.data_ptr(), self.v.data_ptr(), nb, pos, slot, self.slot_stride,
                                             self.nh, self.nkv, hp.head_dim, hp.n_rot, e.rope_tab.data_ptr(), kc, vc, s), "rope_kv_batch")
            cabi.check(lib.ggb_attn_decode_batch(self.q.data_ptr(), kc, vc, pos, slot, self.slot_stride, nb, self.nh, self.nkv,
                                                 hp.head_dim, e.n_ctx, self.att.data_ptr(), apdl, s), "attn_decode_batch")
            row_split(self.att, L["wo"], qd)
            prep(self.x, L["ffn_norm"], hp.d, L["wg"])
            gemv([(L["wg"].ptr, L["wg"].type, L["wg"].rows, self.h.data_ptr()),
                  (L["wu"].ptr, L["wu"].type, L["wu"].rows, 0)], hp.d, cabi.EPI_SWIGLU)
            row_split(self.h, L["wd"], self.ffl)
        if not head:      # prompt tokens whose logits nobody reads: the pass only fills the KV cache
            return
        prep(self.x, e.out_norm, hp.d, e.w_out)
        gemv([(e.w_out.ptr, e.w_out.type, e.w_out.rows, self.logits.data_ptr())], hp.d, cabi.EPI_STORE)
        if not tp:
            cabi.check(lib.ggb_argmax_rows(self.logits.data_ptr(), self.vl, nb, self.next_tok.data_ptr(), s), "argmax_rows")
            return
        cabi.check(lib.ggb_argmax_rows_key(self.logits.data_ptr(), self.vl, nb, e.tp_rank * self.vl, self.keys.data_ptr(), s), "argmax_rows_key")
        e.dist.all_reduce(self.keys[:nb], op=e.dist.ReduceOp.MAX, group=e.pg)
        cabi.check(lib.ggb_argmax_keys_unpack(self.keys.data_ptr(), nb, self.next_tok.data_ptr(), s), "argmax_keys_unpack")

    def launches_per_step(self, nb: int) -> int:
        passes = lambda k: -(-nb // 8)   # noqa: E731  (an upper bound: very large K runs more, smaller passes)
        per_layer = 4 + passes(0) * 4 + 2
        return nb + self.hp.n_layer * per_layer + 1 + passes(0) + 1

    # ------------------------------------------------------------------ public
    def step(self, entries, head: bool = True) -> list[int]:
        """entries: [(slot index, token id, position)].  Runs the tokens through the model, writes their K/V at the
        given positions and returns the greedy next token of each entry (logits stay on the device: logits_row).
        A slot may appear several times with CONSECUTIVE positions (a prompt chunk): all K/V of the batch are written
        before attention runs and entry j attends positions <= pos[j], so the result is the token-by-token one.
        head=False skips the lm-head (prompt chunks whose logits nobody reads); returns []."""
        return self.collect(self.launch(entries, head))

    def launch(self, entries, head: bool = True):
        """step() without the wait: enqueues the pass and the read-back of its greedy tokens, returns a handle for collect().
        At most two launches may be outstanding (the host buffers alternate), collected in launch order."""
        nb = len(entries)
        if not 0 < nb <= self.nb_max:
            raise ValueError(f"batch of {nb} entries (1..{self.nb_max})")
        e, torch = self.eng, self.torch
        slots = [en[0] for en in entries]
        seen = {}
        for sl, _, pos in entries:
            if sl in seen and pos != seen[sl] + 1:
                raise ValueError("entries of one slot must be consecutive positions in ascending order")
            seen[sl] = pos
        for sl, tok, pos in entries:
            if not (0 <= sl < len(e.slots) and 0 <= pos < e.n_ctx and 0 <= tok < self.hp.vocab):
                raise ValueError(f"bad batch entry (slot {sl}, token {tok}, position {pos})")
        mh = self.meta_host
        mh[0, :nb] = torch.tensor([en[1] for en in entries], dtype=torch.int32)
        mh[1, :nb] = torch.tensor([en[2] for en in entries], dtype=torch.int32)
        mh[2, :nb] = torch.tensor(slots, dtype=torch.int32)
        with torch.cuda.stream(self.stream):
            self.meta.copy_(mh, non_blocking=True)
        self._last = (nb, slots, [en[2] for en in entries]) if head and len(seen) == nb else None
        for sl, _, pos in entries:
            e.slots[sl].n_past = pos + 1
            e.slots[sl].chain_valid = False     # the slot's own device-side (token, position, x) no longer match
        return self._run(nb, head, False)

    def launch_chained(self):
        """The next decode step of the SAME sequences as the last launch, entirely from device state: token ids = that
        launch's arg-max results, positions + 1.  The host can therefore enqueue step k+1 before it has seen the tokens of
        step k (and emit them while k+1 runs).  Returns None when a sequence would leave its context."""
        if self._last is None:
            raise cabi.GGBError("launch_chained: no decode launch to continue")
        nb, slots, pos = self._last
        pos = [p + 1 for p in pos]
        if max(pos) >= self.eng.n_ctx:
            return None
        self._last = (nb, slots, pos)
        for sl, p in zip(slots, pos):
            self.eng.slots[sl].n_past = p + 1
        return self._run(nb, True, True)

    def _run(self, nb: int, head: bool, chained: bool):
        e, torch = self.eng, self.torch

        def enqueue(s):
            if chained:     # device-side hand-over of (token, position) from the previous step
                self.meta[0, :nb].copy_(self.next_tok[:nb])
                self.meta[1, :nb].add_(1)
            self._enqueue(nb, s, head)

        buf = self._seq & 1
        self._seq += 1
        with torch.cuda.stream(self.stream):
            if not e.use_graph:
                enqueue(self.stream.cuda_stream)
            else:
                g = self._graphs.get((nb, head, chained))
                if g is None:
                    if (nb, head) not in self._warm:   # (a chained launch always follows a plain one of the same size)
                        self._enqueue(nb, self.stream.cuda_stream, head)   # first use: sets kernel attributes outside capture
                        self.stream.synchronize()
                        self._warm.add((nb, head))
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=self.stream):
                        enqueue(torch.cuda.current_stream().cuda_stream)
                    self._graphs[(nb, head, chained)] = g
                g.replay()
            if head:
                self.next_host[buf].copy_(self.next_tok, non_blocking=True)
            self.events[buf].record(self.stream)
        return (buf, nb, head)

    def collect(self, handle) -> list[int]:
        buf, nb, head = handle
        self.events[buf].synchronize()
        return self.next_host[buf][:nb].tolist() if head else []

    def prefill(self, slot: int, tokens: list[int], start: int):
        """Write the K/V of `tokens` at positions start.. of one slot, 16 tokens per pass over the weights (bit-identical
        to feeding them one by one)."""
        for c0 in range(0, len(tokens), self.nb_max):
            chunk = tokens[c0:c0 + self.nb_max]
            self.step([(slot, int(t), start + c0 + j) for j, t in enumerate(chunk)], head=False)

    def warmup(self, n_seq: int | None = None):
        """Capture the graph of every batch shape the scheduler can ask for -- decode steps of 2..n_seq sequences (host-fed and
        chained), prompt chunks of 1..nb_max tokens -- so that no request pays for a capture (30-50 ms each on an 8B model:
        the first requests of a 16-stream burst otherwise see half a second of extra latency).  Leaves the slots reset."""
        e = self.eng
        if not e.use_graph:
            return
        n_seq = min(n_seq or len(e.slots), len(e.slots), self.nb_max)
        for nb in range(2, n_seq + 1):
            self.collect(self.launch([(s, 1, 0) for s in range(nb)]))
            h = self.launch_chained()
            if h is not None:
                self.collect(h)
        for nb in range(1, min(self.nb_max, e.n_ctx - 1) + 1):
            self.step([(0, 1, j) for j in range(nb)], head=False)
        self.stream.synchronize()
        for sl in e.slots:
            sl.reset()
        self._last = None

    def candidates(self, nb: int, k: int, cap: int = 256, extra=None):
        """per row of the last step's logits: (token ids, logits, logits of extra[b]) -- every logit >= the k-th largest, plus the raw
        logits of the tokens named in extra = {row: [ids]} (penalty windows) -- or None for a row whose ties overflowed the buffer;
        None altogether when the vocabulary is sharded.  One top-k launch, one gather launch when asked, one copy for all rows."""
        e, torch = self.eng, self.torch
        m = max((len(v) for v in extra.values()), default=0) if extra else 0
        if self.tp > 1 or not (0 < nb <= self.nb_max and 0 < k <= min(cap, self.vl)) or m > cap:
            return None
        NB = self.nb_max
        if getattr(self, "_cand_dev", None) is None or self._cand_cap != cap:
            self._cand_cap = cap
            self._cand_dev = torch.zeros(NB * (4 * cap + 1), dtype=torch.int32, device=e.dev)   # values | indices | extra values | extra ids | counts
            self._cand_host = torch.zeros(NB * (4 * cap + 1), dtype=torch.int32).pin_memory()
            self._cand_ids = torch.zeros(NB * cap, dtype=torch.int32).pin_memory()
        d = self._cand_dev
        p0 = d.data_ptr()
        with torch.cuda.stream(self.stream):
            cabi.check(self.lib.ggb_topk_rows(self.logits.data_ptr(), self.vl, nb, k, cap, p0, p0 + 4 * NB * cap, p0 + 16 * NB * cap,
                                              self.stream.cuda_stream), "topk_rows")
            if m:
                ids = self._cand_ids.numpy()[:nb * m].reshape(nb, m)
                ids[:] = -1
                for b, v in extra.items():
                    ids[b, :len(v)] = np.asarray(v, dtype=np.int32)
                d[3 * NB * cap:3 * NB * cap + nb * m].copy_(self._cand_ids[:nb * m], non_blocking=True)
                cabi.check(self.lib.ggb_gather_rows(self.logits.data_ptr(), self.vl, nb, p0 + 12 * NB * cap, m, p0 + 8 * NB * cap,
                                                    self.stream.cuda_stream), "gather_rows")
            self._cand_host.copy_(d, non_blocking=True)
        self.stream.synchronize()
        h = self._cand_host.numpy()
        out = []
        for b in range(nb):
            c = int(h[4 * NB * cap + b])
            if c > cap:
                out.append(None)
                continue
            ex = h[2 * NB * cap + b * m:2 * NB * cap + b * m + (len(extra.get(b, ())) if extra else 0)].copy().view(np.float32)
            out.append((h[NB * cap + b * cap:NB * cap + b * cap + c].copy(), h[b * cap:b * cap + c].copy().view(np.float32), ex))
        return out

    def logits_row_tensor(self, b: int):
        """row b of this rank's logits (its vocabulary shard under tensor parallelism), on the device, stream drained"""
        self.stream.synchronize()
        return self.logits[b]

    def logits_row(self, b: int) -> np.ndarray:
        if self.logits_host is None:
            self.logits_host = self.torch.zeros(self.vl, dtype=self.torch.float32).pin_memory()
        with self.torch.cuda.stream(self.stream):
            self.logits_host.copy_(self.logits[b], non_blocking=True)
        self.stream.synchronize()
        return self.logits_host.numpy()
