"""Slot scheduler: maps concurrent completion requests onto the engine's sequence slots.

Stands in for llama-server's slot / continuous-batching loop [UPSTREAM-MEM: tools/server/server.cpp]: requests
queue up, each free slot takes one, prompt tokens are fed, and every scheduler tick advances each active slot by
one token.  The gateway in front limits what arrives (`MAX_CONCURRENT_REQUESTS`, scripts/gateway.py:113,132);
`/health` never touches this thread, so probes are answered while a generation runs (gateway.py:326-376).

Greedy requests (temperature 0 / top_k 1) stay on the GPU: the arg-max kernel picks the token and the host only
reads it back.  Other requests read the logits back and sample on the host (temperature, top-k, top-p, seed).

The engine is duck-typed (`slots[i].reset/prefill/decode/read_last_token/feed/read_logits`), so the HTTP
boundary can be tested on a CPU box with a stand-in engine.
"""
from __future__ import annotations

import os
import queue
import threading
import time
from collections import deque
from dataclasses import dataclass, field

import numpy as np


@dataclass
class SamplingParams:
    temperature: float = 0.8
    top_k: int = 40
    top_p: float = 0.95
    seed: int | None = None
    min_p: float = 0.0                 # keep tokens with p >= min_p * p_max (0 = off)
    repeat_penalty: float = 1.0        # llama.cpp: logits of recent tokens divided (if > 0) or multiplied (if < 0) by this
    presence_penalty: float = 0.0      # OpenAI: subtracted once for every token seen in the window
    frequency_penalty: float = 0.0     # OpenAI: subtracted per occurrence in the window
    repeat_last_n: int = 64            # window of the three penalties (prompt tail + generated tokens)

    @property
    def penalised(self) -> bool:
        return self.repeat_penalty != 1.0 or self.presence_penalty != 0.0 or self.frequency_penalty != 0.0

    @property
    def arg_max(self) -> bool:
        return self.temperature <= 0.0 or self.top_k == 1

    @property
    def greedy(self) -> bool:
        """decided on the device: arg-max of the raw logits"""
        return self.arg_max and not self.penalised


@dataclass
class Request:
    prompt_ids: list
    max_tokens: int
    sampling: SamplingParams = field(default_factory=SamplingParams)
    stop: list = field(default_factory=list)
    ignore_eos: bool = False
    cache_prompt: bool = True       # reuse the slot's KV cache for the longest common prefix with its previous sequence
    embed: str = ""                 # "mean" / "last": an embedding request (no generation); the vector arrives as ("embedding", vec, n_tokens)
    events: "queue.Queue" = field(default_factory=queue.Queue)   # ("piece", text, token_id) | ("done", reason, usage) | ("error", msg)
    cancelled: threading.Event = field(default_factory=threading.Event)
    t_submit: float = field(default_factory=time.time)


def apply_penalties(logits: np.ndarray, sp: SamplingParams, history) -> np.ndarray:
    """repeat / presence / frequency penalties over the last repeat_last_n tokens (llama.cpp's `penalties` sampler
    [UPSTREAM-MEM: llama-sampling.cpp]): x = x / r if x > 0 else x * r, then x -= count * frequency + presence."""
    if not sp.penalised or not history:
        return logits
    win = list(history)[-sp.repeat_last_n:] if sp.repeat_last_n > 0 else list(history)
    ids, counts = np.unique(np.asarray(win, dtype=np.int64), return_counts=True)
    keep = (ids >= 0) & (ids < logits.size)
    ids, counts = ids[keep], counts[keep]
    x = logits.astype(np.float32).copy()
    if sp.repeat_penalty != 1.0:
        v = x[ids]
        x[ids] = np.where(v > 0, v / sp.repeat_penalty, v * sp.repeat_penalty)
    x[ids] -= counts.astype(np.float32) * np.float32(sp.frequency_penalty) + np.float32(sp.presence_penalty)
    return x


def _lcp(a, b) -> int:
    n = 0
    for x, y in zip(a, b):
        if x != y:
            break
        n += 1
    return n


TOPK_CAP = 256      # candidates per row the device hands back (ggb_topk_rows); top_k above TOPK_CAP - 16 reads the whole row


def _chain(idx: np.ndarray, xs: np.ndarray, sp: SamplingParams, rng: np.random.Generator) -> int:
    """top-p -> min-p -> temperature -> multinomial on the top-k logits xs (float64, descending) with token ids idx"""
    p = np.exp(xs - xs[0])
    p /= p.sum()
    if 0.0 < sp.top_p < 1.0:
        keep = int(np.searchsorted(np.cumsum(p), sp.top_p) + 1)
        idx, xs, p = idx[:keep], xs[:keep], p[:keep] / p[:keep].sum()
    if 0.0 < sp.min_p < 1.0:
        keep = max(1, int((p >= sp.min_p * p[0]).sum()))          # p is sorted, p[0] is the maximum
        idx, xs = idx[:keep], xs[:keep]
    xt = xs / max(sp.temperature, 1e-6)
    p = np.exp(xt - xt[0])
    p /= p.sum()
    return int(idx[rng.choice(len(idx), p=p)])


PENALTY_WINDOW_MAX = 128     # distinct window tokens whose logits the device gathers for a penalised request


def penalty_window(sp: SamplingParams, history) -> list:
    """the distinct token ids the penalties of this request touch (its last repeat_last_n tokens)"""
    if not sp.penalised or not history:
        return []
    win = list(history)[-sp.repeat_last_n:] if sp.repeat_last_n > 0 else list(history)
    return sorted(set(int(t) for t in win))


def device_topk_ok(sp: SamplingParams, history=None) -> bool:
    """the request can be sampled from what the device hands back -- the top-k candidates; with penalties the top-(k + window)
    candidates plus the logits of the window tokens, which determine the penalised top-k exactly (a token outside the window keeps
    its logit, so at most `window` of the tokens ranked above it can drop below it) -- when that fits the candidate buffer"""
    if sp.arg_max and not sp.penalised:
        return False
    if os.environ.get("GGB_DEVICE_TOPK", "1") == "0":
        return False
    k = 1 if sp.arg_max else sp.top_k
    w = len(penalty_window(sp, history)) if sp.penalised else 0
    return 0 < k and k + w <= TOPK_CAP - 16 and w <= PENALTY_WINDOW_MAX


def sample_from_candidates_penalised(idx, vals, win_ids, win_vals, sp: SamplingParams, rng, history) -> int:
    """sample_token() for a penalised request on the union of the top-(k + window) candidates and the window tokens: the same
    float32 penalty arithmetic on the same raw logits, the same top-k of the result (ties: lower token id first)"""
    raw = {int(i): np.float32(v) for i, v in zip(idx, vals)}
    raw.update({int(i): np.float32(v) for i, v in zip(win_ids, win_vals)})
    ids = np.array(sorted(raw), dtype=np.int64)
    small = np.array([raw[int(i)] for i in ids], dtype=np.float32)
    pos = {int(t): j for j, t in enumerate(ids)}
    win = (list(history)[-sp.repeat_last_n:] if sp.repeat_last_n > 0 else list(history)) if history else []
    win = [pos[int(t)] for t in win if int(t) in pos]                          # (ids outside the vocabulary were never gathered)
    sp_small = SamplingParams(**{**sp.__dict__, "repeat_last_n": 0})          # the window is already cut; ids are positions now
    return int(ids[sample_token(small, sp_small, rng, win)])


def sample_from_candidates(idx: np.ndarray, vals: np.ndarray, sp: SamplingParams, rng: np.random.Generator) -> int:
    """sample_token() on what ggb_topk_rows returned: every logit >= the k-th largest, unordered (ties: lower token id first)"""
    xs = vals.astype(np.float64)
    order = np.lexsort((idx, -xs))[: sp.top_k]
    return _chain(idx[order].astype(np.int64), xs[order], sp, rng)


def sample_token(logits: np.ndarray, sp: SamplingParams, rng: np.random.Generator, history=None) -> int:
    """penalties -> top-k -> top-p -> min-p -> temperature -> multinomial: the order of upstream's default sampler chain
    [UPSTREAM-MEM: common/sampling.cpp; the reference documents the parameters in docs/API_REFERENCE.md:369-379].
    The nucleus and min-p cut-offs are therefore taken on the UNtempered distribution; the temperature only reshapes the
    probabilities of the survivors."""
    logits = apply_penalties(logits, sp, history)
    if sp.arg_max:
        return int(np.argmax(logits))
    x = logits.astype(np.float64)
    k = sp.top_k if 0 < sp.top_k < x.size else x.size
    idx = np.flatnonzero(x >= np.partition(x, -k)[-k]) if k < x.size else np.arange(x.size)
    idx = idx[np.lexsort((idx, -x[idx]))][:k]                    # descending; ties: lower token id first
    return _chain(idx, x[idx], sp, rng)


class _Active:
    def __init__(self, req: Request, slot, tokenizer):
        from .tokenizer import StreamDecoder
        self.req, self.slot = req, slot
        self.dec = StreamDecoder(tokenizer)
        self.n_gen = 0
        self.text = ""
        self.sent = 0          # characters of self.text already emitted
        self.rng = np.random.default_rng(req.sampling.seed)
        n = max(0, req.sampling.repeat_last_n)
        self.history = deque(req.prompt_ids[-n:] if n else req.prompt_ids, maxlen=n or None)   # penalty window
        self.t_first = None
        self.t_start = time.time()


class Scheduler(threading.Thread):
    def __init__(self, engine, tokenizer, ignore_eos: bool = False, log=None):
        super().__init__(daemon=True, name="ggufb200-scheduler")
        self.engine, self.tok = engine, tokenizer
        self.ignore_eos = ignore_eos
        self.pending: deque[Request] = deque()
        self.cv = threading.Condition()
        self.active: dict[int, _Active] = {}
        self.stop_flag = False
        self.slot_tokens: dict[int, list] = {}      # tokens whose K/V a free slot still holds at positions 0..len-1 (prompt cache)
        self.log = log or (lambda *a: None)
        self.stats = {"requests": 0, "prompt_tokens": 0, "completion_tokens": 0, "decode_seconds": 0.0}
        self.fatal: str | None = None
        self._inflight = None                       # (BatchDecoder handle, [(slot index, _Active)]) of a launched, uncollected step

    # ------------------------------------------------------------------ public
    def submit(self, req: Request) -> Request:
        with self.cv:
            if self.fatal:
                req.events.put(("error", self.fatal))
                return req
            self.pending.append(req)
            self.cv.notify()
        return req

    def shutdown(self):
        with self.cv:
            self.stop_flag = True
            self.cv.notify()

    def idle_slots(self) -> int:
        return len(self.engine.slots) - len(self.active)

    # ------------------------------------------------------------------ loop
    def run(self):
        try:
            while True:
                with self.cv:
                    while not self.stop_flag and not self.pending and not self.active and self._inflight is None:
                        self.cv.wait()
                    if self.stop_flag:
                        break
                    admit = bool(self.pending) and len(self.active) < len(self.engine.slots)
                if admit or not self.active:
                    self._drain()                      # the batch composition is about to change: no step may be in flight
                with self.cv:
                    fresh = []
                    while self.pending and len(self.active) < len(self.engine.slots):
                        req = self.pending.popleft()
                        # the free slot whose cache shares the longest prefix with this prompt (llama-server's slot similarity)
                        free = max((i for i in range(len(self.engine.slots)) if i not in self.active),
                                   key=lambda i: (_lcp(self.slot_tokens.get(i, ()), req.prompt_ids) if req.cache_prompt else 0, -i))
                        self.active[free] = _Active(req, self.engine.slots[free], self.tok)
                        fresh.append(free)
                self._start_many(fresh)
                if len(self.active) >= 2 and getattr(self.engine, "batch_capable", False):
                    self._step_batch()
                else:
                    self._drain()
                    for i in list(self.active):
                        self._step(i)
        except Exception as e:  # a CUDA error is fatal: surface it to every waiter, then let the process die loudly
            self.fatal = f"engine failure: {e!r}"
            self.log(self.fatal)
            for a in self.active.values():
                a.req.events.put(("error", self.fatal))
            with self.cv:
                for r in self.pending:
                    r.events.put(("error", self.fatal))
                self.pending.clear()
            raise

    def _start_many(self, fresh: list[int]):
        """Requests admitted in the same scheduling round: prompts long enough for the tensor-core prefill are
        processed TOGETHER (one pass of GEMMs over the concatenated tokens, Engine.prefill_many), the rest one by one."""
        many = getattr(self.engine, "prefill_many", None)
        group, group_tokens = [], 0
        limit = getattr(self.engine, "prefill_chunk", 0)
        floor = getattr(self.engine, "gemm_prefill_min", 1 << 30)

        def flush():
            nonlocal group, group_tokens
            if len(group) > 1:
                jobs = []
                for i in group:
                    a = self.active[i]
                    common = self._reusable_prefix(i)
                    a.slot.reset()                         # counters only: cached K/V below `common` stays
                    jobs.append((a.slot.index, a.req.prompt_ids[common:], common))
                    if common:
                        self.stats["prompt_tokens_cached"] = self.stats.get("prompt_tokens_cached", 0) + common
                try:
                    many(jobs)
                except ValueError as e:                    # a malformed prompt in the group: each request answers for itself
                    self.log(f"grouped prefill rejected ({e}); retrying the requests one by one")
                    for i in group:
                        self.slot_tokens.pop(i, None)
                        self._start(i)
                    group, group_tokens = [], 0
                    return
                for i in group:
                    self._start(i, prefilled=True)
            else:
                for i in group:
                    self._start(i)
            group, group_tokens = [], 0

        for i in fresh:
            a = self.active[i]
            n = len(a.req.prompt_ids) - self._reusable_prefix(i)      # tokens that actually have to be processed
            if many is None or a.req.embed or n < floor or len(a.req.prompt_ids) + 1 >= a.slot.n_ctx or n > limit:
                self._start(i)
                continue
            if group_tokens + n > limit:
                flush()
            group.append(i)
            group_tokens += n
        flush()

    def _start(self, i: int, prefilled: bool = False):
        a = self.active[i]
        req, slot = a.req, a.slot
        n_ctx = slot.n_ctx
        if req.embed:                                      # embedding request: one pass over the prompt, no generation
            try:
                vec = self.engine.embed(req.prompt_ids, slot=slot.index, pooling=req.embed)
            except ValueError as e:
                req.events.put(("error", f"invalid request: {e}"))
            else:
                self.stats["requests"] += 1
                self.stats["prompt_tokens"] += len(req.prompt_ids)
                req.events.put(("embedding", vec, len(req.prompt_ids)))
            del self.active[i]
            self.slot_tokens.pop(i, None)                  # the slot's cached prefix is gone (embed() refilled the slot from 0)
            return
        if len(req.prompt_ids) + 1 >= n_ctx:
            req.events.put(("error", f"the request exceeds the available context size ({len(req.prompt_ids)} prompt tokens, context {n_ctx})"))
            del self.active[i]
            return
        req.max_tokens = max(0, min(req.max_tokens, n_ctx - len(req.prompt_ids) - 1))
        if not prefilled:
            common = self._reusable_prefix(i)
            slot.reset()                                   # counters only: the K/V of positions < common stay valid
            try:
                if common:
                    slot.prefill(req.prompt_ids[common:], start_pos=common)
                    self.stats["prompt_tokens_cached"] = self.stats.get("prompt_tokens_cached", 0) + common
                else:
                    slot.prefill(req.prompt_ids)
            except ValueError as e:                        # this request's input is unusable: it alone fails (HTTP 400)
                req.events.put(("error", f"invalid request: {e}"))
                del self.active[i]
                self.slot_tokens.pop(i, None)
                return
        a.fed = list(req.prompt_ids)                       # tokens whose K/V the slot holds
        self.slot_tokens.pop(i, None)
        self.stats["requests"] += 1
        self.stats["prompt_tokens"] += len(req.prompt_ids)
        a.t_decode0 = time.time()
        if req.max_tokens == 0:
            self._finish(i, "length")
            return
        self._emit(i, self._pick(a))

    def _reusable_prefix(self, i: int) -> int:
        """length of the prompt prefix whose K/V is already in slot i (at least one token is always processed, and short
        overlaps are not worth a different code path)"""
        a = self.active[i]
        if not a.req.cache_prompt:
            return 0
        n = min(_lcp(self.slot_tokens.get(i, ()), a.req.prompt_ids), len(a.req.prompt_ids) - 1)
        return n if n >= 16 else 0

    def _pick(self, a: _Active) -> int:
        """token produced by the step that just ran"""
        if a.req.sampling.greedy:
            return a.slot.read_last_token()
        sp = a.req.sampling
        if device_topk_ok(sp, a.history) and hasattr(a.slot, "read_candidates"):
            win = penalty_window(sp, a.history)
            cand = a.slot.read_candidates((1 if sp.arg_max else sp.top_k) + len(win), TOPK_CAP, win)   # k numbers instead of the vocabulary
            if cand is not None:
                self.stats["device_topk_tokens"] = self.stats.get("device_topk_tokens", 0) + 1
                if sp.penalised:
                    return sample_from_candidates_penalised(cand[0], cand[1], win, cand[2], sp, a.rng, a.history)
                return sample_from_candidates(cand[0], cand[1], sp, a.rng)
        return sample_token(a.slot.read_logits(), sp, a.rng, a.history)

    def _step(self, i: int):
        a = self.active.get(i)
        if a is None:
            return
        if a.req.cancelled.is_set():
            self._finish(i, "cancelled")
            return
        if a.req.sampling.greedy and getattr(a.slot, "chain_valid", True):
            a.slot.decode(1)            # the device already holds the next token, its position and embedding
        else:
            a.slot.feed(a.last_tok)     # sampled on the host, or the slot last advanced inside a batch
        a.fed.append(a.last_tok)
        self._emit(i, self._pick(a))

    def _step_batch(self):
        """Two or more busy slots: their next tokens form one batch -- one pass over the weights (batch.py).

        Pipelined by one step when every sequence is greedy: the arg-max of step k stays on the device and feeds step k+1
        (BatchDecoder.launch_chained), which is enqueued BEFORE the host waits for step k; detokenising and streaming the
        tokens of step k then overlaps the GPU work of step k+1.  A sequence that turns out to have ended at step k (EOS,
        stop string) was carried through step k+1 in vain -- that token is dropped, the K/V row it wrote lies beyond what the
        slot's prompt cache claims -- and the next launch is host-fed again with the new composition."""
        for i in list(self.active):
            if self.active[i].req.cancelled.is_set():
                self._finish(i, "cancelled")
        bd = self.engine.batch
        inf = self._inflight
        if inf is None:
            live = list(self.active)
            if len(live) < 2:
                for i in live:
                    self._step(i)
                return
            acts = [self.active[i] for i in live]
            entries = [(a.slot.index, a.last_tok, a.slot.n_past) for a in acts]
            for a in acts:
                a.fed.append(a.last_tok)
            self.stats["batched_steps"] = self.stats.get("batched_steps", 0) + 1
            self.stats["batched_tokens"] = self.stats.get("batched_tokens", 0) + len(live)
            pairs = list(zip(live, acts))
            if hasattr(bd, "launch"):
                self._inflight = (bd.launch(entries), pairs)           # collected by the next round, after its own launch
            else:
                self._deliver(bd.step(entries), pairs)
            return
        handle, pairs = inf
        nxt = None
        all_greedy = all(a.req.sampling.greedy for _, a in pairs)
        go_on = False
        if ([i for i, _ in pairs] == list(self.active) and all(self.active[i] is a for i, a in pairs)
                and all(a.n_gen + 1 < a.req.max_tokens for _, a in pairs)):
            with self.cv:
                waiting = bool(self.pending) and len(self.active) < len(self.engine.slots)
            go_on = not waiting
        if go_on and all_greedy:
            nxt = bd.launch_chained()                                   # step k+1 from device state; None at the context end
        self._inflight = None
        t0 = time.time()
        toks = bd.collect(handle)
        self.stats["gpu_wait_seconds"] = self.stats.get("gpu_wait_seconds", 0.0) + time.time() - t0
        decided = self._decide(toks, pairs)
        if go_on and not all_greedy and len(decided) == len(pairs) and all(a.slot.n_past + 1 < a.slot.n_ctx for _, a, _ in decided):
            # sampled sequences: the tokens are chosen (k candidates per row came off the GPU), so step k+1 is fed by the host and
            # enqueued BEFORE the tokens of step k are detokenised and streamed -- that work overlaps the GPU like the greedy chain's
            nxt = bd.launch([(a.slot.index, tok, a.slot.n_past) for _, a, tok in decided])
        for i, a, tok in decided:
            if self.active.get(i) is a:
                self._emit(i, tok)
        if nxt is not None:
            for i, a in pairs:
                if self.active.get(i) is a:
                    a.fed.append(a.last_tok)                            # the token step k+1 is processing
            self.stats["batched_steps"] = self.stats.get("batched_steps", 0) + 1
            self.stats["batched_tokens"] = self.stats.get("batched_tokens", 0) + len(pairs)
            self.stats["chained_steps"] = self.stats.get("chained_steps", 0) + 1
            self._inflight = (nxt, pairs)

    def _deliver(self, toks, pairs):
        for i, a, tok in self._decide(toks, pairs):
            if self.active.get(i) is a:
                self._emit(i, tok)

    def _decide(self, toks, pairs):
        """[(slot index, sequence, token)] of the step that was just collected, for the sequences still running: the device's
        arg-max for greedy ones, the host sampler (on the device's top-k candidates where the request allows) for the others"""
        bd = self.engine.batch
        out = []
        cands = None
        wins = {b: penalty_window(a.req.sampling, a.history) for b, (i, a) in enumerate(pairs)
                if self.active.get(i) is a and device_topk_ok(a.req.sampling, a.history)}
        if wins and hasattr(bd, "candidates"):
            k = max((1 if pairs[b][1].req.sampling.arg_max else pairs[b][1].req.sampling.top_k) + len(w) for b, w in wins.items())
            cands = bd.candidates(len(pairs), k, TOPK_CAP, wins)        # one launch + one small copy for every sampled row of the step
        for b, (i, a) in enumerate(pairs):
            if self.active.get(i) is not a:
                continue                                                # ended or cancelled while the step was in flight
            sp = a.req.sampling
            if sp.greedy:
                tok = toks[b]
            elif cands is not None and b in wins and cands[b] is not None:
                self.stats["device_topk_tokens"] = self.stats.get("device_topk_tokens", 0) + 1
                if sp.penalised:
                    tok = sample_from_candidates_penalised(cands[b][0], cands[b][1], wins[b], cands[b][2], sp, a.rng, a.history)
                else:
                    tok = sample_from_candidates(cands[b][0], cands[b][1], sp, a.rng)
            else:
                tok = sample_token(bd.logits_row(b), sp, a.rng, a.history)
            out.append((i, a, tok))
        return out

    def _drain(self):
        """wait for the step in flight (if any) and deliver its tokens"""
        inf, self._inflight = self._inflight, None
        if inf is not None:
            self._deliver(self.engine.batch.collect(inf[0]), inf[1])

    def _emit(self, i: int, tok: int):
        a = self.active[i]
        req = a.req
        a.last_tok = tok
        a.history.append(tok)
        if a.t_first is None:
            a.t_first = time.time()
        if tok in self.tok.eog and not (req.ignore_eos or self.ignore_eos):
            self._flush(a, final=True)
            self._finish(i, "stop")
            return
        a.n_gen += 1
        a.text += a.dec.push(tok)
        # stop strings: hold back text that could be the beginning of one
        hit = None
        for s in req.stop:
            j = a.text.find(s, max(0, a.sent - len(s)))
            if j >= 0 and (hit is None or j < hit):
                hit = j
        if hit is not None:
            a.text = a.text[:hit]
            self._flush(a, final=True, tok=tok)
            self._finish(i, "stop")
            return
        self._flush(a, final=False, tok=tok)
        if a.n_gen >= req.max_tokens:
            a.text += a.dec.flush()
            self._flush(a, final=True, tok=tok)
            self._finish(i, "length")

    def _flush(self, a: _Active, final: bool, tok: int = -1):
        hold = 0 if final else max((len(s) - 1 for s in a.req.stop), default=0)
        upto = len(a.text) - hold
        if upto > a.sent:
            a.req.events.put(("piece", a.text[a.sent:upto], tok))
            a.sent = upto
        elif not final and not a.req.stop:
            a.req.events.put(("piece", "", tok))   # token without printable text yet (partial UTF-8)

    def _finish(self, i: int, reason: str):
        a = self.active.pop(i)
        if getattr(a, "fed", None):
            self.slot_tokens[i] = a.fed                    # what the next request on this slot may reuse
        dt = time.time() - getattr(a, "t_decode0", a.t_start)
        self.stats["completion_tokens"] += a.n_gen
        self.stats["decode_seconds"] += dt
        usage = {"prompt_tokens": len(a.req.prompt_ids), "completion_tokens": a.n_gen,
                 "total_tokens": len(a.req.prompt_ids) + a.n_gen}
        timings = {"prompt_n": len(a.req.prompt_ids), "predicted_n": a.n_gen, "predicted_ms": dt * 1e3,
                   "predicted_per_second": (a.n_gen / dt) if dt > 0 else 0.0,
                   "ttft_ms": ((a.t_first or time.time()) - a.t_start) * 1e3}
        a.req.events.put(("done", reason, usage, timings))
