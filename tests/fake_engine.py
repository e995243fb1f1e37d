"""Test doubles for the HTTP/scheduler boundary on a CPU-only box (test infrastructure, never shipped).

OracleEngine exposes the duck-typed slot API the scheduler drives (reset / prefill / decode / read_last_token /
feed / read_logits) on top of the CPU oracle, so `-m "not gpu"` tests can exercise the real server, scheduler,
tokenizer and (when /root/reference is mounted) the reference's unchanged gateway end to end -- BASELINE.json
config 1 ("TinyLlama ... via gateway /v1/chat/completions, NGL=0 CPU") at test size.
"""
import numpy as np


class OracleSlot:
    def __init__(self, model, n_ctx, index=0):
        self.m, self.n_ctx, self.index = model, n_ctx, index
        self.n_past, self.logits, self.last = 0, None, 0
        self.chain_valid = True

    def reset(self):
        self.n_past = 0          # like the GPU slot: counters only, the K/V of earlier positions stays (prompt cache)

    def prefill(self, tokens, start_pos=None):
        if start_pos is not None:
            self.n_past = start_pos
        if self.n_past + len(tokens) >= self.n_ctx:
            raise ValueError("context")
        for t in tokens:
            self.logits = self.m.forward(int(t), self.n_past)
            self.n_past += 1
        self.last = int(np.argmax(self.logits))
        self.chain_valid = True

    def decode(self, n):
        assert self.chain_valid, "decode() after a batch step without feed()"
        for _ in range(n):
            self.prefill([self.last])

    def feed(self, tok):
        self.prefill([tok])

    def read_last_token(self):
        return self.last

    def read_logits(self):
        return self.logits

    def read_candidates(self, k, cap=256, extra_ids=None):
        """what Slot.read_candidates takes off the GPU (ggb_topk_rows + ggb_gather_rows), restated with numpy"""
        return _candidates(np.asarray(self.logits, dtype=np.float32), k, cap, extra_ids)


def _candidates(row, k, cap, extra_ids):
    idx = np.flatnonzero(row >= np.sort(row)[-k]).astype(np.int32)
    if idx.size > cap:
        return None
    np.random.default_rng(int(idx.sum())).shuffle(idx)          # the device emits them unordered
    ex = np.array([row[t] if 0 <= t < row.size else 0.0 for t in (extra_ids or [])], dtype=np.float32)
    return idx, row[idx], ex


class OracleEngine:
    def __init__(self, path, n_ctx=128, n_slots=1):
        from oracle import oracle as O
        self.slots = [OracleSlot(O.OracleLlama(path, n_ctx=n_ctx, mode="canon"), n_ctx, i) for i in range(n_slots)]
        self.batch_capable = n_slots > 1
        self.batch = OracleBatch(self)
        m = self.slots[0].m
        self.hp = type("HP", (), {"n_layer": getattr(m, "n_layer", 0), "d": getattr(m, "d", 0)})()   # what cli.main reports
        self.weight_bytes = 0

    def warmup(self):
        pass

    def embed(self, tokens, slot=0, pooling="mean"):
        """the oracle's pooled, L2-normalised final hidden states (what Engine.embed computes on the GPU)"""
        s = self.slots[slot]
        s.reset()
        hs = []
        for i, t in enumerate(tokens):
            hs.append(np.asarray(s.m.forward(int(t), i, return_hidden=True), dtype=np.float64))
        s.n_past = len(tokens)
        v = hs[-1] if pooling == "last" else np.mean(hs, axis=0)
        n = np.linalg.norm(v)
        return (v / n if n > 0 else v).astype(np.float32)

    def close(self):
        pass


class OracleBatch:
    """the duck-typed BatchDecoder (llama-gguf-inference_b200/batch.py): one oracle forward per entry"""

    nb_max = 16

    def __init__(self, eng):
        self.eng, self.rows, self.steps = eng, [], 0

    def prefill(self, slot, tokens, start):
        s = self.eng.slots[slot]
        for j, t in enumerate(tokens):
            s.logits = s.m.forward(int(t), start + j)
        s.n_past, s.chain_valid = start + len(tokens), False

    def step(self, entries):
        assert len({e[0] for e in entries}) == len(entries)
        self.rows, out = [], []
        for sl, tok, pos in entries:
            s = self.eng.slots[sl]
            assert pos == s.n_past
            s.logits = s.m.forward(int(tok), pos)
            s.n_past, s.chain_valid = pos + 1, False
            self.rows.append(s.logits)
            out.append(int(np.argmax(s.logits)))
        self.steps += 1
        return out

    # the asynchronous form the scheduler pipelines with (BatchDecoder.launch / launch_chained / collect)
    def launch(self, entries, head=True):
        out = self.step(entries)
        self._last = (list(entries), out)
        return out

    def launch_chained(self):
        entries, out = self._last
        nxt = [(sl, tok, pos + 1) for (sl, _, pos), tok in zip(entries, out)]
        if max(p for _, _, p in nxt) >= self.eng.slots[0].n_ctx:
            return None
        self.chained = getattr(self, "chained", 0) + 1
        return self.launch(nxt)

    def collect(self, handle):
        return handle

    def logits_row(self, b):
        return self.rows[b]

    def candidates(self, nb, k, cap=256, extra=None):
        return [_candidates(np.asarray(self.rows[b], dtype=np.float32), k, cap, (extra or {}).get(b)) for b in range(nb)]
