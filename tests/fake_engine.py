"""Test doubles for the HTTP/scheduler boundary on a CPU-only box (test infrastructure, never shipped).

OracleEngine exposes the duck-typed slot API the scheduler drives (reset / prefill / decode / read_last_token /
feed / read_logits) on top of the CPU oracle, so `-m "not gpu"` tests can exercise the real server, scheduler,
tokenizer and (when /root/reference is mounted) the reference's unchanged gateway end to end -- BASELINE.json
config 1 ("TinyLlama ... via gateway /v1/chat/completions, NGL=0 CPU") at test size.
"""
import numpy as np


class OracleSlot:
    def __init__(self, model, n_ctx):
        self.m, self.n_ctx = model, n_ctx
        self.n_past, self.logits, self.last = 0, None, 0

    def reset(self):
        self.m.reset()
        self.n_past = 0

    def prefill(self, tokens, start_pos=None):
        if self.n_past + len(tokens) >= self.n_ctx:
            raise ValueError("context")
        for t in tokens:
            self.logits = self.m.forward(int(t), self.n_past)
            self.n_past += 1
        self.last = int(np.argmax(self.logits))

    def decode(self, n):
        for _ in range(n):
            self.prefill([self.last])

    def feed(self, tok):
        self.prefill([tok])

    def read_last_token(self):
        return self.last

    def read_logits(self):
        return self.logits


class OracleEngine:
    def __init__(self, path, n_ctx=128, n_slots=1):
        from oracle import oracle as O
        self.slots = [OracleSlot(O.OracleLlama(path, n_ctx=n_ctx, mode="canon"), n_ctx) for _ in range(n_slots)]
