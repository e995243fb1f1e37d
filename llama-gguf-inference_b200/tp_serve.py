"""Serving a tensor-parallel engine: rank 0 owns the HTTP front and the scheduler, the other ranks mirror its engine calls.

Launched as one process per GPU (`torchrun --nproc-per-node N bin/llama-server ...`, csrc/peer.cu exchanges inside every
step), every rank must issue the SAME sequence of engine operations.  The scheduler only runs on rank 0, so its slot calls
go through TPSlot, which first broadcasts (operation, arguments) to the followers and then executes locally; the followers
sit in follower_loop().  Greedy tokens are identical on every rank (the sharded arg-max is all-reduced inside the step);
sampled requests need the full logits, which are vocabulary-sharded, so read_logits() is a collective (all_gather).
"""
from __future__ import annotations

import numpy as np


class TPSlot:
    def __init__(self, leader: "TPLeader", slot):
        self._l, self._s = leader, slot
        self.index, self.n_ctx = slot.index, slot.n_ctx

    @property
    def n_past(self):
        return self._s.n_past

    @property
    def chain_valid(self):
        return getattr(self._s, "chain_valid", True)

    def reset(self):
        self._l.cmd("reset", self.index)
        self._s.reset()

    def prefill(self, tokens, start_pos=None):
        self._l.cmd("prefill", self.index, [int(t) for t in tokens], start_pos)
        self._s.prefill(tokens, start_pos)

    def decode(self, n):
        self._l.cmd("decode", self.index, int(n))
        self._s.decode(n)

    def feed(self, tok):
        self._l.cmd("feed", self.index, int(tok))
        self._s.feed(tok)

    def read_last_token(self):
        return self._s.read_last_token()           # identical on every rank

    def read_logits(self):
        self._l.cmd("logits", self.index)
        return self._l.gather_logits(self._s)


class TPBatch:
    """BatchDecoder (batch.py) behind the same mirroring: every launch is announced, then executed on every rank (its
    all-reduces are collective).  Only rank 0 collects results; the followers' state advances with the launches."""

    def __init__(self, leader: "TPLeader", bd):
        self._l, self._bd = leader, bd
        self.nb_max = bd.nb_max

    def launch(self, entries, head: bool = True):
        entries = [(int(a), int(b), int(c)) for a, b, c in entries]
        self._l.cmd("b_launch", entries, bool(head))
        return self._bd.launch(entries, head)

    def launch_chained(self):
        self._l.cmd("b_chain")                       # whether the context end refuses it is decided identically on every rank
        return self._bd.launch_chained()

    def collect(self, handle):
        return self._bd.collect(handle)

    def step(self, entries, head: bool = True):
        return self.collect(self.launch(entries, head))

    def prefill(self, slot: int, tokens, start: int):
        self._l.cmd("b_prefill", int(slot), [int(t) for t in tokens], int(start))
        self._bd.prefill(slot, tokens, start)

    def logits_row(self, b: int) -> np.ndarray:
        self._l.cmd("b_logits", int(b))
        bd = self._bd
        return _gather_rows(self._l.dist, self._l.group, bd.logits_row_tensor(b) if hasattr(bd, "logits_row_tensor") else bd.logits_row(b))


class TPLeader:
    """What rank 0 hands to the Scheduler instead of the Engine (same duck-typed surface)."""

    def __init__(self, engine, dist, group=None):
        self.engine, self.dist, self.group = engine, dist, group
        self.slots = [TPSlot(self, s) for s in engine.slots]
        self.batch_capable = bool(getattr(engine, "batch_capable", False))   # concurrent requests share one pass over the shards
        self._batch = None

    @property
    def batch(self):
        if self._batch is None:
            self._batch = TPBatch(self, self.engine.batch)
        return self._batch

    def cmd(self, *msg):
        self.dist.broadcast_object_list([msg], src=0, group=self.group)

    def gather_logits(self, slot) -> np.ndarray:
        return _gather_logits(self.dist, self.group, slot)

    def shutdown(self):
        self.cmd("stop")


def _gather_rows(dist, group, part) -> np.ndarray:
    if isinstance(part, np.ndarray):                # duck-typed test engines: numpy logits
        import torch
        part = torch.from_numpy(np.ascontiguousarray(part, dtype=np.float32))
    parts = [part.new_empty(part.shape) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, part.contiguous(), group=group)
    return np.concatenate([p.float().cpu().numpy() for p in parts])


def _gather_logits(dist, group, slot) -> np.ndarray:
    return _gather_rows(dist, group, slot.logits_tensor() if hasattr(slot, "logits_tensor") else slot.read_logits())


def follower_loop(engine, dist, group=None):
    """Ranks != 0: execute whatever rank 0's scheduler does, until it says stop."""
    while True:
        box = [None]
        dist.broadcast_object_list(box, src=0, group=group)
        op, *args = box[0]
        if op == "stop":
            return
        if op.startswith("b_"):                     # batched steps (TPBatch): launched, never collected here
            bd = engine.batch
            if op == "b_launch":
                bd.launch(args[0], args[1])
            elif op == "b_chain":
                bd.launch_chained()
            elif op == "b_prefill":
                bd.prefill(args[0], args[1], args[2])
            elif op == "b_logits":
                _gather_rows(dist, group, bd.logits_row_tensor(args[0]) if hasattr(bd, "logits_row_tensor") else bd.logits_row(args[0]))
            else:
                raise RuntimeError(f"unknown operation from rank 0: {op!r}")
            continue
        slot = engine.slots[args[0]]
        if op == "reset":
            slot.reset()
        elif op == "prefill":
            slot.prefill(args[1], args[2])
        elif op == "decode":
            slot.decode(args[1])
        elif op == "feed":
            slot.feed(args[1])
        elif op == "logits":
            _gather_logits(dist, group, slot)
        else:
            raise RuntimeError(f"unknown operation from rank 0: {op!r}")
