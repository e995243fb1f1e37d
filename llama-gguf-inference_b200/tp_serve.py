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


class TPLeader:
    """What rank 0 hands to the Scheduler instead of the Engine (same duck-typed surface)."""

    batch_capable = False                           # batched decode is single-GPU; tensor-parallel slots are time-sliced

    def __init__(self, engine, dist, group=None):
        self.engine, self.dist, self.group = engine, dist, group
        self.slots = [TPSlot(self, s) for s in engine.slots]

    def cmd(self, *msg):
        self.dist.broadcast_object_list([msg], src=0, group=self.group)

    def gather_logits(self, slot) -> np.ndarray:
        return _gather_logits(self.dist, self.group, slot)

    def shutdown(self):
        self.cmd("stop")


def _gather_logits(dist, group, slot) -> np.ndarray:
    part = slot.logits_tensor() if hasattr(slot, "logits_tensor") else None
    if part is None:                                # duck-typed test engines: numpy logits
        import torch
        part = torch.from_numpy(np.ascontiguousarray(slot.read_logits(), dtype=np.float32))
    parts = [part.new_empty(part.shape) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, part.contiguous(), group=group)
    return np.concatenate([p.float().cpu().numpy() for p in parts])


def follower_loop(engine, dist, group=None):
    """Ranks != 0: execute whatever rank 0's scheduler does, until it says stop."""
    while True:
        box = [None]
        dist.broadcast_object_list(box, src=0, group=group)
        op, *args = box[0]
        if op == "stop":
            return
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
