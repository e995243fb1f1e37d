"""Tensor-parallel sharding of a llama GGUF model across the GPUs of one node (one process per GPU).

The reference's backend has no tensor parallelism of its own design; BASELINE.json config 4
asks for Llama-3-70B split over 2/4/8 B200s with an all-reduce after the attention-output and FFN-down projections:

  column-split (output rows)   attn_q / attn_k / attn_v by heads, ffn_gate / ffn_up by rows, output.weight by vocab
  row-split (K)                attn_output, ffn_down -- on 256-element super-block boundaries (K-quant constraint)
  replicated                   token_embd (one-row gather), all norm vectors, rope table

Each rank keeps its heads' KV cache.  Partial row-split products are exchanged as *unrounded f64 sums*
(GGB_EPI_STORE_F64 + f64 all-reduce), so the single rounding to f32 happens after the cross-rank sum and a
tensor-parallel run is bit-identical with the single-GPU run and with the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import gguf_reader as G

ROWS, COLS, FULL = "rows", "cols", "full"


@dataclass(frozen=True)
class Shard:
    kind: str
    lo: int = 0
    hi: int = 0


def check_divisible(hp, tp: int) -> None:
    if tp < 1:
        raise ValueError("tp size must be >= 1")
    if hp.n_head % tp or hp.n_kv % tp:
        raise ValueError(f"tp={tp} must divide n_head={hp.n_head} and n_kv={hp.n_kv}")
    if (hp.n_head // tp * hp.head_dim) % 256 or (hp.ff // tp) % 256 or hp.ff % tp:
        raise ValueError(f"tp={tp}: the K-slices of attn_output ({hp.n_head // tp * hp.head_dim}) and ffn_down ({hp.ff // tp}) "
                         "must be multiples of the 256-element super-block")
    if hp.vocab % tp or (hp.vocab // tp) % 2:
        raise ValueError(f"tp={tp} must divide the vocabulary ({hp.vocab}) into even shards")


def shard_of(name: str, hp, tp: int, rank: int) -> Shard:
    """which part of tensor `name` rank `rank` holds"""
    if tp == 1:
        return Shard(FULL)
    base = name.split(".")[-2] if name.startswith("blk.") else name.rsplit(".", 1)[0]
    hq, hkv = hp.n_head // tp * hp.head_dim, hp.n_kv // tp * hp.head_dim
    fl = hp.ff // tp
    if base == "attn_q":
        return Shard(ROWS, rank * hq, (rank + 1) * hq)
    if base in ("attn_k", "attn_v"):
        return Shard(ROWS, rank * hkv, (rank + 1) * hkv)
    if base == "attn_output":
        return Shard(COLS, rank * hq, (rank + 1) * hq)
    if base in ("ffn_gate", "ffn_up"):
        return Shard(ROWS, rank * fl, (rank + 1) * fl)
    if base == "ffn_down":
        return Shard(COLS, rank * fl, (rank + 1) * fl)
    if base == "output":
        v = hp.vocab // tp
        return Shard(ROWS, rank * v, (rank + 1) * v)
    return Shard(FULL)


def slice_canonical(raw: np.ndarray, qtype: int, k: int, rows: int, sh: Shard):
    """(bytes of the shard in canonical GGUF block layout, rows_local, k_local)"""
    _, be, bb = G.GGML_TYPES[qtype]
    if sh.kind == FULL:
        return raw, rows, k
    m = raw.reshape(rows, k // be, bb)
    if sh.kind == ROWS:
        return np.ascontiguousarray(m[sh.lo:sh.hi]).reshape(-1), sh.hi - sh.lo, k
    if sh.lo % be or sh.hi % be:
        raise ValueError("K-slice not on a block boundary")
    return np.ascontiguousarray(m[:, sh.lo // be:sh.hi // be]).reshape(-1), rows, sh.hi - sh.lo
