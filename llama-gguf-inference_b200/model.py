"""Decode engine: GGUF file -> HBM-resident tile-SoA weights -> CUDA-graph-replayed decode steps.

Replaces the model-execution core of the reference's backend process (`/app/llama-server -m MODEL -c CTX
-ngl NGL`, /root/reference/scripts/start.sh:473-480).  PyTorch is used for device memory, streams and graph
capture only; every arithmetic operation is a kernel of libggufb200.so reached through cabi.py.

Per decoded token the engine enqueues, per layer, five launches (see csrc/gemv.cu):
    QKV   : rms_norm*attn_norm -> Q8_K -> {Wq,Wk,Wv} GEMV -> RoPE(q,k) -> f16 KV-cache write
    ATTN  : 8-CTA cluster per head, two-pass softmax over the cache through distributed shared memory
    O     : Q8_K(attn) -> Wo GEMV -> + residual
    GATEUP: rms_norm*ffn_norm -> Q8_K -> {Wgate,Wup} GEMV -> silu(g)*u
    DOWN  : Q8_K(h) -> Wdown GEMV -> + residual
then  HEAD: rms_norm*output_norm -> Q8_K -> Woutput GEMV -> arg-max partials -> token, pos+1, next embedding.
Position, token and step counters live in device memory, so one captured graph serves every step.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from dataclasses import dataclass

import numpy as np

from . import cabi
from . import gguf_reader as G
from . import parallel

SM_SMEM = 227 * 1024            # shared memory an SM can hand out to CTAs (B200)
HALF_SM_SMEM = SM_SMEM // 2 + 512
LONG_SEQ = 2560                 # positions from which the decode attention runs as launches over the whole GPU (measured crossover ~2 600 on Llama-3-8B)

QUANT_TYPES = (G.GGML_Q4_K, G.GGML_Q5_K, G.GGML_Q6_K, G.GGML_Q8_0)
LEGACY_TYPES = (G.GGML_Q4_0, G.GGML_Q5_0)


def legacy_to_q8_0(raw: np.ndarray, ggml_type: int) -> np.ndarray:
    """Q4_0 / Q5_0 blocks (32 weights: f16 d, 4- / 5-bit codes, value d * (q - 8) / d * (q - 16)) as Q8_0 blocks with the SAME
    scale and the int8 codes q - 8 / q - 16.  Exact: the dequantised weights are identical, and ggml pairs both formats with
    Q8_0 activations and computes sum_i * d_w * d_a per block exactly as for Q8_0 weights (ggml_vec_dot_q4_0_q8_0 /
    q5_0_q8_0 [UPSTREAM-MEM: ggml-cpu/quants.c]) -- so the Q8_0 kernels give the very terms of the legacy dot.  Costs 34
    instead of 18 / 22 bytes per block in HBM; these formats are the rare case."""
    bb = 18 if ggml_type == G.GGML_Q4_0 else 22
    b = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1, bb)
    out = np.empty((b.shape[0], 34), dtype=np.uint8)
    out[:, 0:2] = b[:, 0:2]
    qs = b[:, bb - 16:]
    lo, hi = (qs & 0x0F).astype(np.int16), (qs >> 4).astype(np.int16)       # elements 0..15 and 16..31
    if ggml_type == G.GGML_Q4_0:
        codes = np.concatenate([lo, hi], axis=1) - 8
    else:
        qh = b[:, 2:6].copy().view(np.uint32).reshape(-1, 1)
        bits = ((qh >> np.arange(32, dtype=np.uint32)) & 1).astype(np.int16)  # bit j = fifth bit of element j
        codes = (np.concatenate([lo, hi], axis=1) | (bits << 4)) - 16
    out[:, 2:] = codes.astype(np.int8).view(np.uint8)
    return out.reshape(-1)


@dataclass
class HParams:
    n_layer: int
    d: int
    ff: int
    n_head: int
    n_kv: int
    head_dim: int
    n_rot: int
    eps: float
    rope_base: float
    vocab: int
    ctx_train: int

    @classmethod
    def from_gguf(cls, f: G.GGUFFile) -> "HParams":
        arch = f.get("general.architecture")
        if arch != "llama":
            raise G.GGUFError(f"unsupported architecture {arch!r} (this engine implements the llama graph)")
        a = lambda k, d=None: f.get(f"{arch}.{k}", d)  # noqa: E731
        d = int(a("embedding_length"))
        n_head = int(a("attention.head_count"))
        hd = int(a("attention.key_length", d // n_head))
        emb = f.tensors["token_embd.weight"]
        return cls(n_layer=int(a("block_count")), d=d, ff=int(a("feed_forward_length")), n_head=n_head,
                   n_kv=int(a("attention.head_count_kv", n_head)), head_dim=hd, n_rot=int(a("rope.dimension_count", hd)),
                   eps=float(a("attention.layer_norm_rms_epsilon", 1e-5)), rope_base=float(a("rope.freq_base", 10000.0)),
                   vocab=int(emb.ne[1]), ctx_train=int(a("context_length", 4096)))


def rope_table(n_ctx: int, n_rot: int, base: float, freq_factors: np.ndarray | None = None) -> np.ndarray:
    """[n_ctx, n_rot/2, 2] (cos, sin) in f32, the way ggml's rope cache is filled [UPSTREAM-MEM: ggml-cpu/ops.cpp]:
    theta_0 = pos, theta_{i+1} = theta_i * base^(-2/n_rot), every step rounded to f32."""
    scale = np.float32(np.power(np.float64(np.float32(base)), np.float64(np.float32(-2.0) / np.float32(n_rot))))
    theta = np.empty((n_ctx, n_rot // 2), dtype=np.float32)
    theta[:, 0] = np.arange(n_ctx, dtype=np.float32)
    for i in range(1, n_rot // 2):
        theta[:, i] = theta[:, i - 1] * scale
    if freq_factors is not None:
        theta = (theta / freq_factors.astype(np.float32)[None, :]).astype(np.float32)
    t64 = theta.astype(np.float64)
    return np.stack([np.cos(t64), np.sin(t64)], axis=-1).astype(np.float32)


class Weight:
    """One 2-D weight in HBM, tile-SoA layout."""

    def __init__(self, tensor, qtype: int, rows: int, k: int):
        self.t, self.type, self.rows, self.k = tensor, qtype, rows, k

    @property
    def ptr(self) -> int:
        return self.t.data_ptr()


class _Stager:
    """Pipelined weight upload (the load path behind start.sh's 30 s /health gate, scripts/start.sh:600-635): reader
    threads copy row chunks of the memory-mapped GGUF file into a ring of pinned staging buffers while the chunks before
    them travel host -> device (async copies from pinned memory) and are re-ordered into the tile-SoA layout on the GPU.
    Buffers are handed out by the consumer in order, so a chunk is always uploaded in the order it was requested."""

    _pinned: dict = {}
    pin_seconds = 0.0

    def __init__(self, torch, lib, stream, dev, n_buf: int = 6, buf_bytes: int = 64 << 20, threads: int = 4):
        import concurrent.futures as cf
        from collections import deque
        self.torch, self.lib, self.stream, self.dev = torch, lib, stream, dev
        self.buf_bytes = buf_bytes
        key = (n_buf, buf_bytes)
        if key not in _Stager._pinned:           # pinning host memory is slow: the ring is allocated once per process
            t0 = time.time()
            _Stager._pinned[key] = [torch.empty(buf_bytes, dtype=torch.uint8).pin_memory() for _ in range(n_buf)]
            _Stager.pin_seconds = time.time() - t0
        self.bufs = _Stager._pinned[key]
        self.nps = [b.numpy() for b in self.bufs]
        self.evs = [torch.cuda.Event() for _ in range(n_buf)]
        self.used = [False] * n_buf
        self.free = deque(range(n_buf))
        self.pending = deque()
        self.pool = cf.ThreadPoolExecutor(max_workers=threads)
        with torch.cuda.stream(stream):
            self.dev_tmp = torch.empty(buf_bytes, dtype=torch.uint8, device=dev)   # canonical bytes of the chunk being re-ordered
        self.bytes = 0

    def _consume(self):
        fut, idx, n, finish = self.pending.popleft()
        fut.result()
        with self.torch.cuda.stream(self.stream):
            finish(self.bufs[idx][:n])
            self.evs[idx].record(self.stream)
        self.used[idx] = True
        self.free.append(idx)

    def submit(self, src: np.ndarray, finish):
        """src: a uint8 view of at most buf_bytes of the mapped file; finish(pinned_tensor) enqueues what happens to it"""
        n = src.size
        assert n <= self.buf_bytes
        if not self.free:
            self._consume()
        idx = self.free.popleft()
        if self.used[idx]:
            self.evs[idx].synchronize()        # the async copy that last read this buffer has run
        fut = self.pool.submit(np.copyto, self.nps[idx][:n], src)
        self.pending.append((fut, idx, n, finish))
        self.bytes += n

    def drain(self):
        while self.pending:
            self._consume()
        self.stream.synchronize()

    def close(self):
        self.drain()
        self.pool.shutdown()
        self.nps, self.dev_tmp = [], None


class Slot:
    """One sequence: KV cache, activations, device-side counters and the captured graphs that advance it."""

    def __init__(self, eng: "Engine", index: int):
        self.eng, self.index = eng, index
        torch, hp, dev = eng.torch, eng.hp, eng.dev
        self.torch, self.lib, self.hp = torch, eng.lib, hp
        self.n_ctx, self.max_new, self.use_pdl = eng.n_ctx, eng.max_new, eng.use_pdl
        self.stream = eng.stream
        tp = eng.tp_size
        self.nh, self.nkv, self.ffl, self.vl = hp.n_head // tp, hp.n_kv // tp, hp.ff // tp, hp.vocab // tp   # this rank's share
        kvd = self.nkv * hp.head_dim
        self.kc, self.vc = eng.k_all[index], eng.v_all[index]   # [n_layer][n_ctx][kvd] views of the engine's cache pool
        self.chain_valid = True   # device-side (token, position, x) describe this slot's next decode step
        f32 = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)  # noqa: E731
        i32 = lambda n: torch.zeros(n, dtype=torch.int32, device=dev)  # noqa: E731
        self.x, self.q, self.attn = f32(hp.d), f32(self.nh * hp.head_dim), f32(self.nh * hp.head_dim)
        self.h, self.logits = f32(self.ffl), f32(self.vl)
        self.y64 = torch.zeros(hp.d, dtype=torch.float64, device=dev)       # row-split partial sums (tensor parallel)
        self.key = torch.zeros(1, dtype=torch.int64, device=dev)            # sharded arg-max key
        self.attn_ws = torch.zeros(max(256, self.lib.ggb_attn_decode_ws_bytes_ctx(self.nh, self.nkv, hp.head_dim, eng.n_ctx)), dtype=torch.uint8, device=dev)
        self.split_ok = self.attn_ws.numel() > 256     # the long-sequence attention path exists for this geometry and context
        self._long = 0
        self._tokpos = i32(2)      # token id and position side by side: the host sets both with ONE 8-byte copy
        self.tok_dev, self.pos_dev, self.step_dev = self._tokpos[0:1], self._tokpos[1:2], i32(1)
        self.out_tokens = i32(self.max_new)
        self.part_val, self.part_idx = f32(1), i32(1)
        a = self._head_args()
        self.n_part = self.lib.ggb_gemv_grid(C.byref(a))
        self.part_val, self.part_idx = f32(self.n_part), i32(self.n_part)
        # pinned staging for (token, position): a ring with one event per entry, so that a host-driven step needs no
        # synchronisation of its own before the buffer can be rewritten (the read-back of the token is the step's only sync)
        self.host_ring = torch.zeros((8, 2), dtype=torch.int32).pin_memory()
        self.host_ring_np = self.host_ring.numpy()      # the same pinned memory: plain stores instead of tensor indexing
        self.host_ev = [torch.cuda.Event() for _ in range(8)]
        self.host_used = [False] * 8
        self.host_i = 0
        self.host_tok = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.host_tok_np = self.host_tok.numpy()
        self.host_logits = None
        self._graphs = {}
        # every GEMV of this model has k % 2048 == 0: all launches share the kernel instance without partial-tile code
        self.full_k = int(all(k % 2048 == 0 for k in (hp.d, self.nh * hp.head_dim, self.ffl)))
        self._build_args()
        self.n_past = 0   # host mirror of the number of positions held in the KV cache

    # ------------------------------------------------------------------ launch descriptions
    def _head_args(self):
        hp, e = self.hp, self.eng
        return cabi.make_gemv_args(
            [(e.w_out.ptr, e.w_out.type, e.w_out.rows, self.logits.data_ptr())], hp.d, self.x.data_ptr(),
            prologue=cabi.PRO_RMSNORM, epilogue=cabi.EPI_ARGMAX, norm_w=e.out_norm.data_ptr(), eps=hp.eps,
            use_pdl=self.use_pdl, part_val=self.part_val.data_ptr(), part_idx=self.part_idx.data_ptr(),
            full_k_model=getattr(self, "full_k", 0))

    def _build_args(self):
        hp, e = self.hp, self.eng
        self._layer_args = []
        self._attn_pdl = []
        for i, L in enumerate(e.layers):
            qkv = cabi.make_gemv_args(
                [(L["wq"].ptr, L["wq"].type, L["wq"].rows, self.q.data_ptr()),
                 (L["wk"].ptr, L["wk"].type, L["wk"].rows, 0),
                 (L["wv"].ptr, L["wv"].type, L["wv"].rows, 0)],
                hp.d, self.x.data_ptr(), prologue=cabi.PRO_RMSNORM, epilogue=cabi.EPI_ROPE_KV,
                norm_w=L["attn_norm"].data_ptr(), eps=hp.eps, use_pdl=self.use_pdl, pos_dev=self.pos_dev.data_ptr(),
                rope_tab=e.rope_tab.data_ptr(), n_rot=hp.n_rot, head_dim=hp.head_dim,
                kcache=self.kc[i].data_ptr(), vcache=self.vc[i].data_ptr(), full_k_model=self.full_k)
            tp = e.tp_size > 1
            # row-split projections: partial sums leave through the fused peer exchange, or as a local f64 vector (NCCL)
            epi_rs = (cabi.EPI_PEER_F64 if e.peer else cabi.EPI_STORE_F64) if tp else cabi.EPI_RESIDUAL
            peer = (e.peer["bases"], e.tp_rank, hp.d) if (tp and e.peer) else None
            o = cabi.make_gemv_args(
                [(L["wo"].ptr, L["wo"].type, L["wo"].rows, self.y64.data_ptr() if tp else self.x.data_ptr())], L["wo"].k,
                self.attn.data_ptr(), prologue=cabi.PRO_PLAIN, epilogue=epi_rs,
                residual=self.x.data_ptr(), use_pdl=self.use_pdl, peer=peer, full_k_model=self.full_k)
            gu = cabi.make_gemv_args(
                [(L["wg"].ptr, L["wg"].type, L["wg"].rows, self.h.data_ptr()),
                 (L["wu"].ptr, L["wu"].type, L["wu"].rows, 0)],
                hp.d, self.x.data_ptr(), prologue=cabi.PRO_RMSNORM, epilogue=cabi.EPI_SWIGLU,
                norm_w=L["ffn_norm"].data_ptr(), eps=hp.eps, use_pdl=self.use_pdl, full_k_model=self.full_k)
            dn = cabi.make_gemv_args(
                [(L["wd"].ptr, L["wd"].type, L["wd"].rows, self.y64.data_ptr() if tp else self.x.data_ptr())], L["wd"].k,
                self.h.data_ptr(), prologue=cabi.PRO_PLAIN, epilogue=epi_rs,
                residual=self.x.data_ptr(), use_pdl=self.use_pdl, peer=peer, full_k_model=self.full_k)
            # The attention releases the output projection early (it then fills its weight ring while the attention runs)
            # only if two CTAs of the projection cannot land on one SM: pad its shared memory past half an SM's when the
            # gate/up launch behind it still fits next to it (csrc/attn.cu, csrc/gemv.cu: min_smem).
            so, sg = self.lib.ggb_gemv_smem_bytes(C.byref(o)), self.lib.ggb_gemv_smem_bytes(C.byref(gu))
            trig = 0
            if self.use_pdl and so > 0 and sg > 0:
                if so >= HALF_SM_SMEM:
                    trig = 2
                elif HALF_SM_SMEM + 2048 + sg + 2048 <= SM_SMEM:
                    o.min_smem = HALF_SM_SMEM + 2048
                    trig = 2
            self._attn_pdl.append(self.use_pdl | trig)
            self._layer_args.append((qkv, o, gu, dn))
        self._head = self._head_args()

    # ------------------------------------------------------------------ enqueue
    def _enqueue_embed(self, s: int):
        e = self.eng
        cabi.check(self.lib.ggb_embed_row(e.emb_type, e.emb_canon.data_ptr(), self.hp.d, self.tok_dev.data_ptr(),
                                          self.x.data_ptr(), s), "embed_row")

    def _allreduce_residual(self, s: int):
        """x += (float) sum over ranks of the f64 partials (the exchange step of a row-split projection)"""
        e = self.eng
        if e.peer:   # the GEMV already pushed its partials into every rank's region (GGB_EPI_PEER_F64)
            cabi.check(self.lib.ggb_peer_reduce_residual(self.x.data_ptr(), e.peer["own"], e.tp_size, self.hp.d, self.hp.d,
                                                         self.use_pdl, s), "peer_reduce_residual")
            return
        e.dist.all_reduce(self.y64, op=e.dist.ReduceOp.SUM, group=e.pg)
        cabi.check(self.lib.ggb_residual_add_f64(self.x.data_ptr(), self.y64.data_ptr(), self.hp.d, 0, s), "residual_add_f64")

    def _enqueue_layers(self, s: int):
        hp, lib = self.hp, self.lib
        tp = self.eng.tp_size > 1
        for i, (qkv, o, gu, dn) in enumerate(self._layer_args):
            cabi.check(lib.ggb_gemv(C.byref(qkv), s), "gemv qkv")
            cabi.check(lib.ggb_attn_decode(self.q.data_ptr(), self.kc[i].data_ptr(), self.vc[i].data_ptr(),
                                           self.pos_dev.data_ptr(), self.nh, self.nkv, hp.head_dim, self.n_ctx,
                                           self.attn_ws.data_ptr(), self.attn.data_ptr(), self._attn_pdl[i] | self._long, s), "attn_decode")
            cabi.check(lib.ggb_gemv(C.byref(o), s), "gemv o")
            if tp:
                self._allreduce_residual(s)
            cabi.check(lib.ggb_gemv(C.byref(gu), s), "gemv gate/up")
            cabi.check(lib.ggb_gemv(C.byref(dn), s), "gemv down")
            if tp:
                self._allreduce_residual(s)

    def _enqueue_head(self, s: int):
        lib, e = self.lib, self.eng
        cabi.check(lib.ggb_gemv(C.byref(self._head), s), "gemv head")
        if e.tp_size > 1:   # vocabulary-sharded arg-max: one sortable key per rank, MAX across ranks
            cabi.check(lib.ggb_argmax_pack(self.part_val.data_ptr(), self.part_idx.data_ptr(), self.n_part,
                                           e.tp_rank * self.vl, self.key.data_ptr(), s), "argmax_pack")
            e.dist.all_reduce(self.key, op=e.dist.ReduceOp.MAX, group=e.pg)
            cabi.check(lib.ggb_argmax_unpack_next(self.key.data_ptr(), self.tok_dev.data_ptr(), self.pos_dev.data_ptr(),
                                                  self.step_dev.data_ptr(), self.out_tokens.data_ptr(), self.max_new, e.emb_type,
                                                  e.emb_canon.data_ptr(), self.hp.d, self.x.data_ptr(), s), "argmax_unpack_next")
            return
        cabi.check(lib.ggb_argmax_next(self.part_val.data_ptr(), self.part_idx.data_ptr(), self.n_part,
                                       self.tok_dev.data_ptr(), self.pos_dev.data_ptr(), self.step_dev.data_ptr(),
                                       self.out_tokens.data_ptr(), self.max_new, e.emb_type,
                                       e.emb_canon.data_ptr(), self.hp.d, self.x.data_ptr(), s), "argmax_next")

    def _run(self, kind: str, pos: int = 0):
        """kind: 'prompt' (embed + layers), 'prompt_last' (embed + layers + head), 'decode' (layers + head).
        pos: the position of the token this step processes -- from LONG_SEQ positions on the attention runs as launches over
        the whole GPU (csrc/attn.cu, use_pdl bit 2), a separately captured graph per kind."""
        torch = self.torch
        self._long = 4 if (self.split_ok and pos >= LONG_SEQ) else 0
        key = (kind, self._long)

        def body(s):
            if kind != "decode":
                self._enqueue_embed(s)
            self._enqueue_layers(s)
            if kind != "prompt":
                self._enqueue_head(s)

        if not self.eng.use_graph:
            body(self.stream.cuda_stream)
            return
        g = self._graphs.get(key)
        if g is None:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=self.stream):
                body(torch.cuda.current_stream().cuda_stream)
            self._graphs[key] = g
        g.replay()

    # ------------------------------------------------------------------ sequence API
    def reset(self):
        with self.torch.cuda.stream(self.stream):
            self.pos_dev.zero_()
            self.step_dev.zero_()
        self.n_past = 0

    def _set_tok_pos(self, tok: int, pos: int):
        i = self.host_i
        self.host_i = (i + 1) % len(self.host_ev)
        if self.host_used[i]:
            self.host_ev[i].synchronize()     # the copy that last read this entry has run (8 entries ago: normally long done)
        self.host_ring_np[i, 0] = tok
        self.host_ring_np[i, 1] = pos
        self._tokpos.copy_(self.host_ring[i], non_blocking=True)
        self.host_ev[i].record(self.torch.cuda.current_stream())
        self.host_used[i] = True

    def warmup(self):
        """Run every graph once (sets kernel attributes, captures graphs), then clear the state."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            graph = self.eng.use_graph
            self.eng.use_graph = False
            self._set_tok_pos(0, 0)
            self._run("prompt_last")
            self.stream.synchronize()
            self.eng.use_graph = graph
            if graph:
                for kind in ("prompt", "prompt_last", "decode"):
                    self._set_tok_pos(0, 0)
                    self._run(kind)
                self.stream.synchronize()
        self.reset()

    def prefill(self, tokens: list[int], start_pos: int | None = None):
        """Feed tokens at positions start_pos.. (default: append).  Long prompts go through the tcgen05 GEMM path
        (tolerance-level numerics, like upstream's batched CUDA path); short ones run token by token through the
        decode kernels (bit-exact integer path).  The last token also runs the head, which emits the first
        generated token (greedy) and the logits."""
        if not tokens:
            raise ValueError("empty prompt")
        self.eng.check_tokens(tokens)
        start = self.n_past if start_pos is None else start_pos
        if start + len(tokens) >= self.n_ctx:
            raise ValueError(f"prompt of {len(tokens)} tokens does not fit the context ({self.n_ctx})")
        e = self.eng
        if e.tp_size == 1 and len(tokens) >= e.gemm_prefill_min:
            for c0 in range(0, len(tokens), e.prefill_chunk):
                # only the final chunk runs the head: it emits the first generated token and advances the step counter
                e.prefill_many([(self.index, tokens[c0:c0 + e.prefill_chunk], start + c0)], head=c0 + e.prefill_chunk >= len(tokens))
            return
        first = 0
        if len(tokens) >= 4:
            # all but the last token only fill the KV cache: 16 per pass through the batched kernels (same arithmetic;
            # under tensor parallelism every rank runs the same passes, csrc exchange = f64 all-reduce, batch.py)
            e.batch.prefill(self.index, tokens[:-1], start)
            first = len(tokens) - 1
        with self.torch.cuda.stream(self.stream):
            for i in range(first, len(tokens)):
                self._set_tok_pos(int(tokens[i]), start + i)
                self._run("prompt_last" if i == len(tokens) - 1 else "prompt", start + i)
        self.n_past = start + len(tokens)
        self.chain_valid = True

    def decode(self, n_steps: int):
        """Enqueue n greedy decode steps (no host synchronisation); tokens land in out_tokens."""
        if self.n_past + n_steps >= self.n_ctx + 1:
            raise ValueError("context window exhausted")
        if not self.chain_valid:
            raise RuntimeError("this slot last advanced inside a batch: feed() its newest token before decode()")
        with self.torch.cuda.stream(self.stream):
            for j in range(n_steps):
                self._run("decode", self.n_past + j)
        self.n_past += n_steps

    def feed(self, tok: int):
        """Sampled / host-driven decoding: the host chose `tok`; run it through the model and produce logits and the next
        greedy token.  One 8-byte H2D copy (token, position) from pinned memory + one graph launch, no host synchronisation."""
        if self.eng.tp_size > 1:
            return self.prefill([tok])
        if isinstance(tok, bool) or not 0 <= int(tok) < self.hp.vocab:
            raise ValueError(f"token id {tok!r} is outside the vocabulary (0..{self.hp.vocab - 1})")
        if self.n_past + 1 >= self.n_ctx:
            raise ValueError(f"prompt of 1 tokens does not fit the context ({self.n_ctx})")
        with self.torch.cuda.stream(self.stream):
            self._set_tok_pos(int(tok), self.n_past)
            self._run("prompt_last", self.n_past)
        self.n_past += 1
        self.chain_valid = True

    def read_last_token(self) -> int:
        """Device -> pinned host read of the newest greedy token (what a streaming server does per step)."""
        with self.torch.cuda.stream(self.stream):
            self.host_tok.copy_(self.tok_dev, non_blocking=True)
        self.stream.synchronize()
        return int(self.host_tok_np[0])

    def read_logits(self) -> np.ndarray:
        if self.host_logits is None:
            self.host_logits = self.torch.zeros(self.vl, dtype=self.torch.float32).pin_memory()
        with self.torch.cuda.stream(self.stream):
            self.host_logits.copy_(self.logits, non_blocking=True)
        self.stream.synchronize()
        return self.host_logits.numpy()

    def read_candidates(self, k: int, cap: int = 256, extra_ids=None):
        """(token ids, logits, logits of extra_ids) -- every logit >= the k-th largest (k of them, more on ties; unordered), which is
        what a top-k sampler needs instead of the whole row, plus the raw logits of a few named tokens (a penalty window): one
        ggb_topk_rows launch, one ggb_gather_rows launch when asked, one copy.  None when the vocabulary is sharded (tensor
        parallelism) or ties overflowed the buffer: the caller then reads the row."""
        e, torch = self.eng, self.torch
        m = len(extra_ids) if extra_ids else 0
        if e.tp_size > 1 or not 0 < k <= min(cap, self.vl) or m > cap:
            return None
        if getattr(self, "_cand_dev", None) is None or self._cand_cap != cap:
            self._cand_cap = cap
            self._cand_dev = torch.zeros(4 * cap + 1, dtype=torch.int32, device=e.dev)   # values (f32 bits) | indices | extra values | extra ids | count
            self._cand_host = torch.zeros(4 * cap + 1, dtype=torch.int32).pin_memory()
            self._cand_ids = torch.zeros(cap, dtype=torch.int32).pin_memory()
        d = self._cand_dev
        p0 = d.data_ptr()
        with torch.cuda.stream(self.stream):
            cabi.check(e.lib.ggb_topk_rows(self.logits.data_ptr(), self.vl, 1, k, cap, p0, p0 + 4 * cap, p0 + 16 * cap, self.stream.cuda_stream), "topk_rows")
            if m:
                self._cand_ids.numpy()[:m] = np.asarray(extra_ids, dtype=np.int32)
                d[3 * cap:3 * cap + m].copy_(self._cand_ids[:m], non_blocking=True)
                cabi.check(e.lib.ggb_gather_rows(self.logits.data_ptr(), self.vl, 1, p0 + 12 * cap, m, p0 + 8 * cap, self.stream.cuda_stream), "gather_rows")
            self._cand_host.copy_(d, non_blocking=True)
        self.stream.synchronize()
        h = self._cand_host.numpy()
        c = int(h[4 * cap])
        if c > cap:
            return None
        return h[cap:cap + c].copy(), h[:c].copy().view(np.float32), h[2 * cap:2 * cap + m].copy().view(np.float32)

    def logits_tensor(self):
        """this rank's (vocabulary-sharded under tensor parallelism) logits, on the device, after the stream has drained"""
        self.stream.synchronize()
        return self.logits

    def tokens(self, n: int) -> list[int]:
        self.stream.synchronize()
        return self.out_tokens[:n].cpu().tolist()

    def generate(self, prompt: list[int], n_new: int, stream_cb=None) -> list[int]:
        """Greedy generation.  With stream_cb the newest token is read back after every step (streaming);
        otherwise all steps are enqueued back-to-back and read once."""
        if len(prompt) + n_new >= self.n_ctx:
            raise ValueError("prompt + n_new exceeds the context window")
        self.reset()
        self.prefill(prompt)
        if stream_cb is None:
            self.decode(n_new - 1)
            return self.tokens(n_new)
        out = [self.read_last_token()]
        stream_cb(out[-1])
        for _ in range(n_new - 1):
            self.decode(1)
            out.append(self.read_last_token())
            stream_cb(out[-1])
        return out

    def last_logits(self) -> np.ndarray:
        self.stream.synchronize()
        return self.logits.cpu().numpy()


class Engine:
    """Weights in HBM + n_slots independent sequences.  The single-sequence methods act on slot 0."""

    def __init__(self, path: str, n_ctx: int = 4096, device: int = 0, use_graph: bool = True, use_pdl: bool = True,
                 max_new: int = 65536, verbose: bool = False, n_slots: int = 1, tp_rank: int = 0, tp_size: int = 1,
                 process_group=None):
        import torch

        if not torch.cuda.is_available():
            raise cabi.GGBError("no CUDA device: the GGUF engine has no CPU fallback")
        self.torch = torch
        self.lib = cabi.lib()
        self.dev = torch.device("cuda", device)
        torch.cuda.set_device(self.dev)
        self.use_graph, self.use_pdl = use_graph, int(bool(use_pdl))
        self.file = G.GGUFFile(path)
        self.hp = HParams.from_gguf(self.file)
        self.tp_rank, self.tp_size, self.pg = int(tp_rank), int(tp_size), process_group
        parallel.check_divisible(self.hp, self.tp_size)
        if self.tp_size > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise cabi.GGBError("tensor parallelism needs torch.distributed to be initialised (one process per GPU)")
            self.dist = dist
        self.n_ctx = int(n_ctx)
        self.max_new = max_new
        self.peer = None
        if self.tp_size > 1 and os.environ.get("GGB_TP_EXCHANGE", "peer") != "nccl":
            self._setup_peer_exchange()
        self.stream = torch.cuda.Stream(device=self.dev)
        t0 = time.time()
        self._load_weights()
        self.load_seconds = time.time() - t0
        self.gemm_prefill_min = int(os.environ.get("GGB_GEMM_PREFILL_MIN", "64"))   # prompts at least this long use the GEMM path
        self.prefill_chunk = 2048
        self.prefill_raw_act = os.environ.get("GGB_PREFILL_RAW_ACT", "0") == "1"
        self.prefill_exact_attn = os.environ.get("GGB_PREFILL_EXACT_ATTN", "0") == "1"
        self._pf = None
        n_slots = max(1, n_slots)
        kvd = (self.hp.n_kv // self.tp_size) * self.hp.head_dim
        # one cache pool for all slots, [slot][layer][position][kv dims]: the batched decode kernels address a token's
        # cache as pool + slot * stride (csrc/attn.cu, ggb_rope_kv_batch)
        self.k_all = torch.zeros((n_slots, self.hp.n_layer, self.n_ctx, kvd), dtype=torch.float16, device=self.dev)
        self.v_all = torch.zeros((n_slots, self.hp.n_layer, self.n_ctx, kvd), dtype=torch.float16, device=self.dev)
        self.slots = [Slot(self, i) for i in range(n_slots)]
        self._batch = None
        self.batch_capable = n_slots > 1
        if verbose:
            print(f"[engine] loaded {path}: {self.hp} in {self.load_seconds:.2f}s, weights {self.weight_bytes/1e9:.3f} GB", flush=True)

    def _setup_peer_exchange(self):
        """Map every rank's exchange region into this process (CUDA IPC over NVLink): the all-reduce after the
        row-split projections then happens inside the GEMV epilogue + one small reduce kernel (csrc/peer.cu).
        GGB_TP_EXCHANGE=nccl keeps the library collective instead (the baseline the fused path is measured against)."""
        lib, dist, n = self.lib, self.dist, self.tp_size
        if n > cabi.PEER_MAX:
            raise cabi.GGBError(f"peer exchange supports up to {cabi.PEER_MAX} ranks (GGB_TP_EXCHANGE=nccl for more)")
        nbytes = lib.ggb_peer_region_bytes(n, self.hp.d)
        own, handle = C.c_void_p(), C.create_string_buffer(64)
        cabi.check(lib.ggb_peer_alloc(nbytes, C.byref(own), handle), "peer_alloc")
        handles = [None] * n
        dist.all_gather_object(handles, handle.raw, group=self.pg)
        bases = []
        for r in range(n):
            if r == self.tp_rank:
                bases.append(own.value)
                continue
            p = C.c_void_p()
            cabi.check(lib.ggb_peer_open(handles[r], C.byref(p)), f"peer_open(rank {r}) -- set GGB_TP_EXCHANGE=nccl if CUDA IPC is unavailable")
            bases.append(p.value)
        dist.barrier(group=self.pg)
        self.peer = {"own": own.value, "bases": bases}

    # ------------------------------------------------------------------ loading
    def _sptr(self) -> int:
        return self.torch.cuda.current_stream().cuda_stream

    def _upload(self, name: str):
        """canonical bytes of a tensor on the device (small tensors directly; large ones through the staging pipeline)"""
        torch = self.torch
        src = self.file.data(name)
        st = getattr(self, "_stager", None)
        if st is None or src.size < (4 << 20):
            raw = np.array(src)  # copy: torch refuses read-only mmap views
            return torch.from_numpy(raw).to(self.dev, non_blocking=False)
        with torch.cuda.stream(self.stream):
            dst = torch.empty(src.size, dtype=torch.uint8, device=self.dev)
        for o in range(0, src.size, st.buf_bytes):
            n = min(st.buf_bytes, src.size - o)
            st.submit(src[o:o + n], lambda pinned, o=o, n=n: dst[o:o + n].copy_(pinned, non_blocking=True))
        return dst

    def _load_matrix(self, name: str, shard_as: str | None = None) -> Weight:
        torch = self.torch
        ti = self.file.tensors[name]
        if ti.ggml_type not in QUANT_TYPES + LEGACY_TYPES:
            raise G.GGUFError(f"{name}: tensor type {ti.type_name} is not supported by the GEMV path (supported: Q4_K, Q5_K, Q6_K, Q8_0, Q4_0, Q5_0)")
        k, rows = ti.ne[0], ti.ne[1]
        if k % 256:
            raise G.GGUFError(f"{name}: K={k} is not a multiple of 256")
        sh = parallel.shard_of(shard_as or name, self.hp, self.tp_size, self.tp_rank)
        if ti.ggml_type in LEGACY_TYPES:      # exact conversion to Q8_0 blocks on the host, then the Q8_0 path
            raw = np.asarray(self.file.data(name))
            if sh.kind != parallel.FULL:
                raw, rows, k = parallel.slice_canonical(raw, ti.ggml_type, k, rows, sh)
            q8 = legacy_to_q8_0(raw, ti.ggml_type)
            canon = torch.from_numpy(q8).to(self.dev)
            stride = self.lib.ggb_repacked_row_stride(G.GGML_Q8_0, k)
            dst = torch.zeros(rows * stride + 16, dtype=torch.uint8, device=self.dev)
            cabi.check(self.lib.ggb_repack(G.GGML_Q8_0, canon.data_ptr(), dst.data_ptr(), rows, k, self._sptr()), f"repack {name}")
            torch.cuda.current_stream().synchronize()
            self.weight_bytes += q8.size
            return Weight(dst, G.GGML_Q8_0, rows, k)
        st = getattr(self, "_stager", None)
        if sh.kind == parallel.FULL and st is not None:
            # row chunks: file -> pinned ring (reader threads) -> device staging (async copy) -> tile-SoA destination (ggb_repack
            # on the chunk's rows); nothing waits for anything but a free staging buffer
            rb = G.row_bytes(ti.ggml_type, k)
            stride = self.lib.ggb_repacked_row_stride(ti.ggml_type, k)
            with torch.cuda.stream(self.stream):     # the zero fill must be ordered before the chunks' re-ordering kernels
                dst = torch.zeros(rows * stride + 16, dtype=torch.uint8, device=self.dev)
            src = self.file.data(name)
            per = max(1, st.buf_bytes // rb)
            for r0 in range(0, rows, per):
                nr = min(per, rows - r0)

                def finish(pinned, r0=r0, nr=nr):
                    tmp = st.dev_tmp[:nr * rb]
                    tmp.copy_(pinned, non_blocking=True)
                    cabi.check(self.lib.ggb_repack(ti.ggml_type, tmp.data_ptr(), dst.data_ptr() + r0 * stride, nr, k, self.stream.cuda_stream), f"repack {name}")
                st.submit(src[r0 * rb:(r0 + nr) * rb], finish)
            self.weight_bytes += rows * rb
            return Weight(dst, ti.ggml_type, rows, k)
        if sh.kind == parallel.FULL:
            canon = self._upload(name)
        else:
            part, rows, k = parallel.slice_canonical(np.asarray(self.file.data(name)), ti.ggml_type, k, rows, sh)
            canon = torch.from_numpy(np.array(part)).to(self.dev)
        stride = self.lib.ggb_repacked_row_stride(ti.ggml_type, k)
        # +16: the GEMV's bulk copies round a partial last tile up to 16 bytes
        dst = torch.zeros(rows * stride + 16, dtype=torch.uint8, device=self.dev)
        cabi.check(self.lib.ggb_repack(ti.ggml_type, canon.data_ptr(), dst.data_ptr(), rows, k, self._sptr()), f"repack {name}")
        torch.cuda.current_stream().synchronize()
        self.weight_bytes += canon.numel()
        del canon
        return Weight(dst, ti.ggml_type, rows, k)

    def _load_f32(self, name: str):
        torch = self.torch
        ti = self.file.tensors[name]
        raw = self._upload(name)
        out = torch.empty(ti.n_elements, dtype=torch.float32, device=self.dev)
        cabi.check(self.lib.ggb_dequant(ti.ggml_type, raw.data_ptr(), out.data_ptr(), ti.n_elements, self._sptr()), f"dequant {name}")
        torch.cuda.current_stream().synchronize()
        return out

    def _load_weights(self):
        hp, f = self.hp, self.file
        self.weight_bytes = 0
        self._stager = None
        if os.environ.get("GGB_LOAD_PIPELINE", "1") != "0":
            self._stager = _Stager(self.torch, self.lib, self.stream, self.dev)
        try:
            self._load_tensors()
        finally:
            if self._stager is not None:
                self._stager.close()
                self._stager = None

    def _load_tensors(self):
        hp, f = self.hp, self.file
        emb = f.tensors["token_embd.weight"]
        self.emb_type = emb.ggml_type
        if emb.ggml_type in LEGACY_TYPES:     # same exact conversion: get_rows of the Q8_0 image gives d * (q - 8) / d * (q - 16)
            self.emb_type = G.GGML_Q8_0
            self.emb_canon = self.torch.from_numpy(legacy_to_q8_0(np.asarray(f.data("token_embd.weight")), emb.ggml_type)).to(self.dev)
        else:
            self.emb_canon = self._upload("token_embd.weight")  # canonical layout: get_rows reads one row
        self.layers = []
        for i in range(hp.n_layer):
            p = f"blk.{i}."
            self.layers.append({
                "attn_norm": self._load_f32(p + "attn_norm.weight"),
                "wq": self._load_matrix(p + "attn_q.weight"),
                "wk": self._load_matrix(p + "attn_k.weight"),
                "wv": self._load_matrix(p + "attn_v.weight"),
                "wo": self._load_matrix(p + "attn_output.weight"),
                "ffn_norm": self._load_f32(p + "ffn_norm.weight"),
                "wg": self._load_matrix(p + "ffn_gate.weight"),
                "wu": self._load_matrix(p + "ffn_up.weight"),
                "wd": self._load_matrix(p + "ffn_down.weight"),
            })
        self.out_norm = self._load_f32("output_norm.weight")
        # tied embeddings (no output.weight, e.g. Llama-3.2 1B/3B): the head reads token_embd -- under tensor parallelism
        # vocabulary-sharded like output.weight, while the embedding gather keeps its replicated canonical copy
        tied = "output.weight" not in f.tensors
        self.w_out = self._load_matrix("token_embd.weight" if tied else "output.weight", shard_as="output.weight")
        ff = None
        if "rope_freqs.weight" in f.tensors:
            ff = self._load_f32("rope_freqs.weight").cpu().numpy()
        self.rope_tab = self.torch.from_numpy(rope_table(self.n_ctx, hp.n_rot, hp.rope_base, ff)).to(self.dev)

    def check_tokens(self, tokens) -> None:
        """token ids index the embedding table on the device without a bound check there: refuse anything outside the
        vocabulary here (ValueError = the caller's input is wrong, not an engine failure)"""
        v = self.hp.vocab
        for t in tokens:
            if isinstance(t, bool) or not isinstance(t, (int, np.integer)) or not 0 <= int(t) < v:
                raise ValueError(f"token id {t!r} is outside the vocabulary (0..{v - 1})")

    def prefill_many(self, jobs, head: bool = True):
        """jobs = [(slot index, tokens, start position)]: the prompt chunks of SEVERAL sequences in one pass through the
        tcgen05 dequant-GEMMs (csrc/gemm.cu) -- the GEMMs, norms and element-wise ops run on the concatenated tokens, RoPE
        and the KV write take per-token (slot, position), attention runs per sequence (csrc/prefill.cu).  Each sequence's
        last hidden state then goes through the decode head (arg-max, next embedding), like the GEMV path."""
        torch, lib, hp = self.torch, self.lib, self.hp
        if self.tp_size != 1:
            raise cabi.GGBError("the GEMM prefill path is single-GPU")
        T = sum(len(j[1]) for j in jobs)
        if T == 0 or T > self.prefill_chunk:
            raise ValueError(f"prefill_many: {T} tokens (1..{self.prefill_chunk})")
        for sl, toks, start in jobs:
            if start + len(toks) >= self.n_ctx or not toks:
                raise ValueError("prompt does not fit the context")
            self.check_tokens(toks)
        B = self.prefill_buffers(T)
        s = self.stream.cuda_stream
        qd, kvd = hp.n_head * hp.head_dim, hp.n_kv * hp.head_dim
        slot_stride = hp.n_layer * self.n_ctx * kvd

        def gemm(w, xb, y):
            cabi.check(lib.ggb_gemm(w.type, w.ptr, w.rows, w.k, xb.data_ptr(), T, y.data_ptr(), w.rows, s), "gemm")

        def to_f16(x, k, w):
            """the GEMM's activation operand: x [T][k] quantised like the CPU path quantises it for a weight of w's format
            (Q8_K, or Q8_0 for Q8_0 weights) and dequantised to f16 -- the product then matches ggml's integer dot to f16
            rounding.  GGB_PREFILL_RAW_ACT=1: plain f16 conversion (no activation quantisation), for comparison."""
            if self.prefill_raw_act:
                cabi.check(lib.ggb_f32_to_f16(x.data_ptr(), B["xb"].data_ptr(), T * k, s), "f32_to_f16")
            else:
                cabi.check(lib.ggb_act_fakequant_f16(x.data_ptr(), B["xb"].data_ptr(), k, T, 1 if w.type == G.GGML_Q8_0 else 0, s), "act_fakequant_f16")

        ids, pos, slots = [], [], []
        for sl, toks, start in jobs:
            ids += [int(t) for t in toks]
            pos += list(range(start, start + len(toks)))
            slots += [sl] * len(toks)
        with torch.cuda.stream(self.stream):
            meta = torch.tensor([ids, pos, slots], dtype=torch.int32).to(self.dev)
            X, XN, Y = B["x"], B["xn"], B["y"]
            cabi.check(lib.ggb_embed_rows(self.emb_type, self.emb_canon.data_ptr(), hp.d, meta[0].data_ptr(), T, X.data_ptr(), s), "embed_rows")
            def norm_quant(add, norm_w, w):
                """X += add (the previous projection's output, or None), then the GEMM operand of rms_norm(X) * norm_w for weight w"""
                if self.prefill_raw_act:
                    if add is not None:
                        cabi.check(lib.ggb_add_f32(X.data_ptr(), add.data_ptr(), T * hp.d, s), "add")
                    cabi.check(lib.ggb_rms_norm(X.data_ptr(), norm_w.data_ptr(), XN.data_ptr(), hp.d, T, hp.eps, s), "rms_norm")
                    to_f16(XN, hp.d, w)
                else:
                    cabi.check(lib.ggb_add_rmsnorm_fakequant_f16(X.data_ptr(), add.data_ptr() if add is not None else 0, norm_w.data_ptr(),
                                                                 B["xb"].data_ptr(), hp.d, T, hp.eps, 1 if w.type == G.GGML_Q8_0 else 0, s),
                               "add_rmsnorm_fakequant_f16")

            for i, L in enumerate(self.layers):
                norm_quant(Y if i > 0 else None, L["attn_norm"], L["wq"])      # the previous layer's ffn_down output joins here
                gemm(L["wq"], B["xb"], B["q"]); gemm(L["wk"], B["xb"], B["k"]); gemm(L["wv"], B["xb"], B["v"])
                cabi.check(lib.ggb_rope_kv_batch(B["q"].data_ptr(), B["k"].data_ptr(), B["v"].data_ptr(), T, meta[1].data_ptr(), meta[2].data_ptr(),
                                                 slot_stride, hp.n_head, hp.n_kv, hp.head_dim, hp.n_rot, self.rope_tab.data_ptr(),
                                                 self.k_all[0, i].data_ptr(), self.v_all[0, i].data_ptr(), s), "rope_kv_batch")
                off = 0
                if self.prefill_exact_attn:   # diagnosis: the decode attention (f64 sums), one batch entry per prompt token
                    cabi.check(lib.ggb_attn_decode_batch(B["q"].data_ptr(), self.k_all[0, i].data_ptr(), self.v_all[0, i].data_ptr(),
                                                         meta[1].data_ptr(), meta[2].data_ptr(), slot_stride, T, hp.n_head, hp.n_kv, hp.head_dim,
                                                         self.n_ctx, B["att"].data_ptr(), 0, s), "attn_decode_batch")
                for sl, toks, start in ([] if self.prefill_exact_attn else jobs):
                    cabi.check(lib.ggb_attn_prefill(B["q"].data_ptr() + off * qd * 4, self.k_all[sl, i].data_ptr(), self.v_all[sl, i].data_ptr(),
                                                    len(toks), start, hp.n_head, hp.n_kv, hp.head_dim, B["att"].data_ptr() + off * qd * 4, s), "attn_prefill")
                    off += len(toks)
                to_f16(B["att"], qd, L["wo"])
                gemm(L["wo"], B["xb"], Y)
                norm_quant(Y, L["ffn_norm"], L["wg"])
                wg, wu = L["wg"], L["wu"]
                if wg.type == wu.type and wg.rows == wu.rows and wg.k == wu.k:     # one launch: 6.05 waves instead of 2 x 3.03
                    cabi.check(lib.ggb_gemm2(wg.type, wg.ptr, wu.ptr, wg.rows, wg.k, B["xb"].data_ptr(), T, B["gate"].data_ptr(),
                                             B["up"].data_ptr(), wg.rows, s), "gemm2")
                else:
                    gemm(wg, B["xb"], B["gate"]); gemm(wu, B["xb"], B["up"])
                if self.prefill_raw_act:
                    cabi.check(lib.ggb_swiglu(B["gate"].data_ptr(), B["up"].data_ptr(), B["gate"].data_ptr(), T * hp.ff, s), "swiglu")
                    to_f16(B["gate"], hp.ff, L["wd"])
                else:   # silu(gate) * up, quantised like the CPU path quantises ffn_down's input, as f16: one pass
                    cabi.check(lib.ggb_swiglu_fakequant_f16(B["gate"].data_ptr(), B["up"].data_ptr(), B["xb"].data_ptr(), hp.ff, T,
                                                            1 if L["wd"].type == G.GGML_Q8_0 else 0, s), "swiglu_fakequant_f16")
                gemm(L["wd"], B["xb"], Y)
            cabi.check(lib.ggb_add_f32(X.data_ptr(), Y.data_ptr(), T * hp.d, s), "add")     # the last layer's ffn_down
            off = 0
            for sl, toks, start in jobs:
                slot = self.slots[sl]
                off += len(toks)
                if head:    # (an intermediate chunk of a long prompt only fills the cache)
                    slot.x.copy_(X[(off - 1) * hp.d:off * hp.d], non_blocking=True)
                    slot._set_tok_pos(int(toks[-1]), start + len(toks) - 1)
                    slot._enqueue_head(s)
            self.stream.synchronize()
        for sl, toks, start in jobs:
            self.slots[sl].n_past = start + len(toks)
            self.slots[sl].chain_valid = True

    def embed(self, tokens, slot: int = 0, pooling: str = "mean") -> np.ndarray:
        """Sentence embedding of a prompt (llama-server --embeddings, docs/API_REFERENCE.md:540-590 of the reference): the
        prompt runs through the tensor-core prefill, the final hidden states go through output_norm, are pooled over the
        tokens ("mean", or "last") and L2-normalised (llama.cpp's default embd_normalize = 2).  Fills the slot's K/V for the
        prompt (positions 0..n-1) like any prefill."""
        if self.tp_size != 1:
            raise cabi.GGBError("embeddings are served by the single-GPU prefill path")
        if pooling not in ("mean", "last"):
            raise ValueError(f"unsupported pooling {pooling!r} (mean, last)")
        tokens = [int(t) for t in tokens]
        if not tokens:
            raise ValueError("empty input")
        hp, lib = self.hp, self.lib
        self.slots[slot].reset()
        acc = np.zeros(hp.d, dtype=np.float64)
        last = None
        for c0 in range(0, len(tokens), self.prefill_chunk):
            chunk = tokens[c0:c0 + self.prefill_chunk]
            self.prefill_many([(slot, chunk, c0)], head=False)
            T = len(chunk)
            B = self._pf
            with self.torch.cuda.stream(self.stream):
                cabi.check(lib.ggb_rms_norm(B["x"].data_ptr(), self.out_norm.data_ptr(), B["xn"].data_ptr(), hp.d, T, hp.eps, self.stream.cuda_stream), "rms_norm")
                h = B["xn"][:T * hp.d].view(T, hp.d)
                h = h[-1:] if pooling == "last" else h
                host = h.cpu().numpy()
            if pooling == "last":
                last = host[0].astype(np.float64)
            else:
                acc += host.astype(np.float64).sum(axis=0)
        v = last if pooling == "last" else acc / len(tokens)
        n = float(np.linalg.norm(v))
        return (v / n if n > 0 else v).astype(np.float32)

    def prefill_buffers(self, T: int) -> dict:
        """activation buffers of the GEMM prefill path, sized for the largest chunk seen so far (shared by the slots)"""
        cap = max(T, 64)
        if self._pf is None or self._pf["cap"] < cap:
            torch, hp, dev = self.torch, self.hp, self.dev
            f32 = lambda n: torch.empty(cap * n, dtype=torch.float32, device=dev)  # noqa: E731
            qd, kvd = hp.n_head * hp.head_dim, hp.n_kv * hp.head_dim
            self._pf = {"cap": cap, "x": f32(hp.d), "xn": f32(hp.d), "y": f32(hp.d), "q": f32(qd), "k": f32(kvd), "v": f32(kvd),
                        "att": f32(qd), "gate": f32(hp.ff), "up": f32(hp.ff),
                        "xb": torch.empty(cap * max(hp.d, hp.ff, qd), dtype=torch.float16, device=dev)}
        return self._pf

    @property
    def batch(self):
        """Batched decode over the slots (batch.py); built on first use."""
        if self._batch is None:
            from .batch import BatchDecoder
            self._batch = BatchDecoder(self)
        return self._batch

    def launches_per_step(self) -> int:
        """kernels of libggufb200 launched by one decode step (layers + head); NCCL kernels are not counted."""
        return self.hp.n_layer * (7 if self.tp_size > 1 else 5) + (4 if self.tp_size > 1 else 3)

    # ------------------------------------------------------------------ single-sequence convenience (slot 0)
    @property
    def _s0(self) -> Slot:
        return self.slots[0]

    def warmup(self):
        for s in self.slots:
            s.warmup()

    def reset(self):
        self._s0.reset()

    def prefill(self, tokens, start_pos=None):
        self._s0.prefill(tokens, start_pos)

    def decode(self, n_steps: int):
        self._s0.decode(n_steps)

    def read_last_token(self) -> int:
        return self._s0.read_last_token()

    def tokens(self, n: int):
        return self._s0.tokens(n)

    def generate(self, prompt, n_new, stream_cb=None):
        return self._s0.generate(prompt, n_new, stream_cb)

    def last_logits(self):
        return self._s0.last_logits()

    @property
    def _layer_args(self):
        return self._s0._layer_args

    @property
    def _head(self):
        return self._s0._head

    def _set_tok_pos(self, tok, pos):
        self._s0._set_tok_pos(tok, pos)

    def close(self):
        for s in self.slots:
            s._graphs.clear()
        if self._batch is not None:
            self._batch._graphs.clear()
        if self.peer:
            self.torch.cuda.synchronize()
            for r, b in enumerate(self.peer["bases"]):
                if r != self.tp_rank:
                    self.lib.ggb_peer_close(b)
            self.lib.ggb_peer_free(self.peer["own"])
            self.peer = None
        self.file.close()
