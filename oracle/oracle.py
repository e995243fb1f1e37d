"""CPU oracle for the GGUF decode path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package never does.

What it restates: the arithmetic of the reference's backend run with NGL=0 (`/app/llama-server`,
/root/reference/scripts/start.sh:473-480,516; CPU image /root/reference/Dockerfile.cpu:11,84-89), i.e.
ggml's CPU path for a llama-architecture model: Q8_K / Q8_0 activation quantisation, integer vec_dot
against Q4_K / Q5_K / Q6_K / Q8_0 / Q4_0 / Q5_0 weights, RMSNorm, NORM-mode RoPE, f16 KV cache, f32 softmax, SwiGLU,
greedy argmax.  The C kernels live in ggml_ref.c (see its header for the parity-pinning status:
dequantisation and Q8_0 quantisation are pinned bit-exact against gguf-py; the rest is PARITY UNPINNED
because neither the reference nor this machine holds upstream's sources or any numeric test vector).

The GGUF container is read with upstream's own `gguf.GGUFReader` (gguf-py 0.19.0), deliberately NOT with
the product's reader, so that the two readers check each other.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libggml_ref.so")

F32, F16, Q4_0, Q5_0, Q8_0, Q4_K, Q5_K, Q6_K = 0, 1, 2, 6, 8, 12, 13, 14
BLOCK = {F32: (1, 4), F16: (1, 2), Q4_0: (32, 18), Q5_0: (32, 22), Q8_0: (32, 34), Q4_K: (256, 144), Q5_K: (256, 176), Q6_K: (256, 210)}


def build(force: bool = False) -> str:
    """Compile ggml_ref.c (gcc, OpenMP).  Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "ggml_ref.c")):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float
        L.gref_dequantize_row.argtypes = [i32, vp, vp, i64]; L.gref_dequantize_row.restype = i32
        L.gref_quantize_row_q8_K.argtypes = [vp, vp, i64]; L.gref_quantize_row_q8_K.restype = None
        L.gref_quantize_row_q8_0.argtypes = [vp, vp, i64]; L.gref_quantize_row_q8_0.restype = None
        for n in ("q4_K_q8_K", "q5_K_q8_K", "q6_K_q8_K", "q8_0_q8_0"):
            f = getattr(L, "gref_vec_dot_" + n); f.argtypes = [i64, vp, vp]; f.restype = f32
        L.gref_matmul.argtypes = [i32, vp, i64, i64, vp, i64, vp, i32]; L.gref_matmul.restype = i32
        L.gref_rms_norm.argtypes = [vp, vp, vp, i64, f32]; L.gref_rms_norm.restype = None
        L.gref_rope_norm.argtypes = [vp, i32, i32, i32, i32, f32, vp]; L.gref_rope_norm.restype = None
        L.gref_rope_table.argtypes = [i32, i32, f32, vp, vp]; L.gref_rope_table.restype = None
        L.gref_swiglu.argtypes = [vp, vp, vp, i64]; L.gref_swiglu.restype = None
        L.gref_attn_decode.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i64]; L.gref_attn_decode.restype = None
        L.gref_argmax.argtypes = [vp, i64]; L.gref_argmax.restype = i64
        L.gref_fp32_to_fp16_row.argtypes = [vp, vp, i64]; L.gref_fp32_to_fp16_row.restype = None
        L.gref_fp16_to_fp32_row.argtypes = [vp, vp, i64]; L.gref_fp16_to_fp32_row.restype = None
        L.gref_num_threads.argtypes = []; L.gref_num_threads.restype = i32
        # order-independent ("canon") variants, see ggml_ref.c
        L.gref_matmul_mode.argtypes = [i32, vp, i64, i64, vp, i64, vp, i32, i32]; L.gref_matmul_mode.restype = i32
        L.gref_exp_ref.argtypes = [f32]; L.gref_exp_ref.restype = f32
        L.gref_matvec_f64.argtypes = [i32, vp, i64, i64, vp, vp, i32]; L.gref_matvec_f64.restype = i32
        L.gref_swiglu_canon.argtypes = [vp, vp, vp, i64]; L.gref_swiglu_canon.restype = None
        L.gref_rope_table_canon.argtypes = [i32, i32, f32, vp, vp]; L.gref_rope_table_canon.restype = None
        L.gref_rope_apply.argtypes = [vp, i32, i32, i32, vp]; L.gref_rope_apply.restype = None
        L.gref_attn_decode_canon.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i64]; L.gref_attn_decode_canon.restype = None
        _lib = L
    return _lib


def _p(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def row_bytes(qtype: int, k: int) -> int:
    be, bb = BLOCK[qtype]
    assert k % be == 0
    return k // be * bb


# ----------------------------------------------------------------------------- primitives
def dequantize(raw: np.ndarray, qtype: int, n: int) -> np.ndarray:
    """raw uint8 bytes of n elements -> float32[n]  (ggml dequantize_row_*)."""
    raw = np.ascontiguousarray(raw, dtype=np.uint8).reshape(-1)
    assert raw.size == row_bytes(qtype, n), (raw.size, qtype, n)
    out = np.empty(n, dtype=np.float32)
    if n == 0:
        return out
    rc = lib().gref_dequantize_row(qtype, _p(raw), _p(out), n)
    assert rc == 0
    return out


def quantize_q8_K(x: np.ndarray) -> np.ndarray:
    """float32[k] -> packed block_q8_K bytes (292 B per 256)."""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    out = np.zeros(x.size // 256 * 292, dtype=np.uint8)
    lib().gref_quantize_row_q8_K(_p(x), _p(out), x.size)
    return out


def q8_K_fields(packed: np.ndarray):
    b = packed.reshape(-1, 292)
    d = b[:, :4].copy().view(np.float32).reshape(-1)
    qs = b[:, 4:260].copy().view(np.int8)
    bsums = b[:, 260:].copy().view(np.int16)
    return d, qs, bsums


def quantize_q8_0(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    out = np.zeros(x.size // 32 * 34, dtype=np.uint8)
    lib().gref_quantize_row_q8_0(_p(x), _p(out), x.size)
    return out


GGML, CANON = "ggml", "canon"  # f32 accumulation in ggml's generic order | order-independent f64 accumulation


def matmul(qtype: int, w_raw: np.ndarray, rows: int, k: int, x: np.ndarray, nthreads: int = 0, mode: str = GGML) -> np.ndarray:
    """Y[m, rows] = W[rows, k] . X[m, k] through activation quantisation + integer vec_dot (CPU mul_mat)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    m = 1 if x.ndim == 1 else x.shape[0]
    assert x.size == m * k
    w_raw = np.ascontiguousarray(w_raw, dtype=np.uint8).reshape(-1)
    assert w_raw.size == rows * row_bytes(qtype, k)
    y = np.empty((m, rows), dtype=np.float32)
    rc = lib().gref_matmul_mode(qtype, _p(w_raw), rows, k, _p(x), m, _p(y), nthreads, 1 if mode == CANON else 0)
    assert rc == 0
    return y[0] if x.ndim == 1 else y


def matvec_f64(qtype: int, w_raw: np.ndarray, rows: int, k: int, x: np.ndarray, nthreads: int = 0) -> np.ndarray:
    """unrounded canon row sums (f64) of W.x -- what tensor-parallel ranks exchange before the single rounding"""
    x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1)
    w_raw = np.ascontiguousarray(w_raw, dtype=np.uint8).reshape(-1)
    assert x.size == k and w_raw.size == rows * row_bytes(qtype, k)
    y = np.empty(rows, dtype=np.float64)
    rc = lib().gref_matvec_f64(qtype, _p(w_raw), rows, k, _p(x), _p(y), nthreads)
    assert rc == 0
    return y


def rms_norm(x: np.ndarray, w: np.ndarray | None, eps: float) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.empty_like(x)
    wp = _p(np.ascontiguousarray(w, dtype=np.float32)) if w is not None else None
    lib().gref_rms_norm(_p(x), wp, _p(y), x.size, eps)
    return y


def rope_norm(x: np.ndarray, n_heads: int, head_dim: int, n_rot: int, pos: int, freq_base: float,
              freq_factors: np.ndarray | None = None) -> np.ndarray:
    y = np.ascontiguousarray(x, dtype=np.float32).copy()
    ff = _p(np.ascontiguousarray(freq_factors, dtype=np.float32)) if freq_factors is not None else None
    lib().gref_rope_norm(_p(y), n_heads, head_dim, n_rot, pos, freq_base, ff)
    return y


def rope_table(pos: int, n_rot: int, freq_base: float, freq_factors: np.ndarray | None = None) -> np.ndarray:
    out = np.empty((n_rot // 2, 2), dtype=np.float32)
    ff = _p(np.ascontiguousarray(freq_factors, dtype=np.float32)) if freq_factors is not None else None
    lib().gref_rope_table(pos, n_rot, freq_base, ff, _p(out))
    return out


def swiglu(g: np.ndarray, u: np.ndarray, mode: str = GGML) -> np.ndarray:
    g = np.ascontiguousarray(g, dtype=np.float32); u = np.ascontiguousarray(u, dtype=np.float32)
    out = np.empty_like(g)
    (lib().gref_swiglu_canon if mode == CANON else lib().gref_swiglu)(_p(g), _p(u), _p(out), g.size)
    return out


def exp_ref(x: float) -> float:
    return float(lib().gref_exp_ref(float(np.float32(x))))


def rope_table_canon(pos: int, n_rot: int, freq_base: float, freq_factors: np.ndarray | None = None) -> np.ndarray:
    out = np.empty((n_rot // 2, 2), dtype=np.float32)
    ff = _p(np.ascontiguousarray(freq_factors, dtype=np.float32)) if freq_factors is not None else None
    lib().gref_rope_table_canon(pos, n_rot, freq_base, ff, _p(out))
    return out


def rope_apply(x: np.ndarray, n_heads: int, head_dim: int, n_rot: int, tab: np.ndarray) -> np.ndarray:
    y = np.ascontiguousarray(x, dtype=np.float32).copy()
    tab = np.ascontiguousarray(tab, dtype=np.float32)
    lib().gref_rope_apply(_p(y), n_heads, head_dim, n_rot, _p(tab))
    return y


def fp32_to_fp16(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, dtype=np.uint16)
    lib().gref_fp32_to_fp16_row(_p(x), _p(out), x.size)
    return out


def attn_decode(q: np.ndarray, kc: np.ndarray, vc: np.ndarray, n_head: int, n_kv: int, hd: int, n_pos: int,
                mode: str = GGML) -> np.ndarray:
    """q f32 [n_head*hd]; kc, vc uint16 (f16 bits) [ctx, n_kv*hd]; attends positions [0, n_pos)."""
    q = np.ascontiguousarray(q, dtype=np.float32)
    assert kc.dtype == np.uint16 and vc.dtype == np.uint16 and kc.flags.c_contiguous and vc.flags.c_contiguous
    out = np.empty(n_head * hd, dtype=np.float32)
    fn = lib().gref_attn_decode_canon if mode == CANON else lib().gref_attn_decode
    fn(_p(q), _p(kc), _p(vc), _p(out), n_head, n_kv, hd, n_pos, n_kv * hd)
    return out


def argmax(x: np.ndarray) -> int:
    x = np.ascontiguousarray(x, dtype=np.float32)
    return int(lib().gref_argmax(_p(x), x.size))


# ----------------------------------------------------------------------------- llama forward
class OracleLlama:
    """Token-by-token llama forward over a GGUF file, graph order of upstream's llm_build_llama
    [UPSTREAM-MEM: src/llama-model.cpp]:

        x = get_rows(token_embd, tok)
        per layer: h = rms_norm(x)*attn_norm ; q,k,v = W h ; rope(q), rope(k) ; cache k,v as f16 ;
                   a = softmax(q k^T / sqrt(hd)) v ; x += Wo a ;
                   h = rms_norm(x)*ffn_norm ; x += Wdown( silu(Wgate h) * (Wup h) )
        logits = Woutput ( rms_norm(x)*output_norm )        (output.weight falls back to token_embd)
    """

    def __init__(self, path: str, n_ctx: int = 512, nthreads: int = 0, mode: str = GGML):
        """mode="ggml": f32 accumulation in ggml's generic order, libm expf/cosf (the restatement of the
        reference's CPU path).  mode="canon": identical integers and per-unit f32 terms, f64 accumulation,
        gref_exp_ref, f64-rounded rope table -- the order-independent definition the CUDA kernels implement
        bit-for-bit (see ggml_ref.c)."""
        assert mode in (GGML, CANON)
        self.mode = mode
        from gguf import GGUFReader  # upstream's reader, on purpose (see module docstring)

        rd = GGUFReader(path)
        self._rd = rd

        def kv(key, default=None):
            f = rd.fields.get(key)
            if f is None:
                return default
            v = f.contents()
            return v

        arch = kv("general.architecture")
        assert arch == "llama", arch
        self.n_layer = int(kv("llama.block_count"))
        self.d = int(kv("llama.embedding_length"))
        self.ff = int(kv("llama.feed_forward_length"))
        self.n_head = int(kv("llama.attention.head_count"))
        self.n_kv = int(kv("llama.attention.head_count_kv", self.n_head))
        self.eps = float(kv("llama.attention.layer_norm_rms_epsilon", 1e-5))
        self.hd = int(kv("llama.attention.key_length", self.d // self.n_head))
        self.n_rot = int(kv("llama.rope.dimension_count", self.hd))
        self.freq_base = float(kv("llama.rope.freq_base", 10000.0))
        self.n_ctx = n_ctx
        self.nthreads = nthreads
        self.t = {}
        for t in rd.tensors:
            shape = [int(s) for s in t.shape]  # ne order: ne0 (innermost) first
            self.t[t.name] = (int(t.tensor_type), shape, np.asarray(t.data).reshape(-1).view(np.uint8))
        self.vocab = self.t["token_embd.weight"][1][1]
        self._norm_cache = {}
        self.freq_factors = None
        if "rope_freqs.weight" in self.t:
            self.freq_factors = self._f32("rope_freqs.weight")
        self.reset()

    def _f32(self, name):
        qt, shape, raw = self.t[name]
        return dequantize(raw, qt, int(np.prod(shape)))

    def _norm(self, name):
        if name not in self._norm_cache:
            self._norm_cache[name] = self._f32(name)
        return self._norm_cache[name]

    def reset(self):
        kvd = self.n_kv * self.hd
        self.kc = np.zeros((self.n_layer, self.n_ctx, kvd), dtype=np.uint16)
        self.vc = np.zeros((self.n_layer, self.n_ctx, kvd), dtype=np.uint16)

    def _mm(self, name, x):
        qt, shape, raw = self.t[name]
        k, rows = shape[0], shape[1]
        return matmul(qt, raw, rows, k, x, self.nthreads, self.mode)

    def embed(self, tok: int) -> np.ndarray:
        qt, shape, raw = self.t["token_embd.weight"]
        rb = row_bytes(qt, shape[0])
        return dequantize(raw[tok * rb:(tok + 1) * rb], qt, shape[0])

    def layer(self, l: int, x: np.ndarray, pos: int) -> np.ndarray:
        """One transformer block on the residual stream x at position pos (writes the KV cache)."""
        p = f"blk.{l}."
        h = rms_norm(x, self._norm(p + "attn_norm.weight"), self.eps)
        q = self._mm(p + "attn_q.weight", h)
        k = self._mm(p + "attn_k.weight", h)
        v = self._mm(p + "attn_v.weight", h)
        if self.mode == CANON:
            tab = rope_table_canon(pos, self.n_rot, self.freq_base, self.freq_factors)
            q = rope_apply(q, self.n_head, self.hd, self.n_rot, tab)
            k = rope_apply(k, self.n_kv, self.hd, self.n_rot, tab)
        else:
            q = rope_norm(q, self.n_head, self.hd, self.n_rot, pos, self.freq_base, self.freq_factors)
            k = rope_norm(k, self.n_kv, self.hd, self.n_rot, pos, self.freq_base, self.freq_factors)
        self.kc[l, pos] = fp32_to_fp16(k)
        self.vc[l, pos] = fp32_to_fp16(v)
        a = attn_decode(q, self.kc[l], self.vc[l], self.n_head, self.n_kv, self.hd, pos + 1, self.mode)
        x = x + self._mm(p + "attn_output.weight", a)
        h = rms_norm(x, self._norm(p + "ffn_norm.weight"), self.eps)
        g = self._mm(p + "ffn_gate.weight", h)
        u = self._mm(p + "ffn_up.weight", h)
        return x + self._mm(p + "ffn_down.weight", swiglu(g, u, self.mode))

    def head(self, x: np.ndarray, return_hidden: bool = False) -> np.ndarray:
        h = rms_norm(x, self._norm("output_norm.weight"), self.eps)
        if return_hidden:
            return h
        out_name = "output.weight" if "output.weight" in self.t else "token_embd.weight"
        return self._mm(out_name, h)

    def forward(self, tok: int, pos: int, return_hidden: bool = False) -> np.ndarray:
        assert pos < self.n_ctx
        x = self.embed(tok)
        for l in range(self.n_layer):
            x = self.layer(l, x, pos)
        return self.head(x, return_hidden)

    def greedy(self, prompt: list[int], n_new: int, return_logits: bool = False):
        """Feed the prompt token by token, then generate n_new tokens by argmax."""
        self.reset()
        logits = None
        for i, t in enumerate(prompt):
            logits = self.forward(t, i)
        out, all_logits = [], []
        pos = len(prompt)
        for _ in range(n_new):
            nxt = argmax(logits)
            out.append(nxt)
            if return_logits:
                all_logits.append(logits)
            if len(out) == n_new:
                break
            logits = self.forward(nxt, pos)
            pos += 1
        return (out, all_logits) if return_logits else out
