"""Seeded synthetic llama-architecture GGUF files ("randomly initialised synthetic GGUF", BASELINE.json
configs; recipe in SURVEY.md section 8d).

There is no network for checkpoints, so every parity test and benchmark runs on files made here.  Weights
are random *valid packed blocks* (uniform nibbles / 6-bit fields / 6-bit sub-scales) whose f16 super-block
scales are set per tensor so the dequantised standard deviation hits a target (1/sqrt(K) for projections,
0.02 for embeddings and the lm-head).  The tensor-type mix for "Q4_K_M" mirrors upstream's
llama_tensor_get_type [UPSTREAM-MEM: src/llama-quant.cpp]: everything Q4_K except output.weight -> Q6_K and
attn_v / ffn_down -> Q6_K in layers where  i < L/8  or  i >= 7L/8  or  (i - L/8) % 3 == 2.

The model file that the reference hands to its backend is `$DATA_DIR/models/$MODEL_NAME`
(/root/reference/scripts/start.sh:309-343); files written here are drop-in replacements for it.
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

import numpy as np

from . import gguf_reader as G

Q4_K_M, Q8_0, Q6_K, Q5_K_M = "Q4_K_M", "Q8_0", "Q6_K", "Q5_K_M"
_FILE_TYPE_ID = {Q4_K_M: 15, Q8_0: 7, Q6_K: 18, Q5_K_M: 17, "Q4_0": 2, "Q5_0": 8}  # general.file_type, gguf/constants.py:4107-4125


@dataclass(frozen=True)
class LlamaConfig:
    name: str
    n_layer: int
    d: int
    n_head: int
    n_kv: int
    head_dim: int
    ff: int
    vocab: int
    rope_base: float = 10000.0
    eps: float = 1e-5
    ctx_train: int = 8192
    resid_scale: float = 1.0   # factor on the std of attn_output / ffn_down (the projections that write the residual stream)

    @property
    def n_params_matmul(self) -> int:
        per_layer = self.d * self.n_head * self.head_dim * 2 + 2 * self.d * self.n_kv * self.head_dim + 3 * self.d * self.ff
        return self.n_layer * per_layer + self.vocab * self.d


PRESETS = {
    # test-sized models; every K is a multiple of 256 (K-quant super-block)
    "tiny": LlamaConfig("tiny", 2, 256, 4, 2, 64, 512, 512),
    "small": LlamaConfig("small", 4, 512, 8, 2, 64, 1536, 2048),
    "medium": LlamaConfig("medium", 8, 1024, 16, 4, 64, 2816, 8192),
    # BASELINE.json configs (dims: SURVEY.md section 8)
    "tinyllama-1.1b": LlamaConfig("tinyllama-1.1b", 22, 2048, 32, 4, 64, 5632, 32000, 10000.0, 1e-5, 2048),
    "llama3-8b": LlamaConfig("llama3-8b", 32, 4096, 32, 8, 128, 14336, 128256, 500000.0, 1e-5, 8192),
    "llama3-70b": LlamaConfig("llama3-70b", 80, 8192, 64, 8, 128, 28672, 128256, 500000.0, 1e-5, 8192),
}


def damped(cfg: LlamaConfig | str, factor: float = 0.003) -> LlamaConfig:
    """The same architecture with the residual-writing projections (attn_output, ffn_down) scaled by factor / sqrt(2L).

    Why: activations are quantised to int8 (Q8_K / Q8_0) before every matmul.  Two implementations whose f32 sums differ
    in the last bit (ggml's generic / AVX2 / NEON kernels, or this engine's order-independent f64 sums) occasionally put
    one activation on the other side of a rounding boundary; with heavy-tailed SwiGLU outputs one flipped code moves a
    layer output by 1e-4..1e-3, after which a percent of ALL later codes flip and the two runs decorrelate down to the
    int8 quantisation noise itself (1-4e-2 on logits, measured between two orderings of the CPU oracle; tools/
    diag_parity.py).  No initialisation of a random net removes that cascade -- GPT-2's 1/sqrt(2L) leaves it unchanged --
    but shrinking every layer's contribution bounds what it can do to the logits.  On this family the CUDA path is held
    to BASELINE.json's bar against the ggml-order oracle itself: logits within 1e-2 at every step, 64 identical tokens
    (tests/test_gpu_engine.py); the plain presets are checked bit-for-bit against the order-independent oracle."""
    from dataclasses import replace
    if isinstance(cfg, str):
        cfg = PRESETS[cfg]
    return replace(cfg, name=f"{cfg.name}-damped", resid_scale=factor * (2.0 * cfg.n_layer) ** -0.5)


def tensor_plan(cfg: LlamaConfig, ftype: str):
    """[(name, ne (ggml order), ggml_type, target_std | None)] in file order."""
    L = cfg.n_layer

    def mixed(i: int) -> bool:  # use_more_bits
        return i < L // 8 or i >= 7 * L // 8 or (i - L // 8) % 3 == 2

    if ftype == Q4_K_M:
        base, big, out = G.GGML_Q4_K, G.GGML_Q6_K, G.GGML_Q6_K
    elif ftype == Q5_K_M:
        base, big, out = G.GGML_Q5_K, G.GGML_Q6_K, G.GGML_Q6_K
    elif ftype == Q8_0:
        base = big = out = G.GGML_Q8_0
    elif ftype == Q6_K:
        base = big = out = G.GGML_Q6_K
    elif ftype in ("Q4_0", "Q5_0"):     # llama.cpp's legacy mixes: every matrix in the 32-element format, the lm-head in Q6_K
        base = big = G.GGML_Q4_0 if ftype == "Q4_0" else G.GGML_Q5_0
        out = G.GGML_Q6_K
    else:
        raise ValueError(f"unknown synthetic file type {ftype}")
    qd, kvd = cfg.n_head * cfg.head_dim, cfg.n_kv * cfg.head_dim
    plan = [("token_embd.weight", (cfg.d, cfg.vocab), base, 0.02)]
    for i in range(L):
        p = f"blk.{i}."
        hi = big if mixed(i) else base
        plan += [
            (p + "attn_norm.weight", (cfg.d,), G.GGML_F32, None),
            (p + "attn_q.weight", (cfg.d, qd), base, cfg.d ** -0.5),
            (p + "attn_k.weight", (cfg.d, kvd), base, cfg.d ** -0.5),
            (p + "attn_v.weight", (cfg.d, kvd), hi, cfg.d ** -0.5),
            (p + "attn_output.weight", (qd, cfg.d), base, qd ** -0.5 * cfg.resid_scale),
            (p + "ffn_norm.weight", (cfg.d,), G.GGML_F32, None),
            (p + "ffn_gate.weight", (cfg.d, cfg.ff), base, cfg.d ** -0.5),
            (p + "ffn_up.weight", (cfg.d, cfg.ff), base, cfg.d ** -0.5),
            (p + "ffn_down.weight", (cfg.ff, cfg.d), hi, cfg.ff ** -0.5 * cfg.resid_scale),
        ]
    plan += [("output_norm.weight", (cfg.d,), G.GGML_F32, None),
             ("output.weight", (cfg.d, cfg.vocab), out, 0.02)]
    return plan


def weight_bytes_per_token(cfg: LlamaConfig, ftype: str) -> dict:
    """Algorithmic HBM bytes of one decoded token in canonical GGUF block bytes: every matmul weight read
    once (token_embd contributes one row).  Matches the table in SURVEY.md section 8d."""
    by_type: dict[str, int] = {}
    norms = 0
    for name, ne, tt, _ in tensor_plan(cfg, ftype):
        if tt == G.GGML_F32:
            norms += ne[0] * 4
            continue
        if name == "token_embd.weight":
            continue
        b = ne[1] * G.row_bytes(tt, ne[0])
        by_type[G.type_name(tt)] = by_type.get(G.type_name(tt), 0) + b
    emb_t = tensor_plan(cfg, ftype)[0][2]
    return {"weights": sum(by_type.values()), "by_type": by_type, "norms": norms,
            "embed_row": G.row_bytes(emb_t, cfg.d),
            "kv_per_pos": 2 * 2 * cfg.n_layer * cfg.n_kv * cfg.head_dim}


N_SPECIAL = 3 + 256  # <unk> <s> </s> + byte-fallback tokens


# ----------------------------------------------------------------------------- random packed blocks
_UNIT_STD = {}  # dequantised std of a block whose f16 scale(s) are 1.0 (closed forms, uniform random fields)


def _unit_std(tt: int) -> float:
    if tt not in _UNIT_STD:
        if tt == G.GGML_Q8_0:      # d*q, q uniform int8
            v = np.mean(np.arange(-128, 128, dtype=np.float64) ** 2)
        elif tt in (G.GGML_Q4_0, G.GGML_Q5_0):   # d*(q - 8) / d*(q - 16), q uniform
            h = 8 if tt == G.GGML_Q4_0 else 16
            v = np.mean(np.arange(-h, h, dtype=np.float64) ** 2)
        elif tt in (G.GGML_Q4_K, G.GGML_Q5_K):  # d*(sc*q - c*m), dmin = c*d with c = E[q] so the mean is zero
            qmax = 15 if tt == G.GGML_Q4_K else 31
            q = np.arange(0, qmax + 1, dtype=np.float64)
            s = np.arange(0, 64, dtype=np.float64)
            c = q.mean()
            v = (s ** 2).mean() * (q ** 2).mean() - 2 * c * s.mean() * q.mean() * s.mean() + c * c * (s ** 2).mean()
        elif tt == G.GGML_Q6_K:    # d*sc*(q-32), sc uniform int8, q uniform 0..63
            v = np.mean(np.arange(-128, 128, dtype=np.float64) ** 2) * np.mean(np.arange(-32, 32, dtype=np.float64) ** 2)
        else:
            raise ValueError(tt)
        _UNIT_STD[tt] = float(np.sqrt(v))
    return _UNIT_STD[tt]


def _f16_bytes(v: np.ndarray) -> np.ndarray:
    return v.astype(np.float16).view(np.uint8).reshape(-1, 2)


def random_blocks(tt: int, n_blocks: int, std: float, rng: np.random.Generator) -> np.ndarray:
    """uint8 [n_blocks, block_bytes]: uniform random packed fields, f16 scales = per-block jitter around the
    value that gives dequantised std `std`."""
    _, _, bb = G.GGML_TYPES[tt]
    n64 = (n_blocks * bb + 7) // 8
    raw = rng.bit_generator.random_raw(n64).view(np.uint8)[: n_blocks * bb].reshape(n_blocks, bb)
    d = (std / _unit_std(tt)) * rng.uniform(0.5, 1.5, size=n_blocks)
    if tt in (G.GGML_Q8_0, G.GGML_Q4_0, G.GGML_Q5_0):
        raw[:, 0:2] = _f16_bytes(d)
    elif tt in (G.GGML_Q4_K, G.GGML_Q5_K):
        raw[:, 0:2] = _f16_bytes(d)
        raw[:, 2:4] = _f16_bytes(d * (7.5 if tt == G.GGML_Q4_K else 15.5))
    elif tt == G.GGML_Q6_K:
        raw[:, 208:210] = _f16_bytes(d)
    return raw


def random_tensor(name: str, ne: tuple, tt: int, std, rng: np.random.Generator) -> np.ndarray:
    if tt == G.GGML_F32:  # norm gains around 1
        return (1.0 + 0.1 * rng.standard_normal(ne[0])).astype(np.float32)
    n = 1
    for x in ne:
        n *= x
    raw = random_blocks(tt, n // G.GGML_TYPES[tt][1], std, rng)
    if name == "output.weight":
        # lm-head rows of the control and byte-fallback tokens get a zero scale: their logit is exactly 0, below
        # the best of the thousands of random word logits, so greedy text is one leading-space word per token
        blocks_per_row = ne[0] // G.GGML_TYPES[tt][1]
        sp = raw[: N_SPECIAL * blocks_per_row]
        if tt in (G.GGML_Q8_0, G.GGML_Q4_0, G.GGML_Q5_0):
            sp[:, 0:2] = 0
        elif tt in (G.GGML_Q4_K, G.GGML_Q5_K):
            sp[:, 0:4] = 0
        elif tt == G.GGML_Q6_K:
            sp[:, 208:210] = 0
    return raw


# ----------------------------------------------------------------------------- vocabulary


def synthetic_vocab(vocab: int):
    """SPM-style vocabulary: 3 control tokens, 256 byte tokens, then one distinct leading-space word per id, so
    whitespace-splitting the generated text (what /root/reference/scripts/benchmark.py:120-125 counts) equals
    the true token count."""
    assert vocab > N_SPECIAL
    toks = ["<unk>", "<s>", "</s>"] + [f"<0x{b:02X}>" for b in range(256)]
    types = [2, 3, 3] + [6] * 256
    toks += [f"▁w{i}" for i in range(N_SPECIAL, vocab)]
    types += [1] * (vocab - N_SPECIAL)
    scores = [0.0] * N_SPECIAL + [-float(i) for i in range(vocab - N_SPECIAL)]
    return toks, scores, types


def write_gguf(path: str, cfg: LlamaConfig | str, ftype: str = Q4_K_M, seed: int = 0xB200) -> dict:
    """Write the synthetic model; returns a small manifest (config, file type, per-token bytes)."""
    if isinstance(cfg, str):
        cfg = PRESETS[cfg]
    w = G.GGUFWriter()
    w.add("general.architecture", G.T_STR, "llama")
    w.add("general.name", G.T_STR, f"synthetic-{cfg.name}-{ftype}")
    w.add("general.file_type", G.T_U32, _FILE_TYPE_ID[ftype])
    w.add("general.quantization_version", G.T_U32, 2)
    w.add("llama.block_count", G.T_U32, cfg.n_layer)
    w.add("llama.context_length", G.T_U32, cfg.ctx_train)
    w.add("llama.embedding_length", G.T_U32, cfg.d)
    w.add("llama.feed_forward_length", G.T_U32, cfg.ff)
    w.add("llama.attention.head_count", G.T_U32, cfg.n_head)
    w.add("llama.attention.head_count_kv", G.T_U32, cfg.n_kv)
    w.add("llama.attention.layer_norm_rms_epsilon", G.T_F32, cfg.eps)
    w.add("llama.rope.dimension_count", G.T_U32, cfg.head_dim)
    w.add("llama.rope.freq_base", G.T_F32, cfg.rope_base)
    w.add("llama.vocab_size", G.T_U32, cfg.vocab)
    if cfg.head_dim * cfg.n_head != cfg.d:
        w.add("llama.attention.key_length", G.T_U32, cfg.head_dim)
        w.add("llama.attention.value_length", G.T_U32, cfg.head_dim)
    toks, scores, types = synthetic_vocab(cfg.vocab)
    w.add("tokenizer.ggml.model", G.T_STR, "llama")
    w.add_array("tokenizer.ggml.tokens", G.T_STR, toks)
    w.add_array("tokenizer.ggml.scores", G.T_F32, scores)
    w.add_array("tokenizer.ggml.token_type", G.T_I32, types)
    w.add("tokenizer.ggml.bos_token_id", G.T_U32, 1)
    w.add("tokenizer.ggml.eos_token_id", G.T_U32, 2)
    w.add("tokenizer.ggml.unknown_token_id", G.T_U32, 0)
    w.add("tokenizer.ggml.add_bos_token", G.T_BOOL, True)
    w.add("tokenizer.ggml.add_eos_token", G.T_BOOL, False)
    for idx, (name, ne, tt, std) in enumerate(tensor_plan(cfg, ftype)):
        def make(idx=idx, name=name, ne=ne, tt=tt, std=std):
            return random_tensor(name, ne, tt, std, np.random.Generator(np.random.PCG64(seed + idx)))
        w.add_tensor(name, ne, tt, make)
    w.write(path)
    return {"config": asdict(cfg), "ftype": ftype, "seed": seed, "bytes_per_token": weight_bytes_per_token(cfg, ftype)}
