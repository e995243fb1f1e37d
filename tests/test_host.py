"""CPU tests of the host-side pieces: GGUF reader/writer against gguf-py, synthetic-model accounting against
SURVEY.md's byte table, tokenizers, sampler, the C-ABI library's exports, and the tile-SoA layout maps."""
import ctypes as C
import os
import re
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ----------------------------------------------------------------------------- GGUF container
def test_reader_matches_gguf_py_on_gguf_py_written_file(tmp_path):
    gguf = pytest.importorskip("gguf")
    from ggufb200 import gguf_reader as G
    p = str(tmp_path / "w.gguf")
    w = gguf.GGUFWriter(p, "llama")
    w.add_uint32("llama.block_count", 3)
    w.add_float32("llama.rope.freq_base", 500000.0)
    w.add_bool("tokenizer.ggml.add_bos_token", True)
    w.add_string("general.name", "héllo ✓")
    w.add_array("tokenizer.ggml.tokens", ["a", "▁b", "<0x0A>"])
    w.add_array("tokenizer.ggml.scores", [0.0, -1.5, -2.0])
    w.add_array("tokenizer.ggml.token_type", [1, 1, 6])
    rng = np.random.default_rng(0)
    q4 = rng.integers(0, 256, size=(7, 144 * 2), dtype=np.uint8)
    w.add_tensor("blk.0.attn_q.weight", q4, raw_dtype=gguf.GGMLQuantizationType.Q4_K)
    w.add_tensor("output_norm.weight", rng.standard_normal(512).astype(np.float32))
    w.add_tensor("h.weight", rng.standard_normal((3, 64)).astype(np.float16))
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    f = G.GGUFFile(p)
    r = gguf.GGUFReader(p)
    assert f.meta["llama.block_count"] == 3 and f.meta["general.name"] == "héllo ✓" and f.meta["tokenizer.ggml.add_bos_token"] is True
    assert f.meta["tokenizer.ggml.tokens"] == ["a", "▁b", "<0x0A>"] and list(f.meta["tokenizer.ggml.token_type"]) == [1, 1, 6]
    for t in r.tensors:
        ti = f.tensors[t.name]
        assert ti.ggml_type == int(t.tensor_type) and ti.ne == tuple(int(x) for x in t.shape) and ti.offset == t.data_offset
        assert np.array_equal(np.asarray(t.data).reshape(-1).view(np.uint8), f.data(t.name))
    assert f.tensors["blk.0.attn_q.weight"].ne == (512, 7)   # ne0 = K first


def test_reader_rejects_bad_files(tmp_path):
    from ggufb200 import gguf_reader as G
    p = tmp_path / "bad.gguf"
    p.write_bytes(b"NOPE" + b"\0" * 40)
    with pytest.raises(G.GGUFError, match="magic"):
        G.GGUFFile(str(p))
    p.write_bytes(struct.pack("<IIQQ", G.GGUF_MAGIC, 99, 0, 0))
    with pytest.raises(G.GGUFError, match="version"):
        G.GGUFFile(str(p))
    p.write_bytes(struct.pack("<IIQQ", G.GGUF_MAGIC, 3, 0, 5))       # promises 5 KV pairs, has none
    with pytest.raises(G.GGUFError, match="truncated"):
        G.GGUFFile(str(p))
    p.write_bytes(b"GG")
    with pytest.raises(G.GGUFError):
        G.GGUFFile(str(p))
    w = G.GGUFWriter()
    w.add("general.architecture", G.T_STR, "llama")
    with pytest.raises(G.GGUFError, match="bytes given"):
        w.add_tensor("x", (256, 2), G.GGML_Q4_K, np.zeros(10, np.uint8))


def test_synthetic_models_match_survey_byte_table_and_recipe():
    from ggufb200 import synth
    B = synth.weight_bytes_per_token
    assert B(synth.PRESETS["llama3-8b"], "Q4_K_M")["weights"] == 4_616_331_264
    assert B(synth.PRESETS["llama3-8b"], "Q4_K_M")["by_type"] == {"Q4_K": 3_359_637_504, "Q6_K": 1_256_693_760}
    assert B(synth.PRESETS["llama3-8b"], "Q8_0")["weights"] == 7_973_699_584
    assert B(synth.PRESETS["llama3-8b"], "Q6_K")["weights"] == 6_156_165_120
    assert B(synth.PRESETS["llama3-70b"], "Q4_K_M")["weights"] == 41_874_309_120
    assert B(synth.PRESETS["tinyllama-1.1b"], "Q4_K_M")["weights"] == 629_846_016
    assert B(synth.PRESETS["llama3-8b"], "Q4_K_M")["kv_per_pos"] == 131_072
    for name, n in (("tinyllama-1.1b", 10), ("llama3-8b", 16), ("llama3-70b", 40)):
        plan = synth.tensor_plan(synth.PRESETS[name], "Q4_K_M")
        assert sum(1 for t in plan if t[0].endswith("attn_v.weight") and t[2] == 14) == n


def test_synthetic_weights_hit_target_std_and_are_seeded(tmp_path, oracle):
    from ggufb200 import synth, gguf_reader as G
    a, b, c = (str(tmp_path / f"{x}.gguf") for x in "abc")
    synth.write_gguf(a, "small", "Q4_K_M", seed=1)
    synth.write_gguf(b, "small", "Q4_K_M", seed=1)
    synth.write_gguf(c, "small", "Q4_K_M", seed=2)
    assert open(a, "rb").read() == open(b, "rb").read() != open(c, "rb").read()
    f = G.GGUFFile(a)
    for name, target in (("blk.0.attn_q.weight", 512 ** -0.5), ("blk.3.ffn_down.weight", 1536 ** -0.5), ("token_embd.weight", 0.02)):
        t = f.tensors[name]
        w = oracle.dequantize(f.data(name), t.ggml_type, t.n_elements)
        assert abs(w.std() / target - 1) < 0.1 and abs(w.mean()) < 0.1 * target
    out = f.tensors["output.weight"]
    w = oracle.dequantize(f.data("output.weight"), out.ggml_type, out.n_elements).reshape(out.ne[1], out.ne[0])
    assert not w[:synth.N_SPECIAL].any() and w[synth.N_SPECIAL:].std() > 0.015   # special-token rows are exactly zero


# ----------------------------------------------------------------------------- tokenizers
def _spm_meta():
    toks = ["<unk>", "<s>", "</s>"] + [f"<0x{b:02X}>" for b in range(256)]
    types = [2, 3, 3] + [6] * 256
    words = ["▁", "h", "e", "l", "o", "▁h", "he", "ll", "hell", "▁hello", "hello", "▁w", "or", "ld", "▁world", "wor", "world", "lo", "▁he", "!", "▁hell", "▁wor", "w", "r", "d"]
    toks += words
    types += [1] * len(words)
    scores = [0.0] * 259 + [-float(i) for i in range(len(words))]
    scores[toks.index("▁hello")] = 5.0
    scores[toks.index("▁world")] = 4.0
    return {"tokenizer.ggml.model": "llama", "tokenizer.ggml.tokens": toks, "tokenizer.ggml.scores": scores,
            "tokenizer.ggml.token_type": types, "tokenizer.ggml.bos_token_id": 1, "tokenizer.ggml.eos_token_id": 2,
            "tokenizer.ggml.unknown_token_id": 0, "tokenizer.ggml.add_bos_token": True}


def test_spm_tokenizer_merges_by_score_byte_fallback_and_roundtrip():
    from ggufb200.tokenizer import StreamDecoder, Tokenizer
    t = Tokenizer(_spm_meta())
    ids = t.encode("hello world!")
    assert ids[0] == t.bos
    assert [t.tokens[i] for i in ids[1:]] == ["▁hello", "▁world", "!"]
    assert t.decode(ids) == " hello world!"
    ids = t.encode("héllo ✓", add_special=False)          # é and ✓ are not in the vocabulary -> UTF-8 byte tokens
    assert t.decode(ids) == " héllo ✓"
    assert sum(1 for i in ids if t.types[i] == 6) == len("é".encode()) + len("✓".encode())
    sd = StreamDecoder(t)                                   # multi-byte characters come out whole
    pieces = [sd.push(i) for i in ids]
    assert "".join(pieces) + sd.flush() == " héllo ✓" and all("�" not in p for p in pieces)
    assert t.encode("</s>", add_special=False) == [2]       # control tokens are matched verbatim
    assert t.encode("", add_special=False) == []


def test_spm_tokenizer_agrees_with_transformers_gguf_converter(tmp_path):
    """independent check: transformers builds a `tokenizers` Unigram/BPE tokenizer from the same GGUF metadata"""
    pytest.importorskip("transformers")
    from ggufb200.tokenizer import Tokenizer
    try:
        from transformers.integrations.ggml import GGUFLlamaConverter
    except Exception:
        pytest.skip("transformers without GGUF converters")
    meta = _spm_meta()
    d = {"tokens": meta["tokenizer.ggml.tokens"], "scores": meta["tokenizer.ggml.scores"], "token_type": meta["tokenizer.ggml.token_type"],
         "bos_token_id": 1, "eos_token_id": 2, "unk_token_id": 0, "tokenizer_type": "llama"}
    try:
        fast = GGUFLlamaConverter(d).converted()
    except Exception as e:
        pytest.skip(f"converter not usable here: {e!r}")
    t = Tokenizer(meta)
    for text in ("hello world!", "hello", "world hello hello", "held"):
        theirs = fast.encode(text, add_special_tokens=False).ids
        assert t.encode(text, add_special=False) == theirs, text


def test_bpe_tokenizer_and_chat_template():
    from ggufb200.tokenizer import Tokenizer, _B2U
    base = [_B2U[b] for b in range(256)]
    merges = ["h e", "l l", "he ll", "hell o", "Ġ w", "o r", "Ġw or", "Ġwor l", "Ġworl d"]
    toks = base + ["he", "ll", "hell", "hello", "Ġw", "or", "Ġwor", "Ġworl", "Ġworld", "<|begin_of_text|>", "<|eot_id|>",
                   "<|start_header_id|>", "<|end_header_id|>"]
    types = [1] * (len(toks) - 4) + [3] * 4
    tmpl = ("{{ bos_token }}{% for m in messages %}<|start_header_id|>{{ m['role'] }}<|end_header_id|>\n\n{{ m['content'] }}<|eot_id|>"
            "{% endfor %}{% if add_generation_prompt %}<|start_header_id|>assistant<|end_header_id|>\n\n{% endif %}")
    meta = {"tokenizer.ggml.model": "gpt2", "tokenizer.ggml.pre": "llama-bpe", "tokenizer.ggml.tokens": toks, "tokenizer.ggml.token_type": types,
            "tokenizer.ggml.merges": merges, "tokenizer.ggml.bos_token_id": toks.index("<|begin_of_text|>"),
            "tokenizer.ggml.eos_token_id": toks.index("<|eot_id|>"), "tokenizer.chat_template": tmpl}
    t = Tokenizer(meta)
    ids = t.encode("hello world", add_special=False)
    assert [t.tokens[i] for i in ids] == ["hello", "Ġworld"] and t.decode(ids) == "hello world"
    assert t.decode(t.encode("naïve ☃ 123", add_special=False)) == "naïve ☃ 123"
    ids = t.encode_chat([{"role": "user", "content": "hello"}])
    names = [t.tokens[i] for i in ids]
    assert names[0] == "<|begin_of_text|>" and names.count("<|begin_of_text|>") == 1      # template already adds BOS
    assert names[1] == "<|start_header_id|>" and "hello" in names and names[-3:] == ["<|end_header_id|>", "Ċ", "Ċ"]
    assert toks.index("<|eot_id|>") in t.eog
    t2 = Tokenizer({k: v for k, v in meta.items() if k != "tokenizer.chat_template"})
    assert t2.apply_chat_template([{"role": "user", "content": "hi"}]) == "<|im_start|>user\nhi<|im_end|>\n<|im_start|>assistant\n"


def test_sampler_is_seeded_and_respects_top_k_top_p():
    from ggufb200.scheduler import SamplingParams, sample_token
    logits = np.array([0.0, 5.0, 4.9, -3.0, 1.0], dtype=np.float32)
    assert sample_token(logits, SamplingParams(temperature=0.0), np.random.default_rng(0)) == 1
    assert sample_token(logits, SamplingParams(temperature=1.0, top_k=1), np.random.default_rng(0)) == 1
    draws = {sample_token(logits, SamplingParams(temperature=1.0, top_k=2, top_p=1.0), np.random.default_rng(s)) for s in range(50)}
    assert draws == {1, 2}
    draws = {sample_token(logits, SamplingParams(temperature=1.0, top_k=0, top_p=0.5), np.random.default_rng(s)) for s in range(50)}
    assert draws == {1}                                              # the top token alone covers p >= 0.5
    a = [sample_token(logits, SamplingParams(temperature=2.0, top_k=0, top_p=1.0), np.random.default_rng(3)) for _ in range(5)]
    b = [sample_token(logits, SamplingParams(temperature=2.0, top_k=0, top_p=1.0), np.random.default_rng(3)) for _ in range(5)]
    assert a == b


def test_sampler_cuts_on_the_untempered_distribution_then_applies_temperature():
    """upstream's default chain: top-k, top-p and min-p act on softmax(logits); temperature comes last.  Hand-computed:
    logits (2, 1, 0): p = (0.665, 0.245, 0.090).  top_p = 0.7 keeps two tokens whatever the temperature (tempering first
    at T = 4 would flatten p to (0.39, 0.33, 0.28) and keep all three); min_p = 0.2 drops only the last (0.090 < 0.133)."""
    from ggufb200.scheduler import SamplingParams, sample_token
    logits = np.array([2.0, 1.0, 0.0], dtype=np.float32)
    draws = [sample_token(logits, SamplingParams(temperature=4.0, top_k=0, top_p=0.7), np.random.default_rng(s)) for s in range(300)]
    assert set(draws) == {0, 1}
    # the survivors are then drawn at temperature 4: p0 = 1 / (1 + exp(-1/4)) = 0.562
    assert abs(draws.count(0) / 300 - 0.562) < 0.09
    draws = {sample_token(logits, SamplingParams(temperature=4.0, top_k=0, top_p=1.0, min_p=0.2), np.random.default_rng(s)) for s in range(200)}
    assert draws == {0, 1}


# ----------------------------------------------------------------------------- C-ABI library
def test_library_loads_and_exports_every_declared_symbol():
    from ggufb200 import cabi
    L = cabi.lib()
    hdr = open(os.path.join(ROOT, "include", "ggufb200.h")).read()
    declared = set(re.findall(r"\b(ggb_[a-z0-9_A-Z]+)\s*\(", hdr))
    assert declared == set(cabi.EXPORTS), declared ^ set(cabi.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.ggb_abi_version() == 3
    # entry points validate their arguments before touching the GPU (no compute call is made here)
    assert L.ggb_repacked_row_stride(12, 4096) == 2304 and L.ggb_repacked_row_stride(14, 5632) == 4624
    assert L.ggb_repacked_row_stride(12, 100) == -1 and L.ggb_repacked_row_stride(2, 4096) == -1
    assert L.ggb_dequant(10, None, None, 256, None) == -3 and b"unsupported" in L.ggb_last_error()
    assert L.ggb_repack(12, None, None, 4, 100, None) == -1
    assert L.ggb_quantize_q8_K(None, None, None, None, 100, 1, None) == -1
    assert L.ggb_attn_decode(None, None, None, None, 8, 2, 128, 64, None, None, 0, None) == -1
    a = cabi.make_gemv_args([(0, 12, 4, 0)], 100, 0)
    assert L.ggb_gemv(C.byref(a), None) == -1 and b"multiple of 256" in L.ggb_last_error()
    assert os.path.getmtime(cabi.LIB_PATH) > 0


def test_engine_refuses_to_run_without_cuda(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ggufb200 import synth
    from ggufb200.cabi import GGBError
    from ggufb200.model import Engine
    p = str(tmp_path / "t.gguf")
    synth.write_gguf(p, "tiny", "Q4_K_M")
    with pytest.raises(GGBError, match="no CPU fallback"):
        Engine(p)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "llama-gguf-inference_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "libggml_ref" not in src and "oracle/_build" not in src, f


# ----------------------------------------------------------------------------- tile-SoA layout (host build of layout.cuh)
@pytest.fixture(scope="module")
def layout_lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("layout")
    src = d / "layout_host.c"
    src.write_text('#include <stdint.h>\n#include "ggufb200.h"\n#include "layout.cuh"\n'
                   "int64_t map(int type, int64_t k, int64_t o) { return ggb_repacked_to_canon(type, k, o); }\n"
                   "int64_t stride(int type, int64_t k) { return ggb_row_stride(type, k); }\n"
                   "int hdr2(const uint8_t* s, int i) { return ggb_hdr2_byte(s, i); }\n"
                   "void dec(const uint8_t* h, int j, int* sc, int* mn) { ggb_hdr2_scale_min(h, j, sc, mn); }\n")
    so = d / "layout_host.so"
    subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-O1", "-shared", "-fPIC", "-x", "c", str(src), "-o", str(so),
                    "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "llama-gguf-inference_b200", "csrc")], check=True)
    L = C.CDLL(str(so))
    L.map.restype = C.c_int64; L.map.argtypes = [C.c_int, C.c_int64, C.c_int64]
    L.stride.restype = C.c_int64; L.stride.argtypes = [C.c_int, C.c_int64]
    return L


@pytest.mark.parametrize("qtype,sbb", [(12, 144), (14, 210), (8, 272), (13, 176)])
@pytest.mark.parametrize("k", [256, 2048, 2304, 5632])
def test_repack_map_is_a_bijection_on_row_bytes(layout_lib, qtype, sbb, k):
    """every canonical byte lands exactly once in the tile-SoA row (except the fields that are re-encoded)"""
    row = k // 256 * sbb
    seen = np.zeros(row, dtype=np.int32)
    special = 0
    for o in range(row):
        s = layout_lib.map(qtype, k, o)
        if s < 0:
            special += 1
        else:
            assert 0 <= s < row
            seen[s] += 1
    assert seen.max() <= 1
    if qtype == 14 or qtype == 8:
        assert special == 0 and seen.min() == 1
    else:   # Q4_K/Q5_K: 12 re-encoded scale bytes per super-block (+ 32 gathered qh bytes for Q5_K)
        per_sb = 12 + (32 if qtype == 13 else 0)
        assert special == per_sb * (k // 256) and (seen == 0).sum() == special
    assert 0 <= layout_lib.stride(qtype, k) - row < 16


def test_header_scale_regrouping_roundtrip(layout_lib):
    rng = np.random.default_rng(0)
    from oracle import oracle as O  # only for its reference get_scale_min semantics via dequant is overkill; restate inline
    for _ in range(200):
        s = rng.integers(0, 256, 12, dtype=np.uint8)
        hdr = np.zeros(16, dtype=np.uint8)
        for i in range(12):
            hdr[4 + i] = layout_lib.hdr2(s.ctypes.data_as(C.POINTER(C.c_uint8)), i)
        for j in range(8):
            if j < 4:
                sc, mn = s[j] & 63, s[j + 4] & 63
            else:
                sc, mn = (s[j + 4] & 0xF) | ((s[j - 4] >> 6) << 4), (s[j + 4] >> 4) | ((s[j] >> 6) << 4)
            a, b = C.c_int(), C.c_int()
            layout_lib.dec(hdr.ctypes.data_as(C.POINTER(C.c_uint8)), j, C.byref(a), C.byref(b))
            assert (a.value, b.value) == (int(sc), int(mn))


def test_sampler_min_p_and_penalties():
    """the samplers the reference's API documents beyond temperature/top-k/top-p (docs/API_REFERENCE.md:369-379)"""
    from ggufb200.scheduler import SamplingParams, apply_penalties, sample_token
    rng = np.random.default_rng(0)
    logits = np.array([2.0, 1.9, -1.0, 0.5, -3.0], dtype=np.float32)
    # min_p keeps only tokens with p >= min_p * p_max: at 0.5 only tokens 0 and 1 survive
    sp = SamplingParams(temperature=1.0, top_k=0, top_p=1.0, min_p=0.5, seed=1)
    assert {sample_token(logits, sp, rng) for _ in range(200)} == {0, 1}
    # repeat penalty: positive logits divided, negative multiplied; frequency counts occurrences; presence once
    sp = SamplingParams(temperature=0.0, repeat_penalty=2.0, presence_penalty=0.25, frequency_penalty=0.5, repeat_last_n=4)
    hist = [4, 0, 0, 2, 0]           # window = last 4 -> token 0 three times, token 2 once; the leading 4 is outside
    x = apply_penalties(logits, sp, hist)
    assert np.allclose(x, [2.0 / 2 - 3 * 0.5 - 0.25, 1.9, -1.0 * 2 - 0.5 - 0.25, 0.5, -3.0])
    assert not sp.greedy and sp.arg_max                    # decided on the host: the device arg-max sees raw logits
    assert sample_token(logits, sp, rng, hist) == 1        # token 0 was the raw arg-max
    assert sample_token(logits, SamplingParams(temperature=0.0), rng, hist) == 0
    assert SamplingParams(temperature=0.0).greedy
    # ids outside the vocabulary in the history are ignored
    assert np.array_equal(apply_penalties(logits, sp, [99, -1]), logits)


def test_sampling_from_device_candidates_equals_sampling_from_the_row():
    """scheduler.sample_from_candidates on {logits >= the k-th largest} (what ggb_topk_rows hands back, unordered) picks the token
    sample_token picks from the whole row, for the same seed: top-k / top-p / min-p / temperature see the same sorted k numbers"""
    from ggufb200.scheduler import SamplingParams, device_topk_ok, sample_from_candidates, sample_token
    rng = np.random.default_rng(7)
    for trial in range(40):
        n = int(rng.integers(300, 5000))
        logits = (rng.standard_normal(n) * 3).astype(np.float32)
        if trial % 5 == 0:
            logits[rng.integers(0, n, 40)] = logits.max()          # ties at the top
        sp = SamplingParams(temperature=float(rng.uniform(0.2, 1.5)), top_k=int(rng.integers(2, 100)), top_p=float(rng.choice([1.0, 0.95, 0.5])),
                            min_p=float(rng.choice([0.0, 0.05])), seed=trial)
        assert device_topk_ok(sp)
        kth = np.sort(logits)[-sp.top_k]
        idx = np.flatnonzero(logits >= kth).astype(np.int32)
        rng.shuffle(idx)
        a = sample_token(logits, sp, np.random.default_rng(trial))
        b = sample_from_candidates(idx, logits[idx], sp, np.random.default_rng(trial))
        assert a == b, trial
    assert not device_topk_ok(SamplingParams(temperature=0.0))                       # greedy: the arg-max kernel
    assert device_topk_ok(SamplingParams(repeat_penalty=1.1), [3, 4, 5])             # penalised: candidates + the window's logits
    assert not device_topk_ok(SamplingParams(repeat_penalty=1.1, repeat_last_n=0), list(range(300)))   # window too wide for the buffer
    assert not device_topk_ok(SamplingParams(top_k=0)) and not device_topk_ok(SamplingParams(top_k=1000))


def test_penalised_sampling_from_candidates_and_window_logits_equals_sampling_from_the_row():
    """penalties only touch the window tokens, so the top-(k + window) raw candidates plus the window's raw logits contain the
    penalised top-k: scheduler.sample_from_candidates_penalised picks sample_token's token, greedy-with-penalties included"""
    from ggufb200.scheduler import SamplingParams, device_topk_ok, penalty_window, sample_from_candidates_penalised, sample_token
    rng = np.random.default_rng(11)
    for trial in range(40):
        n = int(rng.integers(400, 4000))
        logits = (rng.standard_normal(n) * 3).astype(np.float32)
        top = np.argsort(-logits)
        hist = [int(t) for t in rng.choice(top[:30], size=int(rng.integers(1, 50)))] + [int(t) for t in rng.integers(0, n, 10)] + [n + 5]
        sp = SamplingParams(temperature=0.0 if trial % 7 == 0 else float(rng.uniform(0.3, 1.2)), top_k=int(rng.integers(2, 60)),
                            top_p=float(rng.choice([1.0, 0.9])), repeat_penalty=float(rng.choice([1.0, 1.1, 1.5])),
                            presence_penalty=float(rng.choice([0.0, 0.5])), frequency_penalty=float(rng.choice([0.0, 0.3])),
                            repeat_last_n=int(rng.choice([64, 16, 0])), seed=trial)
        if not sp.penalised:
            continue
        assert device_topk_ok(sp, hist)
        win = penalty_window(sp, hist)
        k = (1 if sp.arg_max else sp.top_k) + len(win)
        kth = np.sort(logits)[-min(k, n)]
        idx = np.flatnonzero(logits >= kth).astype(np.int32)
        rng.shuffle(idx)
        win_vals = np.array([logits[t] if 0 <= t < n else 0.0 for t in win], dtype=np.float32)   # what ggb_gather_rows returns
        a = sample_token(logits, sp, np.random.default_rng(trial), hist)
        b = sample_from_candidates_penalised(idx, logits[idx], win, win_vals, sp, np.random.default_rng(trial), hist)
        assert a == b, trial
