"""Tokenizer, detokenizer and chat templating from GGUF metadata (host side).

The reference's backend receives OpenAI `messages` (scripts/benchmark.py:294-301, docs/API_REFERENCE.md:369-379)
and must turn them into token ids and stream text pieces back; upstream does this in src/llama-vocab.cpp and
the server's chat-template code [UPSTREAM-MEM].  Two vocab families are implemented, selected by
`tokenizer.ggml.model`:

  "llama"  SentencePiece-style unigram/BPE hybrid: text -> '▁'-escaped UTF-8 characters, greedy merge of the
           adjacent pair whose concatenation has the highest vocabulary score, unknown pieces fall back to
           <0xXX> byte tokens;
  "gpt2"   byte-level BPE: regex pre-tokeniser (llama-bpe / gpt-2 patterns), bytes mapped to printable code
           points, merges applied by rank.

Special (control / user-defined) tokens are matched verbatim before either algorithm runs, so chat-template
markers such as <|start_header_id|> become single ids.
"""
from __future__ import annotations

import heapq
import json

TT_NORMAL, TT_UNKNOWN, TT_CONTROL, TT_USER, TT_UNUSED, TT_BYTE = 1, 2, 3, 4, 5, 6

# pre-tokeniser patterns [UPSTREAM-MEM: llama-vocab.cpp]; they need the third-party `regex` module (\p classes)
_PRE_LLAMA3 = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+")
_PRE_GPT2 = r"'s|'t|'re|'ve|'m|'ll|'d| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"

CHATML_FALLBACK = "chatml"


def _bytes_to_unicode():
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) + list(range(ord("®"), ord("ÿ") + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {b: chr(c) for b, c in zip(bs, cs)}


_B2U = _bytes_to_unicode()
_U2B = {v: k for k, v in _B2U.items()}


class Tokenizer:
    def __init__(self, meta: dict):
        g = meta.get
        self.model = g("tokenizer.ggml.model", "llama")
        self.tokens: list[str] = list(g("tokenizer.ggml.tokens", []))
        if not self.tokens:
            raise ValueError("GGUF file has no tokenizer.ggml.tokens")
        n = len(self.tokens)
        scores = g("tokenizer.ggml.scores")
        self.scores = [float(x) for x in scores] if scores is not None and len(scores) == n else [0.0] * n
        types = g("tokenizer.ggml.token_type")
        self.types = [int(x) for x in types] if types is not None and len(types) == n else [TT_NORMAL] * n
        self.bos = self._id(g("tokenizer.ggml.bos_token_id"))
        self.eos = self._id(g("tokenizer.ggml.eos_token_id"))
        self.eot = self._id(g("tokenizer.ggml.eot_token_id"))
        self.unk = self._id(g("tokenizer.ggml.unknown_token_id"))
        self.add_bos = bool(g("tokenizer.ggml.add_bos_token", self.model == "llama"))
        self.add_eos = bool(g("tokenizer.ggml.add_eos_token", False))
        self.add_space_prefix = bool(g("tokenizer.ggml.add_space_prefix", True))
        self.chat_template = g("tokenizer.chat_template")
        self.pre = g("tokenizer.ggml.pre", "default")
        self.tok2id = {}
        for i, t in enumerate(self.tokens):
            self.tok2id.setdefault(t, i)
        self.byte_tok = {}
        for i, (t, ty) in enumerate(zip(self.tokens, self.types)):
            if ty == TT_BYTE and len(t) == 6 and t.startswith("<0x") and t.endswith(">"):
                self.byte_tok[int(t[3:5], 16)] = i
        # specials are matched verbatim, longest first
        self.specials = sorted(((t, i) for i, (t, ty) in enumerate(zip(self.tokens, self.types)) if ty in (TT_CONTROL, TT_USER) and t),
                               key=lambda x: -len(x[0]))
        self.eog = {i for i in (self.eos, self.eot) if i is not None}
        for name in ("<|eot_id|>", "<|im_end|>", "<|end|>", "<end_of_turn>", "<|endoftext|>"):
            if name in self.tok2id and self.types[self.tok2id[name]] in (TT_CONTROL, TT_USER):
                self.eog.add(self.tok2id[name])
        if self.model == "gpt2":
            merges = g("tokenizer.ggml.merges", [])
            self.ranks = {}
            for r, m in enumerate(merges):
                a, _, b = m.partition(" ")
                self.ranks[(a, b)] = r
            import regex  # third-party, present in the image
            self._pre_re = regex.compile(_PRE_LLAMA3 if self.pre in ("llama-bpe", "llama3", "llama-v3") else _PRE_GPT2)
            self._bpe_cache: dict[str, list[int]] = {}

    def _id(self, v):
        return None if v is None else int(v)

    @property
    def n_vocab(self) -> int:
        return len(self.tokens)

    # ------------------------------------------------------------------ encode
    def encode(self, text: str, add_special: bool = True, parse_special: bool = True) -> list[int]:
        out: list[int] = []
        if add_special and self.add_bos and self.bos is not None:
            out.append(self.bos)
        first = True
        for chunk, special_id in self._split_specials(text) if parse_special else [(text, None)]:
            if special_id is not None:
                out.append(special_id)
                first = False
            elif chunk:
                out.extend(self._encode_spm(chunk, first) if self.model != "gpt2" else self._encode_bpe(chunk))
                first = False
        if add_special and self.add_eos and self.eos is not None:
            out.append(self.eos)
        return out

    def _split_specials(self, text: str):
        if not self.specials:
            yield text, None
            return
        i, start, n = 0, 0, len(text)
        first_chars = {t[0] for t, _ in self.specials}
        while i < n:
            if text[i] in first_chars:
                for t, tid in self.specials:
                    if text.startswith(t, i):
                        if i > start:
                            yield text[start:i], None
                        yield t, tid
                        i += len(t)
                        start = i
                        break
                else:
                    i += 1
            else:
                i += 1
        if start < n:
            yield text[start:], None

    def _encode_spm(self, text: str, is_first: bool) -> list[int]:
        if self.add_space_prefix and is_first:
            text = " " + text
        text = text.replace(" ", "▁")
        syms = list(text)
        n = len(syms)
        if n == 0:
            return []
        prev = list(range(-1, n - 1))
        nxt = list(range(1, n + 1))
        nxt[-1] = -1
        alive = [True] * n
        heap = []

        def push(i, j):
            if i < 0 or j < 0:
                return
            piece = syms[i] + syms[j]
            tid = self.tok2id.get(piece)
            if tid is not None:
                heapq.heappush(heap, (-self.scores[tid], i, j, piece))

        for i in range(n - 1):
            push(i, i + 1)
        while heap:
            _, i, j, piece = heapq.heappop(heap)
            if not (alive[i] and alive[j]) or nxt[i] != j or syms[i] + syms[j] != piece:
                continue
            syms[i] = piece
            alive[j] = False
            nxt[i] = nxt[j]
            if nxt[j] >= 0:
                prev[nxt[j]] = i
            push(prev[i], i)
            push(i, nxt[i])
        out = []
        i = 0
        while i >= 0:
            tid = self.tok2id.get(syms[i])
            if tid is not None and self.types[tid] != TT_BYTE:
                out.append(tid)
            else:
                for b in syms[i].encode("utf-8"):
                    bt = self.byte_tok.get(b)
                    out.append(bt if bt is not None else (self.unk if self.unk is not None else 0))
            i = nxt[i]
        return out

    def _bpe_word(self, word: str) -> list[int]:
        hit = self._bpe_cache.get(word)
        if hit is not None:
            return hit
        parts = [_B2U[b] for b in word.encode("utf-8")]
        while len(parts) > 1:
            best, bi = None, -1
            for i in range(len(parts) - 1):
                r = self.ranks.get((parts[i], parts[i + 1]))
                if r is not None and (best is None or r < best):
                    best, bi = r, i
            if best is None:
                break
            parts[bi:bi + 2] = [parts[bi] + parts[bi + 1]]
        ids = []
        for p in parts:
            tid = self.tok2id.get(p)
            if tid is not None:
                ids.append(tid)
            else:  # unmergeable piece: emit its single-byte tokens
                ids.extend(self.tok2id.get(ch, self.unk if self.unk is not None else 0) for ch in p)
        if len(self._bpe_cache) < 100000:
            self._bpe_cache[word] = ids
        return ids

    def _encode_bpe(self, text: str) -> list[int]:
        out = []
        for w in self._pre_re.findall(text):
            out.extend(self._bpe_word(w))
        return out

    # ------------------------------------------------------------------ decode
    def token_bytes(self, tid: int, special: bool = False) -> bytes:
        """UTF-8 bytes of one token's text piece (control tokens render as b'' unless special=True)."""
        if tid < 0 or tid >= len(self.tokens):
            return b""
        ty, t = self.types[tid], self.tokens[tid]
        if ty in (TT_CONTROL, TT_UNUSED) or (ty == TT_UNKNOWN and not special):
            return t.encode("utf-8") if special else b""
        if ty == TT_BYTE and tid in self.byte_tok.values():
            return bytes([int(t[3:5], 16)])
        if self.model == "gpt2":
            if ty == TT_USER:
                return t.encode("utf-8")
            return bytes(_U2B.get(ch, ord("?")) if ch in _U2B else 63 for ch in t)
        return t.replace("▁", " ").encode("utf-8")

    def decode(self, ids, special: bool = False) -> str:
        return b"".join(self.token_bytes(i, special) for i in ids).decode("utf-8", errors="replace")

    # ------------------------------------------------------------------ chat template
    def apply_chat_template(self, messages: list[dict], add_generation_prompt: bool = True) -> str:
        msgs = []
        for m in messages:
            content = m.get("content", "")
            if isinstance(content, list):  # OpenAI content parts
                content = "".join(p.get("text", "") for p in content if isinstance(p, dict) and p.get("type") == "text")
            msgs.append({"role": str(m.get("role", "user")), "content": "" if content is None else str(content)})
        if self.chat_template:
            try:
                from jinja2 import BaseLoader
                from jinja2.sandbox import ImmutableSandboxedEnvironment

                env = ImmutableSandboxedEnvironment(loader=BaseLoader(), trim_blocks=True, lstrip_blocks=True)
                env.globals["raise_exception"] = _raise
                env.filters["tojson"] = lambda x, **kw: json.dumps(x, ensure_ascii=False)
                bos = self.tokens[self.bos] if self.bos is not None else ""
                eos = self.tokens[self.eos] if self.eos is not None else ""
                return env.from_string(self.chat_template).render(messages=msgs, add_generation_prompt=add_generation_prompt,
                                                                  bos_token=bos, eos_token=eos)
            except ImportError:
                pass
            except Exception as e:  # a broken template must not take the server down: fall back
                self.template_error = repr(e)
        out = "".join(f"<|im_start|>{m['role']}\n{m['content']}<|im_end|>\n" for m in msgs)
        return out + ("<|im_start|>assistant\n" if add_generation_prompt else "")

    def template_adds_bos(self, rendered: str) -> bool:
        return self.bos is not None and rendered.startswith(self.tokens[self.bos])

    def encode_chat(self, messages: list[dict]) -> list[int]:
        text = self.apply_chat_template(messages)
        ids = self.encode(text, add_special=not self.template_adds_bos(text), parse_special=True)
        return ids


def _raise(msg):
    raise ValueError(msg)


class StreamDecoder:
    """Incremental detokeniser: emits text as soon as the accumulated bytes form valid UTF-8."""

    def __init__(self, tok: Tokenizer):
        self.tok, self.buf = tok, b""

    def push(self, tid: int) -> str:
        self.buf += self.tok.token_bytes(tid)
        try:
            s = self.buf.decode("utf-8")
            self.buf = b""
            return s
        except UnicodeDecodeError as e:
            if e.start > 0:  # emit the valid prefix, keep the incomplete tail
                s = self.buf[:e.start].decode("utf-8")
                self.buf = self.buf[e.start:]
                if len(self.buf) > 4:  # not an incomplete sequence but garbage: flush with replacement
                    s += self.buf.decode("utf-8", errors="replace")
                    self.buf = b""
                return s
            if len(self.buf) > 4:
                s = self.buf.decode("utf-8", errors="replace")
                self.buf = b""
                return s
            return ""

    def flush(self) -> str:
        s = self.buf.decode("utf-8", errors="replace")
        self.buf = b""
        return s
