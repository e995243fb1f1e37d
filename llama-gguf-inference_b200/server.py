"""HTTP boundary of the backend process: the contract the reference's gateway and launcher depend on.

Replaces `/app/llama-server`'s HTTP surface as used by the reference:
  * scripts/gateway.py:699-804  -- every request arrives on a NEW connection, `Connection: close`, header names
    lower-cased, `Authorization: Bearer <backend key>`; the body is relayed in <= 8 KiB reads UNTIL EOF, so each
    response ends by closing the socket (no chunked encoding, SSE events flushed as they are produced);
  * scripts/gateway.py:326-376  -- `GET /health` WITHOUT a key, one `read(4096)`: the answer is public, < 4 KiB,
    JSON, and written with a single send();
  * scripts/start.sh:600-646    -- `/health` must answer 2xx within 30 polls; other paths reject a missing key;
  * docs/API_REFERENCE.md:341-535 -- payload shapes of /v1/chat/completions (JSON + SSE), /v1/completions,
    /v1/models; errors in the OpenAI shape (docs/API_REFERENCE.md:670-682);
  * scripts/benchmark.py:279-384 -- streaming client: lines starting `data:`, first content delta = TTFT.
"""
from __future__ import annotations

import hmac
import json
import socket
import threading
import time
import uuid
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

from .scheduler import Request, SamplingParams

PUBLIC_PATHS = {"/health", "/v1/health", "/models", "/v1/models"}
MAX_BODY = 16 * 1024 * 1024


class ServerState:
    def __init__(self, scheduler, tokenizer, model_name: str, api_key: str | None, n_ctx: int, defaults: SamplingParams,
                 log=print, info: dict | None = None):
        self.sched, self.tok, self.model_name, self.api_key = scheduler, tokenizer, model_name, api_key
        self.n_ctx, self.defaults, self.log = n_ctx, defaults, log
        self.ready = threading.Event()
        self.t_start = time.time()
        self.info = info or {}
        self.embeddings = ""       # "" = not served (501, as llama-server without --embeddings); else the pooling: "mean" / "last"


def _err(code: int, message: str, etype: str) -> tuple[int, bytes]:
    return code, json.dumps({"error": {"code": code, "message": message, "type": etype}}).encode()


class Handler(BaseHTTPRequestHandler):
    protocol_version = "HTTP/1.1"
    server_version = "ggufb200"
    state: ServerState = None  # set on the server class

    # ---- plumbing
    def log_message(self, fmt, *args):  # one plain line per request on stdout (start.sh tees it to the server log)
        self.state.log("srv  request: %s %s" % (self.address_string(), fmt % args))

    def _send(self, code: int, body: bytes, ctype: str = "application/json; charset=utf-8", extra: dict | None = None):
        reason = {200: "OK", 400: "Bad Request", 401: "Unauthorized", 404: "Not Found", 405: "Method Not Allowed",
                  413: "Payload Too Large", 500: "Internal Server Error", 501: "Not Implemented", 503: "Service Unavailable"}.get(code, "OK")
        head = [f"HTTP/1.1 {code} {reason}", f"Content-Type: {ctype}", f"Content-Length: {len(body)}",
                "Connection: close", "Server: ggufb200", "Access-Control-Allow-Origin: *"]
        for k, v in (extra or {}).items():
            head.append(f"{k}: {v}")
        # headers + body in ONE send: the gateway's health probe does a single read(4096)
        self.wfile.write(("\r\n".join(head) + "\r\n\r\n").encode() + body)
        self.wfile.flush()
        self.close_connection = True

    def _authorized(self) -> bool:
        key = self.state.api_key
        if not key:
            return True
        h = self.headers.get("Authorization", "")
        tok = h[7:].strip() if h[:7].lower() == "bearer " else self.headers.get("X-Api-Key", "")
        return hmac.compare_digest(tok.encode(), key.encode())

    def _path(self) -> str:
        return self.path.split("?", 1)[0].rstrip("/") or "/"

    def _gate(self) -> bool:
        """common checks; returns False when a response has been sent"""
        p = self._path()
        if p not in PUBLIC_PATHS and not self._authorized():
            self._send(*_err(401, "Invalid API Key", "authentication_error"))
            return False
        return True

    # ---- GET
    def do_GET(self):
        p = self._path()
        if not self._gate():
            return
        st = self.state
        if p in ("/health", "/v1/health"):
            if st.sched.fatal:
                return self._send(*_err(500, st.sched.fatal, "server_error"))
            if not st.ready.is_set():
                return self._send(*_err(503, "Loading model", "unavailable_error"))
            return self._send(200, json.dumps({"status": "ok", "slots_idle": st.sched.idle_slots(),
                                               "slots_processing": len(st.sched.active)}).encode())
        if p in ("/v1/models", "/models"):
            m = {"id": st.model_name, "object": "model", "created": int(st.t_start), "owned_by": "ggufb200", "meta": st.info}
            return self._send(200, json.dumps({"object": "list", "data": [m]}).encode())
        if p == "/props":
            return self._send(200, json.dumps({"model_path": st.model_name, "n_ctx": st.n_ctx, "total_slots": len(st.sched.engine.slots),
                                               "chat_template": st.tok.chat_template or "", "build_info": st.info}).encode())
        if p == "/metrics":
            s = st.sched.stats
            tps = s["completion_tokens"] / s["decode_seconds"] if s["decode_seconds"] > 0 else 0.0
            body = (f"llamacpp:prompt_tokens_total {s['prompt_tokens']}\nllamacpp:tokens_predicted_total {s['completion_tokens']}\n"
                    f"llamacpp:predicted_tokens_seconds {tps:.3f}\nllamacpp:requests_processing {len(st.sched.active)}\n"
                    f"llamacpp:requests_deferred {len(st.sched.pending)}\n"
                    f"ggufb200:batched_steps_total {s.get('batched_steps', 0)}\nggufb200:chained_steps_total {s.get('chained_steps', 0)}\n"
                    f"ggufb200:gpu_wait_seconds_total {s.get('gpu_wait_seconds', 0.0):.4f}\n"
                    f"ggufb200:device_topk_tokens_total {s.get('device_topk_tokens', 0)}\n")
            return self._send(200, body.encode(), "text/plain; version=0.0.4")
        if p == "/":
            return self._send(200, b'{"status":"ok","server":"ggufb200"}')
        self._send(*_err(404, f"File Not Found: {p}", "not_found_error"))

    def do_OPTIONS(self):
        self._send(200, b"", extra={"Access-Control-Allow-Methods": "GET, POST, OPTIONS",
                                    "Access-Control-Allow-Headers": "Content-Type, Authorization"})

    # ---- POST
    def do_POST(self):
        p = self._path()
        if not self._gate():
            return
        st = self.state
        try:
            n = int(self.headers.get("Content-Length", "0") or 0)
        except ValueError:
            return self._send(*_err(400, "invalid Content-Length", "invalid_request_error"))
        if n < 0:
            return self._send(*_err(400, "invalid Content-Length", "invalid_request_error"))
        if n > MAX_BODY:
            return self._send(*_err(413, "request body too large", "invalid_request_error"))
        raw = self.rfile.read(n) if n else b""
        try:
            body = json.loads(raw.decode("utf-8")) if raw else {}
            if not isinstance(body, dict):
                raise ValueError("body must be a JSON object")
        except Exception as e:
            return self._send(*_err(400, f"invalid JSON body: {e}", "invalid_request_error"))
        if not st.ready.is_set():
            return self._send(*_err(503, "Loading model", "unavailable_error"))
        try:
            if p in ("/v1/chat/completions", "/chat/completions"):
                return self._completion(body, chat=True)
            if p in ("/v1/completions", "/completions", "/completion"):
                return self._completion(body, chat=False)
            if p == "/tokenize":
                ids = st.tok.encode(str(body.get("content", "")), add_special=bool(body.get("add_special", False)))
                return self._send(200, json.dumps({"tokens": ids}).encode())
            if p == "/detokenize":
                return self._send(200, json.dumps({"content": st.tok.decode([int(t) for t in body.get("tokens", [])])}).encode())
            if p in ("/v1/embeddings", "/embeddings", "/embedding"):
                if not st.embeddings:   # upstream's answer when the server was not started with --embeddings
                    return self._send(*_err(501, "This server does not support embeddings. Start it with `--embeddings`", "not_supported_error"))
                return self._embeddings(body)
        except (BrokenPipeError, ConnectionResetError):
            return
        except (ValueError, TypeError) as e:   # malformed field (wrong type, out-of-range token id, ...): this request only
            return self._send(*_err(400, str(e), "invalid_request_error"))
        self._send(*_err(404, f"File Not Found: {p}", "not_found_error"))

    def _build_request(self, body: dict, chat: bool) -> Request:
        st = self.state
        if chat:
            msgs = body.get("messages")
            if not isinstance(msgs, list) or not msgs:
                raise ValueError("'messages' is required and must be a non-empty array")
            for m in msgs:
                if not isinstance(m, dict) or "role" not in m:
                    raise ValueError("every message needs a 'role'")
            ids = st.tok.encode_chat(msgs)
        else:
            prompt = body.get("prompt")
            if isinstance(prompt, list) and prompt and all(isinstance(t, int) and not isinstance(t, bool) for t in prompt):
                # raw token ids go straight to the embedding gather on the device: range-check them here
                n_vocab = st.tok.n_vocab
                bad = [t for t in prompt if not 0 <= t < n_vocab]
                if bad:
                    raise ValueError(f"prompt token id {bad[0]} is outside the vocabulary (0..{n_vocab - 1})")
                ids = list(prompt)
            elif isinstance(prompt, (str, list)):
                ids = st.tok.encode(prompt if isinstance(prompt, str) else "".join(map(str, prompt)))
            else:
                raise ValueError("'prompt' is required")
        if not ids:
            raise ValueError("the prompt is empty after tokenisation")
        def num(key, default, cast):
            """a JSON null means "use the default"; anything that is not a plain number is the client's error (400)"""
            v = body.get(key)
            if v is None:
                return default
            if isinstance(v, bool) or not isinstance(v, (int, float)):
                raise ValueError(f"'{key}' must be a number")
            return cast(v)

        mt = None
        for key in ("max_tokens", "max_completion_tokens", "n_predict"):
            if body.get(key) is not None:
                mt = num(key, None, int)
                break
        max_tokens = st.n_ctx if mt is None or mt < 0 else mt
        d = st.defaults
        seed = num("seed", None, int)
        sp = SamplingParams(temperature=num("temperature", d.temperature, float),
                            top_k=num("top_k", d.top_k, int), top_p=num("top_p", d.top_p, float),
                            seed=(seed if seed not in (None, -1) else d.seed),
                            min_p=num("min_p", d.min_p, float),
                            repeat_penalty=num("repeat_penalty", d.repeat_penalty, float),
                            presence_penalty=num("presence_penalty", d.presence_penalty, float),
                            frequency_penalty=num("frequency_penalty", d.frequency_penalty, float),
                            repeat_last_n=num("repeat_last_n", d.repeat_last_n, int))
        stop = body.get("stop") or []
        if isinstance(stop, str):
            stop = [stop]
        stop = [s for s in stop if isinstance(s, str) and s]
        return Request(prompt_ids=ids, max_tokens=max_tokens, sampling=sp, stop=stop, ignore_eos=bool(body.get("ignore_eos", False)),
                       cache_prompt=bool(body.get("cache_prompt", True)))

    def _embeddings(self, body: dict):
        """POST /v1/embeddings (docs/API_REFERENCE.md:540-590 of the reference): `input` is a string, an array of strings, or
        token-id arrays; one pooled, L2-normalised vector per input, in order."""
        st = self.state
        inp = body.get("input", body.get("content"))
        if isinstance(inp, str) or (isinstance(inp, list) and inp and all(isinstance(t, int) and not isinstance(t, bool) for t in inp)):
            inp = [inp]
        if not isinstance(inp, list) or not inp:
            raise ValueError("'input' is required: a string, an array of strings, or arrays of token ids")
        n_vocab = st.tok.n_vocab
        reqs = []
        for item in inp:
            if isinstance(item, str):
                ids = st.tok.encode(item, add_special=True)
            elif isinstance(item, list) and item and all(isinstance(t, int) and not isinstance(t, bool) for t in item):
                if any(not 0 <= t < n_vocab for t in item):
                    raise ValueError(f"token ids must be in 0..{n_vocab - 1}")
                ids = list(item)
            else:
                raise ValueError("every input must be a non-empty string or a non-empty array of token ids")
            if not ids:
                raise ValueError("an input tokenises to nothing")
            if len(ids) + 1 >= st.n_ctx:
                raise ValueError(f"input of {len(ids)} tokens exceeds the context size {st.n_ctx}")
            reqs.append(st.sched.submit(Request(prompt_ids=ids, max_tokens=0, embed=st.embeddings)))
        data, total = [], 0
        for i, r in enumerate(reqs):
            ev = r.events.get(timeout=600)
            if ev[0] != "embedding":
                code = 400 if str(ev[1]).startswith("invalid request") else 500
                return self._send(*_err(code, str(ev[1]), "invalid_request_error" if code == 400 else "server_error"))
            data.append({"object": "embedding", "embedding": [float(x) for x in ev[1]], "index": i})
            total += ev[2]
        out = {"object": "list", "data": data, "model": str(body.get("model") or st.model_name),
               "usage": {"prompt_tokens": total, "total_tokens": total}}
        self._send(200, json.dumps(out).encode())

    def _completion(self, body: dict, chat: bool):
        st = self.state
        req = self._build_request(body, chat)
        stream = bool(body.get("stream", False))
        rid = ("chatcmpl-" if chat else "cmpl-") + uuid.uuid4().hex[:24]
        created = int(time.time())
        model = str(body.get("model") or st.model_name)
        obj_stream = "chat.completion.chunk" if chat else "text_completion"
        st.sched.submit(req)

        def chunk(delta_or_text, finish=None, usage=None, timings=None):
            if chat:
                ch = {"index": 0, "delta": delta_or_text, "finish_reason": finish}
            else:
                ch = {"index": 0, "text": delta_or_text, "finish_reason": finish}
            o = {"id": rid, "object": obj_stream, "created": created, "model": model, "choices": [ch]}
            if usage:
                o["usage"] = usage
            if timings:
                o["timings"] = timings
            return b"data: " + json.dumps(o, ensure_ascii=False).encode() + b"\n\n"

        if stream:
            head = ("HTTP/1.1 200 OK\r\nContent-Type: text/event-stream\r\nCache-Control: no-cache\r\nConnection: close\r\n"
                    "Server: ggufb200\r\nAccess-Control-Allow-Origin: *\r\n\r\n").encode()
            self.close_connection = True
            try:
                self.connection.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                first = head + (chunk({"role": "assistant", "content": None}) if chat else b"")
                self.wfile.write(first)
                self.wfile.flush()
                while True:
                    ev = req.events.get()
                    if ev[0] == "piece":
                        if ev[1]:
                            self.wfile.write(chunk({"content": ev[1]} if chat else ev[1]))
                            self.wfile.flush()
                    elif ev[0] == "done":
                        self.wfile.write(chunk({} if chat else "", finish=ev[1], usage=ev[2], timings=ev[3]) + b"data: [DONE]\n\n")
                        self.wfile.flush()
                        return
                    else:  # error after the stream began: report it in-band
                        self.wfile.write(b"data: " + json.dumps({"error": {"code": 500, "message": ev[1], "type": "server_error"}}).encode() + b"\n\n")
                        self.wfile.flush()
                        return
            except (BrokenPipeError, ConnectionResetError, OSError):
                req.cancelled.set()   # client went away: free the slot
                return
        text = []
        while True:
            ev = req.events.get()
            if ev[0] == "piece":
                text.append(ev[1])
            elif ev[0] == "done":
                _, reason, usage, timings = ev
                content = "".join(text)
                if chat:
                    choice = {"index": 0, "message": {"role": "assistant", "content": content}, "finish_reason": reason}
                    obj = "chat.completion"
                else:
                    choice = {"index": 0, "text": content, "finish_reason": reason}
                    obj = "text_completion"
                out = {"id": rid, "object": obj, "created": created, "model": model, "choices": [choice], "usage": usage, "timings": timings}
                if not chat:
                    out["content"] = content   # llama-server's native /completion field
                return self._send(200, json.dumps(out, ensure_ascii=False).encode())
            else:
                code = 400 if ("context size" in ev[1] or ev[1].startswith("invalid request")) else 500
                return self._send(*_err(code, ev[1], "invalid_request_error" if code == 400 else "server_error"))


class Server(ThreadingHTTPServer):
    daemon_threads = True
    allow_reuse_address = True
    request_queue_size = 128


def make_server(host: str, port: int, state: ServerState) -> Server:
    handler = type("BoundHandler", (Handler,), {"state": state})
    return Server((host, port), handler)
