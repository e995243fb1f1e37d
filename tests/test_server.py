"""CPU tests of the drop-in boundary: llama-server argv contract, HTTP contract, scheduler, tokenizer -- and, when
the reference checkout is mounted, the reference's UNCHANGED gateway and benchmark client in front of it.
The compute behind the boundary is a test double built on the oracle (tests/fake_engine.py); the GPU engine is
exercised through the same server in tests/test_gpu_server.py."""
import http.client
import json
import os
import socket
import subprocess
import sys
import threading
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF_DIR", "/root/reference")
KEY = "gateway-" + "A" * 43


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.fixture(scope="module")
def tiny_path(model_dir):
    from ggufb200 import synth
    p = os.path.join(model_dir, "srv-tiny.gguf")
    synth.write_gguf(p, "tiny", "Q4_K_M", seed=0xB200)
    return p


@pytest.fixture(scope="module")
def backend(oracle, tiny_path):
    from fake_engine import OracleEngine
    from ggufb200.gguf_reader import GGUFFile
    from ggufb200.scheduler import SamplingParams, Scheduler
    from ggufb200.server import ServerState, make_server
    from ggufb200.tokenizer import Tokenizer
    tok = Tokenizer(GGUFFile(tiny_path).meta)
    eng = OracleEngine(tiny_path, n_ctx=160, n_slots=2)
    sched = Scheduler(eng, tok, ignore_eos=True)
    sched.start()
    st = ServerState(sched, tok, "srv-tiny.gguf", KEY, 160, SamplingParams(temperature=0.0), log=lambda *a: None)
    st.ready.set()
    port = free_port()
    httpd = make_server("127.0.0.1", port, st)
    threading.Thread(target=httpd.serve_forever, daemon=True).start()
    yield {"port": port, "tok": tok, "state": st, "path": tiny_path}
    httpd.shutdown()
    sched.shutdown()


def call(port, method, path, body=None, key=KEY, raw=False):
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=60)
    h = {"Content-Type": "application/json"}
    if key:
        h["Authorization"] = f"Bearer {key}"
    c.request(method, path, json.dumps(body) if body is not None else None, h)
    r = c.getresponse()
    data = r.read()
    c.close()
    return (r.status, data, dict(r.getheaders())) if raw else (r.status, json.loads(data) if data else None)


def expected_text(oracle, backend, messages, n):
    tok = backend["tok"]
    ids = tok.encode_chat(messages)
    m = oracle.OracleLlama(backend["path"], n_ctx=160, mode="canon")
    return tok.decode(m.greedy(ids, n)), len(ids)


MSG = [{"role": "user", "content": "Write a short poem about the sea"}]


def test_health_is_public_single_packet_json(backend):
    s = socket.create_connection(("127.0.0.1", backend["port"]))
    s.sendall(b"GET /health HTTP/1.1\r\nHost: x\r\nConnection: close\r\n\r\n")  # no Authorization, like gateway.py:336-340
    data = s.recv(4096)                                                            # ONE read, like gateway.py:344
    s.close()
    head, _, body = data.partition(b"\r\n\r\n")
    assert head.startswith(b"HTTP/1.1 200")
    assert b"chunked" not in head.lower()
    assert json.loads(body)["status"] == "ok" and len(data) < 4096


def test_key_is_enforced_except_on_public_paths(backend):
    p = backend["port"]
    assert call(p, "POST", "/v1/chat/completions", {"messages": MSG, "max_tokens": 2}, key=None)[0] == 401
    assert call(p, "POST", "/v1/chat/completions", {"messages": MSG, "max_tokens": 2}, key="gateway-" + "B" * 43)[0] == 401
    st, body = call(p, "GET", "/v1/models", key=None)
    assert st == 200 and body["object"] == "list" and body["data"][0]["object"] == "model"
    st, body = call(p, "POST", "/v1/chat/completions", {"messages": MSG, "max_tokens": 2}, key=None)
    assert body["error"]["type"] == "authentication_error"


def test_chat_completion_matches_oracle_tokens(oracle, backend):
    st, body = call(backend["port"], "POST", "/v1/chat/completions", {"model": "default", "messages": MSG, "max_tokens": 12, "temperature": 0})
    assert st == 200 and body["object"] == "chat.completion"
    text, n_prompt = expected_text(oracle, backend, MSG, 12)
    ch = body["choices"][0]
    assert ch["message"] == {"role": "assistant", "content": text}
    assert ch["finish_reason"] == "length"
    assert body["usage"] == {"prompt_tokens": n_prompt, "completion_tokens": 12, "total_tokens": n_prompt + 12}
    # the synthetic vocabulary gives one leading-space word per token (benchmark.py counts whitespace words)
    assert len(text.split()) == 12


def test_streaming_sse_framing_and_close(oracle, backend):
    s = socket.create_connection(("127.0.0.1", backend["port"]))
    body = json.dumps({"model": "default", "messages": MSG, "max_tokens": 8, "stream": True, "temperature": 0}).encode()
    s.sendall(b"POST /v1/chat/completions HTTP/1.1\r\nhost: x\r\nauthorization: Bearer " + KEY.encode() +
              b"\r\ncontent-type: application/json\r\ncontent-length: " + str(len(body)).encode() + b"\r\nConnection: close\r\n\r\n" + body)
    buf = b""
    while True:
        d = s.recv(8192)
        if not d:
            break   # the backend closes the socket when the response is complete (gateway.py:777-783 relays until EOF)
        buf += d
    s.close()
    head, _, payload = buf.partition(b"\r\n\r\n")
    assert b"text/event-stream" in head and b"chunked" not in head.lower()
    events = [l[5:].strip() for l in payload.decode().split("\n") if l.startswith("data:")]
    assert events[-1] == "[DONE]"
    chunks = [json.loads(e) for e in events[:-1]]
    assert all(c["object"] == "chat.completion.chunk" for c in chunks)
    assert chunks[0]["choices"][0]["delta"].get("role") == "assistant"
    text = "".join(c["choices"][0]["delta"].get("content") or "" for c in chunks)
    assert text == expected_text(oracle, backend, MSG, 8)[0]
    assert chunks[-1]["choices"][0]["finish_reason"] == "length"
    assert all(c["choices"][0]["finish_reason"] is None for c in chunks[:-1])
    assert chunks[-1]["usage"]["completion_tokens"] == 8


def test_completions_endpoint_stop_and_errors(oracle, backend):
    p = backend["port"]
    tok = backend["tok"]
    m = oracle.OracleLlama(backend["path"], n_ctx=160, mode="canon")
    ids = tok.encode("hello world")
    full = tok.decode(m.greedy(ids, 10))
    st, body = call(p, "POST", "/v1/completions", {"prompt": "hello world", "max_tokens": 10, "temperature": 0})
    assert st == 200 and body["choices"][0]["text"] == full
    third = full.split()[2]
    st, body = call(p, "POST", "/v1/completions", {"prompt": "hello world", "max_tokens": 10, "temperature": 0, "stop": [third]})
    assert body["choices"][0]["finish_reason"] == "stop" and body["choices"][0]["text"] == full[:full.index(third)]
    assert call(p, "POST", "/v1/chat/completions", {"max_tokens": 4})[0] == 400
    assert call(p, "POST", "/v1/chat/completions", {"messages": [{"role": "user", "content": "x" * 4000}], "max_tokens": 4})[0] == 400  # context
    assert call(p, "GET", "/nope")[0] == 404
    c = http.client.HTTPConnection("127.0.0.1", p, timeout=10)
    c.request("POST", "/v1/chat/completions", "{not json", {"Authorization": f"Bearer {KEY}"})
    assert c.getresponse().status == 400


def test_malformed_requests_get_400_and_leave_the_backend_healthy(backend):
    """raw token ids outside the vocabulary, non-numeric fields, null fields, a negative Content-Length: each is the
    client's error (400) -- none may reach the engine or kill the scheduler"""
    p = backend["port"]
    nv = backend["tok"].n_vocab
    for bad in ([1, nv], [1, -5], [1, 10 ** 12]):
        st, body = call(p, "POST", "/v1/completions", {"prompt": bad, "max_tokens": 2})
        assert st == 400 and "vocabulary" in body["error"]["message"]
    assert call(p, "POST", "/v1/completions", {"prompt": [1, 300], "max_tokens": [1]})[0] == 400
    assert call(p, "POST", "/v1/completions", {"prompt": [1, 300], "max_tokens": 2, "temperature": "hot"})[0] == 400
    st, body = call(p, "POST", "/v1/completions", {"prompt": [1, 300], "max_tokens": 2, "temperature": None, "top_p": None, "seed": None})
    assert st == 200 and body["usage"]["completion_tokens"] == 2          # JSON null = default
    s = socket.create_connection(("127.0.0.1", p), timeout=10)
    s.sendall(b"POST /v1/completions HTTP/1.1\r\nHost: x\r\nAuthorization: Bearer " + KEY.encode() + b"\r\nContent-Length: -5\r\n\r\n")
    assert b" 400 " in s.recv(4096).split(b"\r\n")[0]
    s.close()
    assert backend["state"].sched.is_alive() and backend["state"].sched.fatal is None
    assert call(p, "POST", "/v1/completions", {"prompt": [1, 300], "max_tokens": 2})[0] == 200


def test_sampling_is_seeded_and_differs_from_greedy(backend):
    p = backend["port"]
    req = {"messages": MSG, "max_tokens": 12, "temperature": 1.5, "top_k": 50, "seed": 7}
    a = call(p, "POST", "/v1/chat/completions", req)[1]["choices"][0]["message"]["content"]
    b = call(p, "POST", "/v1/chat/completions", req)[1]["choices"][0]["message"]["content"]
    g = call(p, "POST", "/v1/chat/completions", {**req, "temperature": 0})[1]["choices"][0]["message"]["content"]
    assert a == b and a != g


def test_concurrent_requests_and_health_during_generation(backend):
    p = backend["port"]
    out, errs = [None] * 5, []

    def worker(i):
        try:
            out[i] = call(p, "POST", "/v1/chat/completions", {"messages": MSG, "max_tokens": 24, "temperature": 0})
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(5)]
    [t.start() for t in ts]
    t0 = time.time()
    st, h = call(p, "GET", "/health", key=None)      # must not queue behind the generations
    assert st == 200 and time.time() - t0 < 1.0
    [t.join() for t in ts]
    assert not errs and all(o[0] == 200 for o in out)
    assert len({o[1]["choices"][0]["message"]["content"] for o in out}) == 1   # same prompt, greedy -> same text on every slot
    # two busy slots advance together: the scheduler batched their decode steps, and -- all greedy -- ran them one step
    # ahead of the host (step k+1 launched from device state before step k's tokens were read back)
    assert backend["state"].sched.stats.get("batched_steps", 0) > 0
    assert backend["state"].sched.stats.get("chained_steps", 0) > 0


def test_pipelined_batch_survives_sequences_ending_in_flight(oracle, backend):
    """greedy requests of different lengths, one cut by a stop string: a sequence that ends at step k has already been carried
    into the speculatively launched step k+1 -- its extra token must vanish and the others must not notice"""
    p = backend["port"]
    want40, _ = expected_text(oracle, backend, MSG, 40)
    want9, _ = expected_text(oracle, backend, MSG, 9)
    stop = want40[len(want40) // 2:len(want40) // 2 + 3]
    res = {}

    def worker(name, body):
        res[name] = call(p, "POST", "/v1/chat/completions", body)

    ts = [threading.Thread(target=worker, args=("long", {"messages": MSG, "max_tokens": 40, "temperature": 0})),
          threading.Thread(target=worker, args=("short", {"messages": MSG, "max_tokens": 9, "temperature": 0})),
          threading.Thread(target=worker, args=("stop", {"messages": MSG, "max_tokens": 40, "temperature": 0, "stop": [stop]}))]
    before = backend["state"].sched.stats.get("chained_steps", 0)
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert all(r[0] == 200 for r in res.values())
    assert res["long"][1]["choices"][0]["message"]["content"] == want40
    assert res["short"][1]["choices"][0]["message"]["content"] == want9
    assert res["stop"][1]["choices"][0]["message"]["content"] == want40[:want40.find(stop)]
    assert res["stop"][1]["choices"][0]["finish_reason"] == "stop"
    assert backend["state"].sched.stats.get("chained_steps", 0) > before
    # the slots' prompt caches are intact: the same request again gives the same text
    assert call(p, "POST", "/v1/chat/completions", {"messages": MSG, "max_tokens": 40, "temperature": 0})[1]["choices"][0]["message"]["content"] == want40


def test_batched_and_single_steps_interleave(oracle, backend):
    """a long request keeps running alone after a short one that shared batches with it has finished (the slot goes
    batch -> feed -> device-side chain), and a sampled request rides in the same batches as a greedy one"""
    p = backend["port"]
    want, _ = expected_text(oracle, backend, MSG, 40)
    res = {}

    def worker(name, body):
        res[name] = call(p, "POST", "/v1/chat/completions", body)

    ts = [threading.Thread(target=worker, args=("long", {"messages": MSG, "max_tokens": 40, "temperature": 0})),
          threading.Thread(target=worker, args=("short", {"messages": MSG, "max_tokens": 6, "temperature": 1.2, "seed": 3}))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert res["long"][0] == 200 and res["short"][0] == 200
    assert res["long"][1]["choices"][0]["message"]["content"] == want
    assert res["short"][1]["usage"]["completion_tokens"] == 6


def test_embeddings_endpoint(oracle, backend):
    """POST /v1/embeddings (docs/API_REFERENCE.md:540-590 of the reference): 501 with upstream's message unless the server was
    started with --embeddings; then one pooled, L2-normalised vector per input, OpenAI list shape, usage counted."""
    p, st = backend["port"], backend["state"]
    code, body = call(p, "POST", "/v1/embeddings", {"model": "any", "input": "Hello world"})
    assert code == 501 and "--embeddings" in body["error"]["message"]
    st.embeddings = "mean"
    try:
        code, body = call(p, "POST", "/v1/embeddings", {"model": "any", "input": ["Hello world", "The sea"]})
        assert code == 200 and body["object"] == "list" and [d["index"] for d in body["data"]] == [0, 1]
        tok = backend["tok"]
        m = oracle.OracleLlama(backend["path"], n_ctx=160, mode="canon")
        for d, text in zip(body["data"], ["Hello world", "The sea"]):
            ids = tok.encode(text, add_special=True)
            hs = [np.asarray(m.forward(t, i, return_hidden=True), dtype=np.float64) for i, t in enumerate(ids)]
            want = np.mean(hs, axis=0)
            want /= np.linalg.norm(want)
            got = np.asarray(d["embedding"])
            assert d["object"] == "embedding" and got.shape == want.shape
            assert abs(np.linalg.norm(got) - 1.0) < 1e-5 and np.abs(got - want).max() < 1e-6
        n_tok = sum(len(tok.encode(t, add_special=True)) for t in ["Hello world", "The sea"])
        assert body["usage"] == {"prompt_tokens": n_tok, "total_tokens": n_tok}
        assert call(p, "POST", "/v1/embeddings", {"input": []})[0] == 400
        assert call(p, "POST", "/v1/embeddings", {"input": [[1, 10 ** 9]]})[0] == 400
        # a completion still works on the slot the embedding used (its prompt cache was dropped, not corrupted)
        code, body = call(p, "POST", "/v1/chat/completions", {"messages": MSG, "max_tokens": 6, "temperature": 0})
        assert code == 200 and body["choices"][0]["message"]["content"] == expected_text(oracle, backend, MSG, 6)[0]
    finally:
        st.embeddings = ""


def test_cli_version_and_argv_contract():
    exe = os.path.join(ROOT, "bin", "llama-server")
    r = subprocess.run([exe, "--version"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.strip().startswith("version:")
    from ggufb200.cli import build_parser
    # the exact argv scripts/start.sh:473-494 builds, plus pass-through extras
    argv = ["-m", "/data/models/m.gguf", "--host", "127.0.0.1", "--port", "8080", "-c", "16384", "-ngl", "99",
            "--api-key-file", "/dev/shm/llama-keys/backend-1.key", "-t", "8", "--flash-attn", "--parallel", "16"]
    a, unknown = build_parser().parse_known_args(argv)
    assert (a.model, a.host, a.port, a.ctx_size, a.ngl, a.parallel) == ("/data/models/m.gguf", "127.0.0.1", 8080, 16384, "99", 16)
    assert unknown == ["--flash-attn"]
    r = subprocess.run([exe, "-m", "/nonexistent.gguf", "--port", "1"], capture_output=True, text=True, timeout=60)
    assert r.returncode != 0 and "not found" in r.stderr


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "scripts", "gateway.py")), reason="reference checkout not mounted (set REF_DIR)")
def test_unchanged_reference_gateway_and_benchmark_in_front(oracle, backend):
    """The reference's own gateway.py and benchmark.py, executed from where they are mounted, against this backend."""
    gport = free_port()
    env = {**os.environ, "GATEWAY_PORT": str(gport), "PORT_BACKEND": str(backend["port"]), "BACKEND_HOST": "127.0.0.1",
           "BACKEND_API_KEY": KEY, "AUTH_ENABLED": "false", "MAX_CONCURRENT_REQUESTS": "2", "DATA_DIR": "/tmp"}
    gw = subprocess.Popen([sys.executable, os.path.join(REF, "scripts", "gateway.py")], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    try:
        for _ in range(100):
            try:
                if call(gport, "GET", "/ping", key=None, raw=True)[0] == 200:
                    break
            except OSError:
                time.sleep(0.1)
        st, h = call(gport, "GET", "/health", key=None)
        assert st == 200 and h["backend"]["status"] == "ok"          # gateway surfaces our JSON (gateway.py:356-363)
        st, body = call(gport, "POST", "/v1/chat/completions", {"model": "default", "messages": MSG, "max_tokens": 10, "temperature": 0}, key=None)
        assert st == 200 and body["choices"][0]["message"]["content"] == expected_text(oracle, backend, MSG, 10)[0]
        # the reference's benchmark client (streaming, whitespace token count)
        r = subprocess.run([sys.executable, os.path.join(REF, "scripts", "benchmark.py"), "--url", f"http://127.0.0.1:{gport}",
                            "--requests", "2", "--warmup", "0", "--max-tokens", "12", "--concurrency", "2", "--output", "json"],
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        rep = json.loads(r.stdout)
        inf = rep["inference"]
        assert inf["requests_success"] == 2 and inf["requests_failed"] == 0
        assert inf["tokens_per_sec"]["count"] == 2 and inf["ttft"]["count"] == 2   # it saw content deltas and counted words
    finally:
        gw.terminate()
        gw.wait(timeout=10)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "scripts", "start.sh")), reason="reference checkout not mounted (set REF_DIR)")
def test_reference_start_sh_runs_the_whole_stack(oracle, tiny_path, tmp_path):
    """The reference's launcher, scripts/start.sh, run by bash from where it is mounted.  Only its two container paths are
    redirected while it is piped to bash (/app/llama-server -> a stand-in that runs the product's cli.main with the oracle
    engine; /opt/app/scripts -> the mounted scripts directory): version probe (start.sh:359-365), argv (473-494), the
    `| tee` log pipe (516-517), the authenticated /health poll and the no-key probe (600-646), health server, gateway,
    a chat completion through the gateway with a key from AUTH_KEYS_FILE, and the SIGTERM shutdown (402-430)."""
    import shutil
    import signal
    if not shutil.which("bash") or not shutil.which("curl"):
        pytest.skip("needs bash and curl")
    app = tmp_path / "app"
    app.mkdir()
    fake = app / "llama-server"
    fake.write_text(f"#!/bin/sh\nexec {sys.executable} {os.path.join(ROOT, 'tests', 'fake_llama_server.py')} \"$@\"\n")
    fake.chmod(0o755)
    data = tmp_path / "data"
    (data / "models").mkdir(parents=True)
    os.symlink(tiny_path, data / "models" / "tiny.gguf")
    user_key = "sk-test-" + "k" * 40
    (data / "api_keys.txt").write_text(f"tester:{user_key}\n")
    script = open(os.path.join(REF, "scripts", "start.sh")).read()
    assert "/app/llama-server" in script and "/opt/app/scripts" in script
    script = script.replace("/app/llama-server", str(fake)).replace("/opt/app/scripts", os.path.join(REF, "scripts"))
    gport, bport, hport = free_port(), free_port(), free_port()
    env = {**os.environ, "DATA_DIR": str(data), "MODEL_NAME": "tiny.gguf", "PORT": str(gport), "PORT_BACKEND": str(bport),
           "PORT_HEALTH": str(hport), "NGL": "0", "CTX": "160", "THREADS": "2", "EXTRA_ARGS": "--parallel 2 --temp 0 --ignore-eos",
           "AUTH_ENABLED": "true", "INSTANCE_ID": f"pytest-{os.getpid()}", "MAX_CONCURRENT_REQUESTS": "2"}
    proc = subprocess.Popen(["bash", "-s"], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            env=env, start_new_session=True)
    proc.stdin.write(script)
    proc.stdin.close()
    lines = []
    threading.Thread(target=lambda: lines.extend(proc.stdout), daemon=True).start()
    try:
        ok = False
        for _ in range(600):
            assert proc.poll() is None, "".join(lines)[-4000:]
            try:
                if call(gport, "GET", "/ping", key=None, raw=True)[0] == 200 and any("Services running" in ln for ln in lines):
                    ok = True
                    break
            except OSError:
                pass
            time.sleep(0.1)
        assert ok, "".join(lines)[-4000:]
        log = "".join(lines)
        assert "llama-server: version:" in log                                   # start.sh:364-365 printed OUR --version line
        assert "Backend responds correctly" in log                               # start.sh:600-635, authenticated /health poll
        # start.sh:637-646 then probes /health WITHOUT a key: upstream's llama-server keeps /health public (and gateway.py:336-344
        # relies on that), so the launcher prints its warning -- for the real binary and for this one alike -- and carries on
        assert "Backend responded without authentication" in log
        st, h = call(gport, "GET", "/health", key=None)
        assert st == 200 and h["backend"]["status"] == "ok"
        body = {"model": "default", "messages": MSG, "max_tokens": 10, "temperature": 0}
        assert call(gport, "POST", "/v1/chat/completions", body, key=None)[0] == 401           # the gateway's own auth
        st, r = call(gport, "POST", "/v1/chat/completions", body, key=user_key)
        tok_path = {"tok": __import__("ggufb200.tokenizer", fromlist=["Tokenizer"]).Tokenizer(
            __import__("ggufb200.gguf_reader", fromlist=["GGUFFile"]).GGUFFile(tiny_path).meta), "path": tiny_path}
        assert st == 200 and r["choices"][0]["message"]["content"] == expected_text(oracle, tok_path, MSG, 10)[0]
        logs = list((data / "logs" / "llama").glob("*_server_*.log"))           # the tee'd backend log (start.sh:505-517)
        assert logs and "server is listening" in logs[0].read_text()
        os.killpg(proc.pid, signal.SIGTERM)
        assert proc.wait(timeout=60) is not None
        time.sleep(0.5)
        with pytest.raises(OSError):                                             # the backend is gone with its launcher
            call(bport, "GET", "/health", key=None)
    finally:
        if proc.poll() is None:
            os.killpg(proc.pid, signal.SIGKILL)


def test_prompt_cache_reuses_the_common_prefix(oracle, backend):
    """a follow-up request whose prompt starts like the previous one only processes the new tokens (llama-server's
    cache_prompt); the text must be what a cold run of the full prompt gives"""
    p = backend["port"]
    sched = backend["state"].sched
    first = [{"role": "user", "content": "Hi"}]            # the test backend has a 160-token context
    st, r1 = call(p, "POST", "/v1/chat/completions", {"messages": first, "max_tokens": 4, "temperature": 0})
    assert st == 200
    follow = first + [{"role": "assistant", "content": r1["choices"][0]["message"]["content"]}, {"role": "user", "content": "Go"}]
    before = sched.stats.get("prompt_tokens_cached", 0)
    st, r2 = call(p, "POST", "/v1/chat/completions", {"messages": follow, "max_tokens": 10, "temperature": 0})
    assert st == 200, r2
    assert sched.stats.get("prompt_tokens_cached", 0) - before >= 16
    want, n_prompt = expected_text(oracle, backend, follow, 10)
    assert r2["choices"][0]["message"]["content"] == want and r2["usage"]["prompt_tokens"] == n_prompt
    before = sched.stats.get("prompt_tokens_cached", 0)
    r3 = call(p, "POST", "/v1/chat/completions", {"messages": follow, "max_tokens": 10, "temperature": 0, "cache_prompt": False})[1]
    assert sched.stats.get("prompt_tokens_cached", 0) == before
    assert r3["choices"][0]["message"]["content"] == want


def test_sampled_requests_use_device_candidates_and_give_the_whole_row_tokens(oracle, backend, monkeypatch):
    """the scheduler's candidate path (top-k candidates + penalty-window logits instead of whole rows; sampled batches pipelined):
    the same seeded requests -- plain chain, penalised, greedy-with-penalties, alone and two at a time -- return the texts the
    whole-row path returns (GGB_DEVICE_TOPK=0), and the candidate path is what ran"""
    p, st = backend["port"], backend["state"]
    bodies = [{"messages": MSG, "max_tokens": 14, "temperature": 0.9, "top_k": 30, "top_p": 0.9, "seed": 11},
              {"messages": MSG, "max_tokens": 9, "temperature": 0.7, "top_k": 20, "repeat_penalty": 1.3, "presence_penalty": 0.3, "seed": 12},
              {"messages": MSG, "max_tokens": 11, "temperature": 0.0, "frequency_penalty": 0.5, "repeat_penalty": 1.2}]

    def run_all():
        out = [call(p, "POST", "/v1/chat/completions", b) for b in bodies]                  # one at a time
        res = {}
        ts = [threading.Thread(target=lambda j=j: res.__setitem__(j, call(p, "POST", "/v1/chat/completions", bodies[j]))) for j in (0, 1)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        return out + [res[0], res[1]]

    before = st.sched.stats.get("device_topk_tokens", 0)
    fast = run_all()
    used = st.sched.stats.get("device_topk_tokens", 0) - before
    monkeypatch.setenv("GGB_DEVICE_TOPK", "0")
    slow = run_all()
    assert st.sched.stats.get("device_topk_tokens", 0) - before == used                    # the whole-row path did not count
    for a, b in zip(fast, slow):
        assert a[0] == 200 and b[0] == 200
        assert a[1]["choices"][0]["message"]["content"] == b[1]["choices"][0]["message"]["content"]
    assert fast[0][1]["choices"][0]["message"]["content"] == fast[3][1]["choices"][0]["message"]["content"]   # alone == in a batch
    assert used >= 14 + 9 + 11 + 14 + 9 - 5
