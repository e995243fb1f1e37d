"""-m gpu: the real `bin/llama-server` process (GPU engine) behind its HTTP contract, launched with the argv
scripts/start.sh:473-494 builds; output checked against the CPU oracle."""
import http.client
import json
import os
import signal
import socket
import subprocess
import time

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = "gateway-" + "Z" * 43
MSG = [{"role": "user", "content": "Write a short poem about the sea"}]


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def call(port, method, path, body=None, key=KEY):
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=120)
    h = {"Content-Type": "application/json"}
    if key:
        h["Authorization"] = f"Bearer {key}"
    c.request(method, path, json.dumps(body) if body is not None else None, h)
    r = c.getresponse()
    data = r.read()
    c.close()
    return r.status, (json.loads(data) if data else None)


def test_llama_server_process_contract_and_parity(oracle, model_dir, tmp_path):
    from ggufb200 import synth
    from ggufb200.gguf_reader import GGUFFile
    from ggufb200.tokenizer import Tokenizer
    path = os.path.join(model_dir, "proc-small.gguf")
    synth.write_gguf(path, "small", "Q4_K_M", seed=0xB200)
    keyfile = tmp_path / "backend.key"
    keyfile.write_text(KEY + "\n")
    port = free_port()
    argv = [os.path.join(ROOT, "bin", "llama-server"), "-m", path, "--host", "127.0.0.1", "--port", str(port), "-c", "256",
            "-ngl", "99", "--api-key-file", str(keyfile), "-t", "4", "--parallel", "2", "--temp", "0", "--ignore-eos"]
    # parity run: keep the prompt on the bit-exact integer path (prompts >= 64 tokens otherwise take the bf16 tensor-core
    # prefill, whose logits agree with the reference only to tolerance -- tests/test_gpu_engine.py covers that path)
    env = dict(os.environ, GGB_GEMM_PREFILL_MIN="1000000")
    proc = subprocess.Popen(argv, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    try:
        ok = False
        for _ in range(120):                    # start.sh polls /health for ~30 s; 503 while loading, then 200
            try:
                st, body = call(port, "GET", "/health", key=None)
                if st == 200 and body["status"] == "ok":
                    ok = True
                    break
            except OSError:
                pass
            assert proc.poll() is None, proc.stdout.read()
            time.sleep(0.25)
        assert ok
        assert call(port, "POST", "/v1/chat/completions", {"messages": MSG, "max_tokens": 4}, key=None)[0] == 401
        st, body = call(port, "POST", "/v1/chat/completions", {"model": "default", "messages": MSG, "max_tokens": 48})
        assert st == 200
        tok = Tokenizer(GGUFFile(path).meta)
        ids = tok.encode_chat(MSG)
        ref = oracle.OracleLlama(path, n_ctx=256, mode="canon")
        assert body["choices"][0]["message"]["content"] == tok.decode(ref.greedy(ids, 48))
        assert body["usage"]["completion_tokens"] == 48 and body["timings"]["predicted_per_second"] > 0
        proc.send_signal(signal.SIGTERM)
        assert proc.wait(timeout=30) == 0       # start.sh gives 30 s before SIGKILL (start.sh:414-428)
    finally:
        if proc.poll() is None:
            proc.kill()


def test_tensor_parallel_llama_server_under_torchrun(oracle, model_dir, tmp_path):
    """`torchrun --nproc-per-node 2 --no-python bin/llama-server ...`: rank 0 serves HTTP, rank 1 mirrors its engine calls
    (tp_serve.py); greedy text equals the single-process oracle, a sampled request (which gathers the vocabulary-sharded
    logits from both ranks) is served too."""
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from dataclasses import replace
    from ggufb200 import synth
    from ggufb200.gguf_reader import GGUFFile
    from ggufb200.tokenizer import Tokenizer
    cfg = replace(synth.PRESETS["medium"], n_layer=4, ff=3072)
    path = os.path.join(model_dir, "tp-serve-medium.gguf")
    synth.write_gguf(path, cfg, "Q4_K_M", seed=0xB200)
    keyfile = tmp_path / "backend.key"
    keyfile.write_text(KEY + "\n")
    port = free_port()
    argv = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
            "--master-port", str(free_port()), "--no-python", os.path.join(ROOT, "bin", "llama-server"), "-m", path, "--host", "127.0.0.1",
            "--port", str(port), "-c", "256", "--api-key-file", str(keyfile), "--parallel", "2", "--temp", "0", "--ignore-eos"]
    env = dict(os.environ, GGUFB200_PYTHON=sys.executable)
    proc = subprocess.Popen(argv, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, start_new_session=True)
    try:
        ok = False
        for _ in range(480):
            try:
                st, body = call(port, "GET", "/health", key=None)
                if st == 200 and body["status"] == "ok":
                    ok = True
                    break
            except OSError:
                pass
            assert proc.poll() is None, proc.stdout.read()
            time.sleep(0.25)
        assert ok
        st, body = call(port, "POST", "/v1/chat/completions", {"model": "default", "messages": MSG, "max_tokens": 24})
        assert st == 200
        tok = Tokenizer(GGUFFile(path).meta)
        ids = tok.encode_chat(MSG)
        ref = oracle.OracleLlama(path, n_ctx=256, mode="canon")
        assert body["choices"][0]["message"]["content"] == tok.decode(ref.greedy(ids, 24))
        st, body = call(port, "POST", "/v1/chat/completions", {"messages": MSG, "max_tokens": 8, "temperature": 1.0, "seed": 5})
        assert st == 200 and body["usage"]["completion_tokens"] == 8
        # concurrent requests share ONE pass over the sharded weights per step (TPBatch: launches mirrored on rank 1, f64
        # all-reduce per row-split projection), pipelined one step ahead when all are greedy; a sampled one rides along
        import threading
        res = {}

        def worker(name, req):
            res[name] = call(port, "POST", "/v1/chat/completions", req)

        ts = [threading.Thread(target=worker, args=("a", {"messages": MSG, "max_tokens": 24})),
              threading.Thread(target=worker, args=("b", {"messages": MSG, "max_tokens": 16})),
              threading.Thread(target=worker, args=("c", {"messages": MSG, "max_tokens": 8, "temperature": 1.0, "seed": 5}))]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert all(r[0] == 200 for r in res.values()), res
        assert res["a"][1]["choices"][0]["message"]["content"] == tok.decode(ref.greedy(ids, 24))
        assert res["b"][1]["choices"][0]["message"]["content"] == tok.decode(ref.greedy(ids, 16))
        assert res["c"][1]["usage"]["completion_tokens"] == 8
        c = http.client.HTTPConnection("127.0.0.1", port, timeout=30)
        c.request("GET", "/metrics", headers={"Authorization": f"Bearer {KEY}"})
        metrics = c.getresponse().read().decode()
        c.close()
        batched = [float(ln.split()[1]) for ln in metrics.splitlines() if ln.startswith("ggufb200:batched_steps_total")]
        assert batched and batched[0] > 0, metrics
    finally:
        if proc.poll() is None:
            os.killpg(proc.pid, signal.SIGTERM)     # the process group WE started (torchrun + both ranks)
            try:
                proc.wait(timeout=60)
            except subprocess.TimeoutExpired:
                os.killpg(proc.pid, signal.SIGKILL)


def test_concurrent_sampled_requests_are_seeded_pipelined_and_sampled_from_device_candidates(model_dir, tmp_path):
    """four concurrent requests with llama-server's default chain (temp 0.8, top-k 40, top-p 0.95), different lengths so that
    sequences end while a pipelined step is in flight: every request returns its own max_tokens, the same seeds give the same
    texts on a second round (batched logits are bit-identical whatever the batch composition), and the tokens were sampled from
    the device's top-k candidates (/metrics)"""
    import threading
    from ggufb200 import synth
    path = os.path.join(model_dir, "proc-small.gguf")
    if not os.path.exists(path):
        synth.write_gguf(path, "small", "Q4_K_M", seed=0xB200)
    keyfile = tmp_path / "backend.key"
    keyfile.write_text(KEY + "\n")
    port = free_port()
    argv = [os.path.join(ROOT, "bin", "llama-server"), "-m", path, "--host", "127.0.0.1", "--port", str(port), "-c", "256",
            "--api-key-file", str(keyfile), "--parallel", "4", "--ignore-eos"]
    proc = subprocess.Popen(argv, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, GGB_GEMM_PREFILL_MIN="1000000"))
    try:
        for _ in range(120):
            try:
                if call(port, "GET", "/health", key=None)[0] == 200:
                    break
            except OSError:
                pass
            assert proc.poll() is None, proc.stdout.read()
            time.sleep(0.25)
        lens = [9, 23, 40, 31]

        def round_():
            out = [None] * 4

            def one(j):
                msg = [{"role": "user", "content": f"Tell me about topic number {j}"}]
                out[j] = call(port, "POST", "/v1/chat/completions", {"messages": msg, "max_tokens": lens[j], "seed": 100 + j, "cache_prompt": False})
            th = [threading.Thread(target=one, args=(j,)) for j in range(4)]
            [t.start() for t in th]
            [t.join() for t in th]
            return out
        a, b = round_(), round_()
        for j in range(4):
            assert a[j][0] == 200 and a[j][1]["usage"]["completion_tokens"] == lens[j]
            assert a[j][1]["choices"][0]["message"]["content"] == b[j][1]["choices"][0]["message"]["content"], j
        assert len({a[j][1]["choices"][0]["message"]["content"][:20] for j in range(4)}) > 1      # sampled, different prompts / seeds
        c = http.client.HTTPConnection("127.0.0.1", port, timeout=30)
        c.request("GET", "/metrics", headers={"Authorization": f"Bearer {KEY}"})
        metrics = c.getresponse().read().decode()
        c.close()
        topk = [float(l.split()[1]) for l in metrics.splitlines() if l.startswith("ggufb200:device_topk_tokens_total")][0]
        assert topk >= 2 * sum(lens) - 8, metrics          # (the first token of a request comes from its prompt pass)
        proc.send_signal(signal.SIGTERM)
        assert proc.wait(timeout=30) == 0
    finally:
        if proc.poll() is None:
            proc.kill()
