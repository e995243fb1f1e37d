"""BASELINE.json config 5: N concurrent streaming /v1/chat/completions requests against the real `bin/llama-server`
process (what the reference's scripts/benchmark.py --concurrent N does through the gateway; the gateway itself is a
pure byte-forwarding proxy in front of this port).  Reports aggregate completion tokens/s, per-request tok/s, TTFT.

usage: python tools/serve_bench.py [--model llama3-8b] [--ftype Q4_K_M] [--concurrent 1,16] [--max-tokens 128]"""
import argparse
import http.client
import json
import os
import signal
import socket
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KEY = "gateway-" + "B" * 43


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def stream_one(port, idx, max_tokens, out, rep=0):
    # the round number comes first: no two requests share a prefix long enough for the server's prompt cache
    body = {"model": "default", "stream": True, "max_tokens": max_tokens, "temperature": 0,
            "messages": [{"role": "user", "content": f"{rep}/{idx}: write a short story about the number {idx}, round {rep}."}]}
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=600)
    t0 = time.perf_counter()
    c.request("POST", "/v1/chat/completions", json.dumps(body), {"Content-Type": "application/json", "Authorization": f"Bearer {KEY}"})
    r = c.getresponse()
    n, t_first, usage = 0, None, None
    for line in r:
        if not line.startswith(b"data: "):
            continue
        data = line[6:].strip()
        if data == b"[DONE]":
            break
        ev = json.loads(data)
        if ev.get("usage"):
            usage = ev["usage"]
        ch = ev.get("choices") or []
        if ch and ch[0].get("delta", {}).get("content"):
            if t_first is None:
                t_first = time.perf_counter()
            n += 1
    t1 = time.perf_counter()
    c.close()
    out[idx] = {"status": r.status, "chunks": n, "completion_tokens": (usage or {}).get("completion_tokens", n),
                "ttft_s": (t_first or t1) - t0, "total_s": t1 - t0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--ftype", default="Q4_K_M")
    ap.add_argument("--concurrent", default="1,16")
    ap.add_argument("--max-tokens", type=int, default=128)
    ap.add_argument("--ctx", type=int, default=1024)
    args = ap.parse_args()
    path = bench.model_path(args.model, args.ftype, 0xB200)
    levels = [int(x) for x in args.concurrent.split(",")]
    port = free_port()
    keyfile = f"/tmp/serve_bench_{os.getpid()}.key"
    open(keyfile, "w").write(KEY + "\n")
    argv = [os.path.join(ROOT, "bin", "llama-server"), "-m", path, "--host", "127.0.0.1", "--port", str(port), "-c", str(args.ctx),
            "-ngl", "99", "--api-key-file", keyfile, "--parallel", str(max(levels)), "--temp", "0", "--ignore-eos"]
    proc = subprocess.Popen(argv, stdout=subprocess.DEVNULL, stderr=subprocess.STDOUT)
    res = {"model": args.model, "ftype": args.ftype, "max_tokens": args.max_tokens, "levels": []}
    try:
        for _ in range(1200):
            try:
                c = http.client.HTTPConnection("127.0.0.1", port, timeout=5)
                c.request("GET", "/health")
                if c.getresponse().status == 200:
                    break
            except OSError:
                pass
            if proc.poll() is not None:
                raise SystemExit("server exited")
            time.sleep(0.25)
        for n in levels:
            for rep in range(2):      # first round warms up (graph capture for this batch size); the second is reported
                out = {}
                ts = [threading.Thread(target=stream_one, args=(port, i, args.max_tokens, out, rep)) for i in range(n)]
                t0 = time.perf_counter()
                [t.start() for t in ts]
                [t.join() for t in ts]
                wall = time.perf_counter() - t0
            toks = sum(o["completion_tokens"] for o in out.values())
            row = {"concurrent": n, "ok": sum(o["status"] == 200 for o in out.values()), "completion_tokens": toks, "wall_s": wall,
                   "aggregate_tok_s": toks / wall, "per_request_tok_s": sum(o["completion_tokens"] / o["total_s"] for o in out.values()) / n,
                   "ttft_s_mean": sum(o["ttft_s"] for o in out.values()) / n, "ttft_s_max": max(o["ttft_s"] for o in out.values())}
            res["levels"].append(row)
            print(json.dumps(row), flush=True)
    finally:
        proc.send_signal(signal.SIGTERM)
        try:
            proc.wait(timeout=30)
        except subprocess.TimeoutExpired:
            proc.kill()
        os.remove(keyfile)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"serve_bench_{args.model}_{args.ftype}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
