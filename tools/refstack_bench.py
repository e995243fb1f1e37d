"""The reference's own stack, unchanged, in front of the GPU engine (BASELINE.json configs 1 and 5).

  scripts/start.sh (bash, as is; only its two container paths are redirected while it is piped to bash:
      /app/llama-server -> <repo>/bin/llama-server, /opt/app/scripts -> $REF_DIR/scripts)
    -> bin/llama-server  (this repo: GPU engine, argv built by start.sh:473-494)
    -> scripts/health_server.py, scripts/gateway.py  (reference, unchanged; key auth on)
  scripts/benchmark.py --url <gateway>  (reference, unchanged; streaming /v1/chat/completions)

$REF_DIR defaults to <repo>/baseline/_ref (git-ignored; `tools/refstack_bench.py --stage` copies /root/reference/scripts there
so that it travels to the GPU box, where /root/reference does not exist), else /root/reference.

    python tools/refstack_bench.py --config 1   # TinyLlama-1.1B Q4_K_M, 128-token greedy, one request at a time
    python tools/refstack_bench.py --config 5   # Llama-3-8B Q4_K_M, 16 concurrent streaming requests (MAX_CONCURRENT_REQUESTS=16)
Prints one JSON object: the reference benchmark's own report + what was run.
"""
import argparse
import http.client
import json
import os
import shutil
import signal
import socket
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ref_dir() -> str:
    for c in (os.environ.get("REF_DIR"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if c and os.path.exists(os.path.join(c, "scripts", "start.sh")):
            return c
    raise SystemExit("reference scripts not found (REF_DIR, baseline/_ref, /root/reference)")


def free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def get(port: int, path: str, key: str | None = None):
    c = http.client.HTTPConnection("127.0.0.1", port, timeout=30)
    c.request("GET", path, headers={"Authorization": f"Bearer {key}"} if key else {})
    r = c.getresponse()
    data = r.read()
    c.close()
    return r.status, data


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, choices=(1, 5), default=1)
    ap.add_argument("--stage", action="store_true", help="copy /root/reference/scripts to baseline/_ref/scripts and exit")
    ap.add_argument("--requests", type=int, default=0)
    ap.add_argument("--max-tokens", type=int, default=128)
    ap.add_argument("--temp", type=float, default=0.0, help="server default temperature (0 = greedy; 0.8 = llama-server's default sampling chain)")
    args = ap.parse_args()
    if args.stage:
        dst = os.path.join(ROOT, "baseline", "_ref", "scripts")
        shutil.rmtree(dst, ignore_errors=True)
        shutil.copytree("/root/reference/scripts", dst, ignore=shutil.ignore_patterns("__pycache__", "tests", "dev"))
        print("staged", dst)
        return 0

    import bench
    ref = ref_dir()
    preset, conc = ("tinyllama-1.1b", 1) if args.config == 1 else ("llama3-8b", 16)
    nreq = args.requests or (8 if conc == 1 else 48)
    model = bench.model_path(preset, "Q4_K_M", 0xB200)
    tmp = tempfile.mkdtemp(prefix="refstack-")
    data = os.path.join(tmp, "data")
    os.makedirs(os.path.join(data, "models"))
    os.symlink(model, os.path.join(data, "models", "model.gguf"))
    user_key = "sk-bench-" + "k" * 40
    with open(os.path.join(data, "api_keys.txt"), "w") as f:
        f.write(f"bench:{user_key}\n")
    script = open(os.path.join(ref, "scripts", "start.sh")).read()
    assert "/app/llama-server" in script and "/opt/app/scripts" in script
    script = script.replace("/app/llama-server", os.path.join(ROOT, "bin", "llama-server")).replace("/opt/app/scripts", os.path.join(ref, "scripts"))
    gport, bport, hport = free_port(), free_port(), free_port()
    ctx = 1024 * conc          # llama-server divides -c by --parallel: 1024 positions per slot
    env = {**os.environ, "DATA_DIR": data, "MODEL_NAME": "model.gguf", "PORT": str(gport), "PORT_BACKEND": str(bport),
           "PORT_HEALTH": str(hport), "NGL": "99", "CTX": str(ctx), "THREADS": "4",
           "EXTRA_ARGS": f"--parallel {conc} --temp {args.temp:g} --ignore-eos", "AUTH_ENABLED": "true", "INSTANCE_ID": f"refstack-{os.getpid()}",
           "MAX_CONCURRENT_REQUESTS": str(conc), "MAX_QUEUE_SIZE": "64", "MAX_REQUESTS_PER_MINUTE": "100000"}
    t_start = time.time()
    proc = subprocess.Popen(["bash", "-s"], stdin=subprocess.PIPE, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            env=env, start_new_session=True)
    proc.stdin.write(script)
    proc.stdin.close()
    lines: list[str] = []
    threading.Thread(target=lambda: lines.extend(proc.stdout), daemon=True).start()
    out = {"config": args.config, "model": f"{preset} Q4_K_M synthetic (seed 0xb200)", "concurrency": conc, "max_tokens": args.max_tokens,
           "stack": "reference scripts/start.sh -> bin/llama-server (GPU) + reference health_server.py + gateway.py (key auth); reference benchmark.py"}
    rc = 1
    try:
        ok = False
        for _ in range(1800):
            if proc.poll() is not None:
                break
            try:
                if get(gport, "/ping")[0] == 200 and any("Services running" in ln for ln in lines):
                    ok = True
                    break
            except OSError:
                pass
            time.sleep(0.1)
        if not ok:
            sys.stderr.write("".join(lines)[-6000:])
            return 1
        out["seconds_to_services_running"] = round(time.time() - t_start, 2)
        log = "".join(lines)
        out["launcher_saw"] = {k: (k in log) for k in ("llama-server: version:", "Backend responds correctly", "Services running")}
        st, h = get(gport, "/health")
        out["gateway_health"] = json.loads(h) if st == 200 else st
        t_b = time.time()
        r = subprocess.run([sys.executable, os.path.join(ref, "scripts", "benchmark.py"), "--url", f"http://127.0.0.1:{gport}",
                            "--api-key", user_key, "--requests", str(nreq), "--warmup", str(min(2, conc) if conc == 1 else conc),
                            "--max-tokens", str(args.max_tokens), "--concurrency", str(conc), "--output", "json"],
                           capture_output=True, text=True, timeout=1500)
        if r.returncode != 0:
            sys.stderr.write(r.stderr[-4000:])
            return 1
        rep = json.loads(r.stdout)
        out["benchmark_py"] = rep
        out["benchmark_wall_s"] = round(time.time() - t_b, 2)
        # the reference client counts whitespace-separated words of the streamed deltas; the backend's own count is in /metrics
        bkey = None
        for d in ("/dev/shm/llama-keys", "/tmp/llama-keys"):
            kf = os.path.join(d, f"backend-{env['INSTANCE_ID']}.key")
            if os.path.exists(kf):
                bkey = open(kf).read().strip()
        st, m = get(bport, "/metrics", key=bkey)
        if st == 200:
            out["backend_metrics"] = {ln.split()[0]: float(ln.split()[1]) for ln in m.decode().splitlines()
                                      if ln and not ln.startswith("#") and len(ln.split()) == 2}
        st, m = get(gport, "/metrics", key=user_key)
        if st == 200:
            try:
                out["gateway_metrics"] = json.loads(m)
            except ValueError:
                pass
        os.killpg(proc.pid, signal.SIGTERM)
        t0 = time.time()
        proc.wait(timeout=60)
        out["sigterm_to_exit_s"] = round(time.time() - t0, 2)
        rc = 0
    finally:
        if proc.poll() is None:
            os.killpg(proc.pid, signal.SIGKILL)
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(out))
    return rc


if __name__ == "__main__":
    sys.exit(main())
