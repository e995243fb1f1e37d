"""`llama-server` drop-in entry point: the process contract of the reference's launcher.

scripts/start.sh runs `/app/llama-server --version` (:359-365), then spawns
`llama-server -m MODEL --host 127.0.0.1 --port P -c CTX -ngl NGL --api-key-file F [-t T] [EXTRA_ARGS...]`
(:473-494, :516), expects it alive after 2 s (:527-551), bound to 127.0.0.1 only (:566-590), answering
`GET /health` 2xx within 30 polls (:600-635), logging to stdout/stderr (:516) and exiting on SIGTERM (:402-430)
-- or on a closed stdout pipe, because `$!` of the tee pipeline is tee (:516-517).
"""
from __future__ import annotations

import argparse
import os
import signal
import sys
import threading
import time

from . import __version__

VERSION_LINE = f"version: {__version__} (ggufb200: B200-native GGUF decode engine, sm_100a, CUDA kernels via libggufb200.so)"


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="llama-server", add_help=True, allow_abbrev=False,
                                 description="B200-native GGUF decode engine with llama-server's process and HTTP contract")
    ap.add_argument("--version", action="store_true")
    ap.add_argument("-m", "--model")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=8080)
    ap.add_argument("-c", "--ctx-size", type=int, default=4096)
    ap.add_argument("-ngl", "--gpu-layers", "--n-gpu-layers", dest="ngl", default="99")
    ap.add_argument("--api-key-file")
    ap.add_argument("--api-key")
    ap.add_argument("-t", "--threads", default=None)
    ap.add_argument("-np", "--parallel", type=int, default=1)
    ap.add_argument("-a", "--alias")
    ap.add_argument("--temp", "--temperature", dest="temp", type=float, default=0.8)
    ap.add_argument("--top-k", type=int, default=40)
    ap.add_argument("--top-p", type=float, default=0.95)
    ap.add_argument("-s", "--seed", type=int, default=None)
    ap.add_argument("--ignore-eos", action="store_true")
    ap.add_argument("--embeddings", "--embedding", dest="embeddings", action="store_true", help="serve POST /v1/embeddings")
    ap.add_argument("--pooling", default="mean", choices=["mean", "last"], help="pooling of the embeddings endpoint")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-pdl", action="store_true", help="disable programmatic dependent launch")
    ap.add_argument("-v", "--verbose", action="store_true")
    return ap


class Log:
    """stdout logger; a closed pipe (tee died: start.sh's shutdown path) ends the process quietly"""

    def __init__(self):
        self.lock = threading.Lock()

    def __call__(self, *a):
        try:
            with self.lock:
                print(time.strftime("%H:%M:%S"), *a, flush=True)
        except (BrokenPipeError, OSError):
            os._exit(0)


def read_key(args) -> str | None:
    if args.api_key:
        return args.api_key.strip()
    if args.api_key_file:
        with open(args.api_key_file) as f:
            for line in f:
                line = line.strip()
                if line:
                    return line
        raise SystemExit(f"error: --api-key-file {args.api_key_file} holds no key")
    return None


def main(argv=None, engine_factory=None) -> int:
    """engine_factory(args) -> engine: dependency injection for the CPU tests of the process contract (tests/fake_llama_server.py
    puts the oracle behind this very main loop); the product entry point never passes it -- the engine is the GPU one or nothing."""
    argv = sys.argv[1:] if argv is None else argv
    args, unknown = build_parser().parse_known_args(argv)
    if args.version:
        print(VERSION_LINE)
        return 0
    log = Log()
    # one process per GPU under torchrun: tensor parallelism; rank 0 serves, the others follow (tp_serve.py)
    rank, world, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        args.device = local_rank
        if rank != 0:
            from .model import Engine
            from .tp_serve import follower_loop
            eng = Engine(args.model, n_ctx=args.ctx_size, device=local_rank, use_graph=not args.no_graph, use_pdl=not args.no_pdl,
                         n_slots=max(1, args.parallel), tp_rank=rank, tp_size=world)
            eng.warmup()
            follower_loop(eng, dist)
            eng.close()
            dist.destroy_process_group()
            return 0
    if unknown:
        log(f"warn: ignoring unsupported arguments: {' '.join(unknown)}")
    if not args.model:
        print("error: -m/--model is required", file=sys.stderr)
        return 1
    if not os.path.isfile(args.model):
        print(f"error: model file not found: {args.model}", file=sys.stderr)
        return 1
    if str(args.ngl) == "0":
        log("warn: -ngl 0 requested; this engine has no CPU path, every layer runs on the GPU")
    key = read_key(args)

    from .scheduler import SamplingParams, Scheduler
    from .server import ServerState, make_server
    from .gguf_reader import GGUFFile
    from .tokenizer import Tokenizer

    log(VERSION_LINE)
    log(f"main: binding HTTP server to {args.host}:{args.port}")
    # tokenizer first (cheap), then the listening socket, then the weights: /health says 503 "Loading model" meanwhile
    gf = GGUFFile(args.model)
    tok = Tokenizer(gf.meta)
    name = args.alias or os.path.basename(args.model)
    info = {"vocab": tok.n_vocab, "n_ctx": args.ctx_size, "tokenizer": tok.model, "version": __version__}
    gf.close()

    class _Pending:  # scheduler placeholder while the engine loads
        fatal = None
        active, pending, stats = {}, [], {"prompt_tokens": 0, "completion_tokens": 0, "decode_seconds": 0.0, "requests": 0}
        engine = type("E", (), {"slots": []})()

        def idle_slots(self):
            return 0

    defaults = SamplingParams(temperature=args.temp, top_k=args.top_k, top_p=args.top_p, seed=args.seed)
    state = ServerState(_Pending(), tok, name, key, args.ctx_size, defaults, log=log, info=info)
    try:
        httpd = make_server(args.host, args.port, state)
    except OSError as e:
        print(f"error: cannot bind {args.host}:{args.port}: {e}", file=sys.stderr)
        return 1
    threading.Thread(target=httpd.serve_forever, daemon=True, name="http").start()

    stop = threading.Event()

    def on_signal(signum, frame):
        log(f"main: received signal {signum}, shutting down")
        stop.set()

    signal.signal(signal.SIGTERM, on_signal)
    signal.signal(signal.SIGINT, on_signal)

    try:
        from .model import Engine
        t0 = time.time()
        log(f"main: loading model {args.model}")
        if engine_factory is not None:
            eng = engine_factory(args)
        else:
            eng = Engine(args.model, n_ctx=args.ctx_size, device=args.device, use_graph=not args.no_graph,
                         use_pdl=not args.no_pdl, n_slots=max(1, args.parallel), verbose=args.verbose,
                         tp_rank=rank if world > 1 else 0, tp_size=world)
        eng.warmup()
        if world == 1 and len(eng.slots) > 1 and os.environ.get("GGB_WARM_BATCH", "1") != "0" and hasattr(getattr(eng, "batch", None), "warmup"):
            eng.batch.warmup()      # every batch size's graph now, not under the first burst of requests
        info.update({"n_layer": eng.hp.n_layer, "n_embd": eng.hp.d, "weights_gb": round(eng.weight_bytes / 1e9, 3)})
        log(f"main: model loaded in {time.time() - t0:.2f} s ({eng.weight_bytes / 1e9:.2f} GB of weights in HBM, "
            f"{len(eng.slots)} slot(s), context {args.ctx_size})")
    except Exception as e:
        print(f"error: failed to load model: {e!r}", file=sys.stderr)
        httpd.shutdown()
        if world > 1:   # followers that did load are blocked in their command broadcast: release them
            _tp_release(dist)
        return 1
    leader = None
    if world > 1:
        from .tp_serve import TPLeader
        leader = TPLeader(eng, dist)
    sched = Scheduler(leader or eng, tok, ignore_eos=args.ignore_eos, log=log)
    sched.start()
    state.sched = sched
    if args.embeddings:
        if leader is not None or not hasattr(eng, "embed"):
            log("warn: --embeddings is served by the single-GPU prefill path only; /v1/embeddings stays 501")
        else:
            state.embeddings = args.pooling
    state.ready.set()
    log(f"main: server is listening on http://{args.host}:{args.port} - starting the main loop")

    while not stop.is_set():
        if not sched.is_alive():
            log("main: the scheduler thread died (engine failure), exiting")
            httpd.shutdown()
            if leader is not None:
                _tp_release(dist, leader)
            return 1
        stop.wait(0.5)
    sched.shutdown()
    httpd.shutdown()
    if leader is not None:
        sched.join(timeout=30)
        leader.shutdown()
        dist.destroy_process_group()
    log("main: clean exit")
    return 0


def _tp_release(dist, leader=None):
    """rank 0 is leaving on an error path: tell the follower ranks to stop (they wait in broadcast_object_list and would
    otherwise hold their GPUs until the NCCL timeout) and tear the process group down; best effort, never raises"""
    try:
        if leader is not None:
            leader.shutdown()
        else:
            dist.broadcast_object_list([("stop",)], src=0)
        dist.destroy_process_group()
    except Exception as e:  # the collective itself may be what failed
        print(f"warn: could not release the tensor-parallel followers: {e!r}", file=sys.stderr)


if __name__ == "__main__":
    sys.exit(main())
