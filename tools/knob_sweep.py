"""Sweep decode-chain knobs (environment variables read when an Engine is built) on the headline workload:
Llama-3-8B Q4_K_M, bs=1 greedy decode after a 256-token prompt, device-timed graph replays.
    python tools/knob_sweep.py "" "GGB_ATTN_CL=4" ...
Every configuration must produce the same tokens as the first one (the knobs are hints, never arithmetic)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ggufb200.model import Engine  # noqa: E402

KNOBS = ("GGB_ATTN_CL", "GGB_GEMV_CTAS_PER_SM", "GGB_LIB_PATH", "GGB_GEMV_GRID", "GGB_ATTN_NW")


def run(path, steps, warm):
    eng = Engine(path, n_ctx=1024)
    eng.warmup()
    eng.reset()
    eng.prefill([1] + list(range(300, 300 + int(os.environ.get('SWEEP_PROMPT', '256')) - 1)))
    eng.decode(warm)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        if eng._s0.n_past + steps >= 1000:
            break
        torch.cuda.synchronize()
        e0.record(eng.stream)
        eng.decode(steps)
        e1.record(eng.stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    toks = eng.tokens(warm + 32)
    eng.close()
    del eng
    torch.cuda.empty_cache()
    return best, toks


def main():
    model = os.environ.get("SWEEP_MODEL", "llama3-8b")
    ftype = os.environ.get("SWEEP_FTYPE", "Q4_K_M")
    steps = int(os.environ.get("SWEEP_STEPS", "128"))
    path = bench.model_path(model, ftype, 0xB200)
    ref = None
    for spec in sys.argv[1:] or [""]:
        for k in KNOBS:
            os.environ.pop(k, None)
        for kv in spec.split():
            k, v = kv.split("=")
            os.environ[k] = v
        ms, toks = run(path, steps, 16)
        if ref is None:
            ref = toks
        same = "same tokens" if toks == ref else "TOKENS DIFFER"
        print(f"{spec or '(defaults)':60s} {ms * 1e3:8.1f} us/token  {1e3 / ms:7.1f} tok/s  {same}", flush=True)


if __name__ == "__main__":
    main()
