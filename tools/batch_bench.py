"""Aggregate decode throughput of B concurrent sequences: batched (gemv_batch.cu) vs time-sliced single-sequence steps.
usage: python tools/batch_bench.py [--model llama3-8b] [--ftype Q4_K_M] [--batches 2,4,8,16] [--steps 64]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ggufb200.model import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--ftype", default="Q4_K_M")
    ap.add_argument("--batches", default="2,4,8,16")
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--ctx", type=int, default=1024)
    ap.add_argument("--prompt", type=int, default=32)
    args = ap.parse_args()
    path = bench.model_path(args.model, args.ftype, 0xB200)
    bs = [int(b) for b in args.batches.split(",")]
    eng = Engine(path, n_ctx=args.ctx, n_slots=max(bs))
    eng.warmup()
    out = {"model": args.model, "ftype": args.ftype, "steps": args.steps, "prompt_tokens": args.prompt, "rows": []}
    # single-sequence reference (device-side chain, as bench.py times it)
    s0 = eng.slots[0]
    s0.reset(); s0.prefill([1] + list(range(300, 300 + args.prompt - 1))); s0.decode(8)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(eng.stream):
        ev0.record(eng.stream)
    s0.decode(args.steps)
    with torch.cuda.stream(eng.stream):
        ev1.record(eng.stream)
    torch.cuda.synchronize()
    single_ms = ev0.elapsed_time(ev1) / args.steps
    out["single_ms_per_token"] = single_ms
    for B in bs:
        for s in range(B):
            eng.slots[s].reset()
            eng.slots[s].prefill([1] + [300 + s + j for j in range(args.prompt - 1)])
        last = [eng.slots[s].read_last_token() for s in range(B)]
        for _ in range(4):   # warm-up (captures the graph of this batch size)
            last = eng.batch.step([(s, last[s], eng.slots[s].n_past) for s in range(B)])
        torch.cuda.synchronize()
        # device time of the graph alone
        g = eng.batch._graphs[(B, True, False)]
        with torch.cuda.stream(eng.stream):
            ev0.record(eng.stream)
            for _ in range(10):
                g.replay()
            ev1.record(eng.stream)
        torch.cuda.synchronize()
        dev_ms = ev0.elapsed_time(ev1) / 10
        t0 = time.perf_counter()
        for _ in range(args.steps):
            last = eng.batch.step([(s, last[s], eng.slots[s].n_past) for s in range(B)])
        wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        row = {"batch": B, "device_ms_per_step": dev_ms, "e2e_ms_per_step": wall_ms, "agg_tok_s_device": B / dev_ms * 1e3,
               "agg_tok_s_e2e": B / wall_ms * 1e3, "time_sliced_tok_s": 1e3 / single_ms, "speedup_vs_time_sliced": (B / wall_ms) * single_ms}
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/batch_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
