#!/usr/bin/env python
"""bench.py -- headline benchmark: Q4_K_M batch-1 greedy decode tok/s (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model llama3-8b]

A "step" is one decoded token (one pass of the hot path: 32 layers + lm-head + arg-max over the synthetic
Llama-3-8B Q4_K_M weights resident in HBM).  Prints ONE JSON line (rank 0).

  value      device-timed tok/s over K graph-replayed decode steps (CUDA events on the launching stream,
             weights 4.6 GB >> 126 MB L2, so no L2 flush is needed between steps);
  e2e        the same metric through the public API (Engine.generate with per-token streaming read-back:
             prompt tokens copied host->device from pinned memory, every generated token read device->host);
  roofline   the GEMV kernel: canonical GGUF weight bytes of one token / device time of one token's GEMV
             launches (timed alone, CUDA events), against MEASURED_PEAKS.json's HBM copy bandwidth;
  cpu_baseline  the CPU oracle (oracle/, a restatement of ggml's CPU path -- "port") on the box's host cores,
             on a bounded sample (one transformer layer per timed step + a few lm-head passes, scaled to a token).

`--impl reference` times that same CPU port as the reference arm (the reference repo ships no engine source:
its backend is a prebuilt third-party binary, SURVEY.md section 0, so oracle/_ref cannot exist).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "decode_tok_per_s_q4_k_m_bs1"
PROMPT = [1, 300, 301, 302, 303]


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def model_path(preset: str, ftype: str, seed: int, sample_layers: int | None = None) -> str:
    """Synthetic GGUF on tmpfs (written once per box).  sample_layers trims the depth for the CPU sample."""
    from ggufb200 import synth
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    tag = f"-L{sample_layers}" if sample_layers else ""
    path = os.path.join(base, f"ggufb200-{preset}{tag}-{ftype}-{seed:x}.gguf")
    if not os.path.exists(path):
        cfg = synth.PRESETS[preset]
        if sample_layers:
            from dataclasses import replace
            cfg = replace(cfg, n_layer=sample_layers)
        tmp = path + f".tmp{os.getpid()}"
        synth.write_gguf(tmp, cfg, ftype, seed)
        os.replace(tmp, path)
    return path


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_sample(preset: str, ftype: str, seed: int, steps: int, warmup: int, budget_s: float):
    """Time the CPU oracle on a bounded sample: each step = ONE transformer layer of the model at its real
    dimensions (decode position ~ a few tokens in); the lm-head is timed separately; a token costs
    n_layer * t_layer + t_head.  Returns (tok_per_s, cores, sample description, ms per token)."""
    from oracle import oracle as O
    from ggufb200 import synth
    cfg = synth.PRESETS[preset]
    path = model_path(preset, ftype, seed, sample_layers=2)
    m = O.OracleLlama(path, n_ctx=64)
    cores = int(O.lib().gref_num_threads())
    x = m.embed(PROMPT[1])
    t_layer, pos, t_start = [], 0, time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        x = m.layer(i % 2, x, pos % 32)
        dt = time.perf_counter() - t0
        if i >= warmup:
            t_layer.append(dt)
        pos += (i % 2)
        if i >= warmup and time.perf_counter() - t_start > budget_s:
            break
    t_head = []
    for _ in range(3):
        t0 = time.perf_counter()
        m.head(x)
        t_head.append(time.perf_counter() - t0)
    tl, th = float(np.mean(t_layer)), float(np.min(t_head))
    t_tok = cfg.n_layer * tl + th
    desc = (f"{len(t_layer)} timed layer-steps (1 of {cfg.n_layer} layers each, real dims, {ftype}) + 3 lm-head passes; "
            f"token = {cfg.n_layer}*{tl*1e3:.2f} ms + {th*1e3:.2f} ms; CPU restatement of ggml (not the llama.cpp binary)")
    return 1.0 / t_tok, cores, desc, t_tok * 1e3, len(t_layer)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tps, cores, desc, ms, n = cpu_sample(args.model, args.ftype, args.seed, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8xint4/6->int32, f32 accumulate", "data": "synthetic",
        "config": {"workload": f"{args.model} {args.ftype} synthetic GGUF, bs=1 greedy decode", "timed_layer_steps": n},
        "cpu_baseline": {"value": tps, "unit": "tok/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": tps, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from ggufb200 import synth
    from ggufb200.model import Engine

    cfg = synth.PRESETS[args.model]
    bpt = synth.weight_bytes_per_token(cfg, args.ftype)
    if rank == 0:
        path = model_path(args.model, args.ftype, args.seed)
    if world > 1:
        dist.barrier()
    path = model_path(args.model, args.ftype, args.seed)
    n_ctx = max(1024, len(PROMPT) + args.steps + args.warmup + 64)
    tp = world if (world > 1 and not args.replicas) else 1
    eng = Engine(path, n_ctx=n_ctx, device=local_rank, use_graph=True, use_pdl=not args.no_pdl,
                 tp_rank=rank if tp > 1 else 0, tp_size=tp)
    eng.warmup()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed decode: W warm-up steps, then exactly K steps
    eng.reset()
    eng.prefill(PROMPT)
    eng.decode(args.warmup)
    sampler = ClockSampler(local_rank)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
    with torch.cuda.stream(eng.stream):
        ev0.record(eng.stream)
    eng.decode(args.steps)
    with torch.cuda.stream(eng.stream):
        ev1.record(eng.stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    pos_end = len(PROMPT) + args.warmup + args.steps

    # ---- the GEMV kernel alone: one token's worth of GEMV launches (no attention, no sampler), graph-replayed
    import ctypes as C
    from ggufb200 import cabi
    g = torch.cuda.CUDAGraph()
    n_gemv = 0
    with torch.cuda.stream(eng.stream):
        with torch.cuda.graph(g, stream=eng.stream):
            s = torch.cuda.current_stream().cuda_stream
            for qkv, o, gu, dn in eng._layer_args:
                for a in (qkv, o, gu, dn):
                    cabi.check(eng.lib.ggb_gemv(C.byref(a), s))
                    n_gemv += 1
            cabi.check(eng.lib.ggb_gemv(C.byref(eng._head), s))
            n_gemv += 1
        for _ in range(3):
            g.replay()
        gv0, gv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        gv0.record(eng.stream)
        for _ in range(reps):
            g.replay()
        gv1.record(eng.stream)
    torch.cuda.synchronize()
    gemv_ms_per_token = gv0.elapsed_time(gv1) / reps
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public API, streaming read-back per token
    barrier()
    t0 = time.perf_counter()
    toks = eng.generate(PROMPT, args.steps, stream_cb=lambda t: None)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    assert len(toks) == args.steps

    tmax = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_all, e2e_ms_all = float(tmax[0]), float(tmax[1])

    exchange = ("over NVLink peer memory, fused into the GEMV epilogue (csrc/peer.cu)" if getattr(eng, "peer", None)
                else "by NCCL all-reduce")
    if rank == 0:
        peak, peak_src = load_peaks()
        streams = 1 if tp > 1 else world          # tensor parallel: one token stream over all GPUs; replicas: one per GPU
        value = streams * args.steps / (ms_all / 1e3)
        e2e = streams * args.steps / (e2e_ms_all / 1e3)
        gemv_gbs = bpt["weights"] / tp / (gemv_ms_per_token / 1e3) / 1e9      # per GPU
        kv_mid = bpt["kv_per_pos"] * (pos_end - args.steps / 2)
        step_bytes = bpt["weights"] + bpt["norms"] + bpt["embed_row"] + kv_mid
        line = {
            "metric": METRIC, "value": value, "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_all / args.steps, "higher_is_better": True, "scaling": "strong" if tp > 1 else "weak", "vs_baseline": None,
            "dtype": "int8xint4/6->int32 (dp4a), f32 accumulate", "data": "synthetic",
            "config": {
                "workload": f"{args.model} {args.ftype} synthetic GGUF (seed {args.seed:#x}), bs=1 greedy decode, {args.steps} tokens after a {len(PROMPT)}-token prompt",
                "parallelism": "single GPU" if world == 1 else (
                    f"tp{tp}: q/k/v/gate/up/lm-head column-split, attn_output/ffn_down row-split, f64 partials exchanged "
                    f"{exchange} after each row-split projection ({2 * cfg.n_layer} per token) + one 8-byte NCCL arg-max all-reduce" if tp > 1
                    else f"{world} independent replicas"),
                "l2": f"inputs_larger_than_l2 ({bpt['weights'] / tp / 1e9:.1f} GB of weights per GPU per step vs 126 MB L2)",
                "launch": "CUDA graph per step" + ("" if args.no_pdl else " + programmatic dependent launch"),
                "step_bytes": int(step_bytes),
                "step_roofline_frac_of_measured_hbm": (step_bytes * value / world / 1e9) / peak,
                "roofline_tok_s_at_measured_hbm": peak * 1e9 / step_bytes * (world if tp > 1 else 1),
                "n_ctx": n_ctx,
            },
            "roofline": {"bound": "hbm", "kernel": "gemv_kernel (all GEMV launches of one token)", "achieved": gemv_gbs, "peak": peak,
                         "unit": "GB/s", "frac": gemv_gbs / peak, "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_token": bpt["weights"], "per_gpu": tp > 1, "launches_per_token": n_gemv,
                         "avg_launch_us": gemv_ms_per_token * 1e3 / n_gemv},
            "e2e": {"value": e2e, "unit": "tok/s", "h2d_bytes_per_step": len(PROMPT) * 8 / args.steps, "d2h_bytes_per_step": 4,
                    "api": "Engine.generate(prompt, n, stream_cb) -- token read back to pinned host memory every step"},
            "gpu_launches": eng.launches_per_step() * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            tps, cores, desc, _, _ = cpu_sample(args.model, args.ftype, args.seed, 48, 4, budget_s=25.0)
            line["cpu_baseline"] = {"value": tps, "unit": "tok/s", "cores": cores, "kind": "port", "sample": desc}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--ftype", default="Q4_K_M")
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=0xB200)
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    ap.add_argument("--replicas", action="store_true", help="N>1: independent replicas instead of tensor parallelism")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
