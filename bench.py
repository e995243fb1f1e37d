#!/usr/bin/env python
"""bench.py -- headline benchmark: Q4_K_M batch-1 greedy decode tok/s (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model llama3-8b]

A "step" is one decoded token (one pass of the hot path: 32 layers + lm-head + arg-max over the synthetic
Llama-3-8B Q4_K_M weights resident in HBM) at a context of ~256 positions (the KV cache is pre-filled outside the
timed region, so that a short run sits where the 512-token configuration spends its time).  Prints ONE JSON line
(rank 0).

  value      device-timed tok/s over K graph-replayed decode steps (CUDA events on the launching stream,
             weights 4.6 GB >> 126 MB L2, so no L2 flush is needed between steps);
  e2e        the same metric through the public host-driven API, every step: the host hands the engine the previous
             token (Slot.feed: token id + position copied host->device from pinned memory), the engine runs the step,
             the host reads the new token back (device->host into pinned memory);
  roofline   the GEMV kernel: canonical GGUF weight bytes of one token / device time of one token's GEMV
             launches (timed alone, CUDA events), against MEASURED_PEAKS.json's HBM copy bandwidth; `traffic` is the
             DRAM byte count of the same launches from the committed ncu capture (profiles/r02_gemv_traffic.json);
  cpu_baseline  the CPU oracle (oracle/, a restatement of ggml's CPU path -- "port") on the box's host cores,
             on a bounded sample (one transformer layer per timed step + a few lm-head passes, scaled to a token);
  parity     the first 16 greedy tokens of the TIMED model and engine against the CPU oracle (N=1: computed live in the
             cpu_baseline leg; N>1: against tests/golden/bench_tokens.json, written by the same oracle);
  config4    (N >= 2 only) BASELINE.json config 4 next to the headline: Llama-3-70B Q4_K_M tensor-parallel over the
             same N GPUs -- tok/s, ms/step, per-GPU GEMV roofline fraction.

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
PROMPT = [1, 300, 301, 302, 303]                       # parity block: short prompt, bit-exact integer path end to end
SEED_DEFAULT = 0xB200
CONTEXT = 256                                           # positions in the KV cache when the timed region starts


def context_prompt(vocab: int) -> list[int]:
    """CONTEXT deterministic token ids (pre-fills the KV cache outside the timed region)"""
    return [1] + [300 + (i * 7919) % (min(vocab, 30000) - 300) for i in range(CONTEXT - 1)]


def workload(args) -> str:
    """the same string in both arms (the driver compares them)"""
    return (f"{args.model} {args.ftype} synthetic GGUF (seed {args.seed:#x}), bs=1 greedy decode at ~{CONTEXT} positions "
            f"of context")


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(model: str, ftype: str):
    """DRAM bytes (read + write) of one token's GEMV launches from the committed ncu --set full capture, or None"""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_gemv_traffic.json")) as f:
            t = json.load(f)
        return t.get(f"{model}/{ftype}")
    except Exception:
        return None


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

    def wait_first(self, timeout_s: float = 8.0):
        """block until nvidia-smi has delivered its first row (it needs 0.3-2 s to start on a multi-GPU box): the timed region
        that follows is short, and a sampler that only starts reporting after it would say nothing about it"""
        t0 = time.time()
        while self.proc and not self.rows and time.time() - t0 < timeout_s:
            time.sleep(0.02)

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


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_sample(preset: str, ftype: str, seed: int, steps: int, warmup: int, budget_s: float):
    """Time the CPU oracle on a bounded sample: each step = ONE transformer layer of the model at its real
    dimensions at ~CONTEXT positions of context; the lm-head is timed separately; a token costs
    n_layer * t_layer + t_head.  The OpenMP thread count is set explicitly to the cores this process may use
    (torchrun exports OMP_NUM_THREADS=1).  Returns (tok_per_s, cores, sample description, ms per token, timed steps)."""
    from oracle import oracle as O
    from ggufb200 import synth
    cfg = synth.PRESETS[preset]
    path = model_path(preset, ftype, seed, sample_layers=2)
    cores = host_threads()
    m = O.OracleLlama(path, n_ctx=CONTEXT + 64, nthreads=cores)
    x = m.embed(PROMPT[1])
    t_layer, t_start = [], time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        x = m.layer(i % 2, x, CONTEXT + (i // 2) % 32)   # attention over ~CONTEXT cached positions
        dt = time.perf_counter() - t0
        if i >= warmup:
            t_layer.append(dt)
        if i >= warmup and time.perf_counter() - t_start > budget_s:
            break
    t_head = []
    for _ in range(3):
        t0 = time.perf_counter()
        m.head(x)
        t_head.append(time.perf_counter() - t0)
    tl, th = float(np.mean(t_layer)), float(np.min(t_head))
    t_tok = cfg.n_layer * tl + th
    desc = (f"{len(t_layer)} timed layer-steps (1 of {cfg.n_layer} layers each, real dims, {ftype}, ~{CONTEXT} cached positions) + 3 lm-head "
            f"passes; token = {cfg.n_layer}*{tl*1e3:.2f} ms + {th*1e3:.2f} ms; {cores} OpenMP threads; CPU restatement of ggml "
            f"(not the llama.cpp binary)")
    return 1.0 / t_tok, cores, desc, t_tok * 1e3, len(t_layer)


def oracle_tokens(preset: str, ftype: str, seed: int, n: int) -> list[int]:
    """first n greedy tokens of the full model after PROMPT, CPU oracle (canon mode)"""
    from oracle import oracle as O
    m = O.OracleLlama(model_path(preset, ftype, seed), n_ctx=len(PROMPT) + n + 8, nthreads=host_threads(), mode="canon")
    return [int(t) for t in m.greedy(PROMPT, n)]


def golden_tokens(preset: str, ftype: str, seed: int):
    try:
        with open(os.path.join(ROOT, "tests", "golden", "bench_tokens.json")) as f:
            return json.load(f)[f"{preset}/{ftype}/{seed:#x}"]["tokens"]
    except Exception:
        return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tps, cores, desc, ms, n = cpu_sample(args.model, args.ftype, args.seed, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": "tok/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int8xint4/6->int32, f32 accumulate", "data": "synthetic",
        "config": {"workload": workload(args), "timed_layer_steps": n},
        "cpu_baseline": {"value": tps, "unit": "tok/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": tps, "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def time_decode(eng, steps, warmup, prompt, barrier, sampler=None, rank=0):
    """W untimed + K timed graph-replayed decode steps after `prompt`; device time (ms) of the K steps"""
    import torch
    eng.reset()
    eng.prefill(prompt)
    eng.decode(warmup)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if sampler is not None and rank == 0:
        sampler.start()          # the fork happens BEFORE the barriers: no rank waits for it inside the timed region
        sampler.wait_first()
    barrier()
    barrier()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()   # no-op unless a profiler attached with --profile-from-start off: then exactly the timed region
    ev0.record(eng.stream)
    eng.decode(steps)
    ev1.record(eng.stream)
    barrier()
    torch.cuda.profiler.stop()
    return ev0.elapsed_time(ev1)


def time_gemv_only(eng, reps=20):
    """one token's GEMV launches alone (no attention, no sampler), graph-replayed; ms per token, launches per token"""
    import ctypes as C
    import torch
    from ggufb200 import cabi
    g = torch.cuda.CUDAGraph()
    n_gemv = 0
    with torch.cuda.stream(eng.stream):
        with torch.cuda.graph(g, stream=eng.stream):
            s = torch.cuda.current_stream().cuda_stream
            for qkv, o, gu, dn in eng._layer_args:
                for a in (qkv, o, gu, dn):
                    if a.epilogue == cabi.EPI_PEER_F64:
                        a = cabi.copy_args(a, epilogue=cabi.EPI_STORE_F64)   # timed alone: no peer to complete the exchange
                    cabi.check(eng.lib.ggb_gemv(C.byref(a), s))
                    n_gemv += 1
            cabi.check(eng.lib.ggb_gemv(C.byref(eng._head), s))
            n_gemv += 1
        for _ in range(3):
            g.replay()
        gv0, gv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gv0.record(eng.stream)
        for _ in range(reps):
            g.replay()
        gv1.record(eng.stream)
    torch.cuda.synchronize()
    return gv0.elapsed_time(gv1) / reps, n_gemv


def time_e2e(eng, steps, prompt, barrier):
    """host-driven decode through the public API: per step feed(previous token) [H2D from pinned memory] + read_last_token()
    [D2H into pinned memory]; seconds for `steps` steps (the prompt is processed before the clock starts)"""
    import torch
    slot = eng.slots[0]
    eng.reset()
    eng.prefill(prompt)
    tok = slot.read_last_token()
    for _ in range(3):
        slot.feed(tok)
        tok = slot.read_last_token()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        slot.feed(tok)
        tok = slot.read_last_token()
    torch.cuda.synchronize()
    return time.perf_counter() - t0


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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*vals):
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def bench_model(model: str, ftype: str, steps: int, warmup: int, with_e2e: bool, sampler):
        cfg = synth.PRESETS[model]
        bpt = synth.weight_bytes_per_token(cfg, ftype)
        if rank == 0:
            model_path(model, ftype, args.seed)
        barrier()
        path = model_path(model, ftype, args.seed)
        n_ctx = max(1024, CONTEXT + steps + warmup + 64)
        tp = world if (world > 1 and not args.replicas) else 1
        eng = Engine(path, n_ctx=n_ctx, device=local_rank, use_graph=True, use_pdl=not args.no_pdl,
                     tp_rank=rank if tp > 1 else 0, tp_size=tp)
        eng.warmup()
        ctx_prompt = context_prompt(cfg.vocab)
        ms = time_decode(eng, steps, warmup, ctx_prompt, barrier, sampler, rank)
        gemv_ms, n_gemv = time_gemv_only(eng)
        clocks = sampler.stop() if (sampler is not None and rank == 0) else None
        e2e_s = time_e2e(eng, steps, ctx_prompt, barrier) if with_e2e else 0.0
        # greedy tokens of THIS engine on the short prompt (bit-exact integer path): checked against the oracle
        toks = eng.generate(PROMPT, 16)
        ms_all, e2e_ms_all, gemv_all = max_over_ranks(ms, e2e_s * 1e3, gemv_ms)
        res = {"cfg": cfg, "bpt": bpt, "tp": tp, "ms": ms_all, "e2e_ms": e2e_ms_all, "gemv_ms": gemv_all, "n_gemv": n_gemv,
               "clocks": clocks, "tokens": [int(t) for t in toks], "n_ctx": n_ctx, "launches": eng.launches_per_step(),
               "peer": bool(getattr(eng, "peer", None))}
        eng.close()
        del eng
        torch.cuda.empty_cache()
        return res

    sampler = ClockSampler(local_rank)
    R = bench_model(args.model, args.ftype, args.steps, args.warmup, True, sampler)
    tp, cfg, bpt = R["tp"], R["cfg"], R["bpt"]
    c4 = None
    if world > 1 and tp > 1 and not args.no_config4:
        try:
            c4 = bench_model("llama3-70b", "Q4_K_M", args.steps, args.warmup, False, None)
        except Exception as e:   # never lose the headline line to the secondary block
            c4 = {"error": repr(e)}

    exchange = ("over NVLink peer memory, fused into the GEMV epilogue (csrc/peer.cu)" if R["peer"] else "by NCCL all-reduce")
    if rank == 0:
        peak, peak_src = load_peaks()
        streams = 1 if tp > 1 else world          # tensor parallel: one token stream over all GPUs; replicas: one per GPU
        value = streams * args.steps / (R["ms"] / 1e3)
        e2e = streams * args.steps / (R["e2e_ms"] / 1e3)
        gemv_gbs = bpt["weights"] / tp / (R["gemv_ms"] / 1e3) / 1e9      # per GPU
        kv_mid = bpt["kv_per_pos"] * (CONTEXT + args.warmup + args.steps / 2)
        step_bytes = bpt["weights"] + bpt["norms"] + bpt["embed_row"] + kv_mid
        traffic = load_traffic(args.model, args.ftype) if tp == 1 else None
        line = {
            "metric": METRIC, "value": value, "unit": "tok/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": R["ms"] / args.steps, "higher_is_better": True, "scaling": "strong" if tp > 1 else "weak", "vs_baseline": None,
            "dtype": "int8xint4/6->int32 (dp4a), f32 accumulate", "data": "synthetic",
            "config": {
                "workload": workload(args),
                "parallelism": "single GPU" if world == 1 else (
                    f"tp{tp}: q/k/v/gate/up/lm-head column-split, attn_output/ffn_down row-split, f64 partials exchanged "
                    f"{exchange} after each row-split projection ({2 * cfg.n_layer} per token) + one 8-byte NCCL arg-max all-reduce" if tp > 1
                    else f"{world} independent replicas"),
                "l2": f"inputs_larger_than_l2 ({bpt['weights'] / tp / 1e9:.1f} GB of weights per GPU per step vs 126 MB L2)",
                "launch": "CUDA graph per step" + ("" if args.no_pdl else " + programmatic dependent launch"),
                "step_bytes": int(step_bytes),
                "step_roofline_frac_of_measured_hbm": (step_bytes * value / world / 1e9) / peak,
                "roofline_tok_s_at_measured_hbm": peak * 1e9 / step_bytes * (world if tp > 1 else 1),
                "n_ctx": R["n_ctx"], "context_positions_at_start": CONTEXT + args.warmup,
            },
            "roofline": {"bound": "hbm", "kernel": "ggb_dq_gemv_kernel (all GEMV launches of one token)", "achieved": gemv_gbs, "peak": peak,
                         "unit": "GB/s", "frac": gemv_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_token": bpt["weights"], "per_gpu": tp > 1, "launches_per_token": R["n_gemv"],
                         "avg_launch_us": R["gemv_ms"] * 1e3 / R["n_gemv"]},
            "e2e": {"value": e2e, "unit": "tok/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": 4,
                    "api": "per step Slot.feed(previous token) -- token id and position copied host->device from pinned memory -- "
                           "then Slot.read_last_token() -- device->host into pinned memory"},
            "gpu_launches": R["launches"] * args.steps,
            "clocks": R["clocks"],
        }
        if world == 1 and not args.no_cpu:
            tps, cores, desc, _, _ = cpu_sample(args.model, args.ftype, args.seed, 48, 4, budget_s=20.0)
            line["cpu_baseline"] = {"value": tps, "unit": "tok/s", "cores": cores, "kind": "port", "sample": desc}
            ref = oracle_tokens(args.model, args.ftype, args.seed, 16)
            src = "CPU oracle (canon mode), run live on this box"
        else:
            ref = golden_tokens(args.model, args.ftype, args.seed)
            src = "tests/golden/bench_tokens.json (CPU oracle, canon mode)"
        line["parity"] = {"greedy_tokens": R["tokens"], "oracle_tokens": ref[:16] if ref else None,
                          "identical": (R["tokens"] == ref[:16]) if ref else None, "oracle": src,
                          "prompt": PROMPT}
        if c4 is not None:
            if "error" in c4:
                line["config4"] = c4
            else:
                b4 = c4["bpt"]
                g4 = b4["weights"] / c4["tp"] / (c4["gemv_ms"] / 1e3) / 1e9
                sb4 = b4["weights"] + b4["norms"] + b4["embed_row"] + b4["kv_per_pos"] * (CONTEXT + args.warmup + args.steps / 2)
                v4 = args.steps / (c4["ms"] / 1e3)
                g70 = golden_tokens("llama3-70b", "Q4_K_M", args.seed)
                line["config4"] = {
                    "workload": f"llama3-70b Q4_K_M synthetic GGUF (seed {args.seed:#x}), bs=1 greedy decode at ~{CONTEXT} positions, tp{c4['tp']}",
                    "value": v4, "unit": "tok/s", "ms_per_step": c4["ms"] / args.steps, "steps": args.steps, "warmup": args.warmup,
                    "gemv_gbs_per_gpu": g4, "gemv_roofline_frac_per_gpu": g4 / peak,
                    "step_roofline_frac_of_measured_hbm": (sb4 * v4 / world / 1e9) / peak,
                    "greedy_tokens": c4["tokens"], "oracle_tokens": g70[:16] if g70 else None,
                    "identical": (c4["tokens"] == g70[:16]) if g70 else None}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=16)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--ftype", default="Q4_K_M")
    ap.add_argument("--seed", type=lambda s: int(s, 0), default=SEED_DEFAULT)
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample and the live oracle parity check")
    ap.add_argument("--no-config4", action="store_true", help="N>1: skip the Llama-3-70B block")
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
