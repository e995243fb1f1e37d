"""Prefill benchmark (BASELINE.json config 3): T-token prompt through the tcgen05 GEMM path, plus each GEMM shape alone."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ggufb200 import cabi, synth  # noqa: E402
from ggufb200.model import Engine  # noqa: E402


def gemm_alone(L, qt, rows, k, T, reps=10):
    dev = torch.device("cuda", 0)
    stride = L.ggb_repacked_row_stride(qt, k)
    w = torch.randint(0, 256, (rows * stride + 16,), dtype=torch.uint8, device=dev)
    x = torch.randn(T, k, device=dev).to(torch.float16)
    y = torch.empty(T, rows, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        cabi.check(L.ggb_gemm(qt, w.data_ptr(), rows, k, x.data_ptr(), T, y.data_ptr(), rows, s))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        cabi.check(L.ggb_gemm(qt, w.data_ptr(), rows, k, x.data_ptr(), T, y.data_ptr(), rows, s))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, 2.0 * rows * k * T / ms / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--ftype", default="Q4_K_M")
    ap.add_argument("--tokens", type=int, default=2048)
    ap.add_argument("--gemm-only", action="store_true")
    ap.add_argument("--one", action="store_true", help="a single GEMM shape (for ncu)")
    ap.add_argument("--profile-last", action="store_true", help="cudaProfilerStart/Stop around the last timed prefill (ncu --profile-from-start off)")
    a = ap.parse_args()
    L = cabi.lib()
    T = a.tokens
    names = {12: "Q4_K", 14: "Q6_K", 8: "Q8_0"}
    for qt in ((12,) if a.one else (12, 14, 8)):
        for rows, k in (((14336, 4096),) if a.one else ((4096, 4096), (14336, 4096), (4096, 14336))):
            ms, tf = gemm_alone(L, qt, rows, k, T)
            print(f"gemm {names[qt]} rows={rows:6d} k={k:6d} tokens={T}: {ms:7.3f} ms  {tf:7.1f} TFLOP/s", flush=True)
    if a.gemm_only or a.one:
        return
    cfg = synth.PRESETS[a.model]
    path = f"/dev/shm/pf-{a.model}-{a.ftype}.gguf"
    if not os.path.exists(path):
        synth.write_gguf(path, cfg, a.ftype, 0xB200)
    eng = Engine(path, n_ctx=T + 64)
    eng.warmup()
    rng = np.random.default_rng(0)
    prompt = [1] + [int(t) for t in rng.integers(300, cfg.vocab, size=T - 1)]
    for rep in range(3):
        eng.reset()
        torch.cuda.synchronize()
        if a.profile_last and rep == 2:
            torch.cuda.profiler.start()
        t0 = time.perf_counter()
        eng.prefill(prompt)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if a.profile_last and rep == 2:
            torch.cuda.profiler.stop()
        flops = 2.0 * cfg.n_params_matmul * T
        print(f"prefill {a.model} {a.ftype} {T} tokens: {dt*1e3:8.1f} ms  {T/dt:9.0f} tok/s  {flops/dt/1e12:6.1f} TFLOP/s (matmul flops only)", flush=True)
    eng.decode(16)
    print("first tokens after prefill:", eng.tokens(8))


if __name__ == "__main__":
    main()
