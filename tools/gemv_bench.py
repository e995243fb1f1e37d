"""Micro-benchmark of the fused GEMV launches of one Llama-3-8B layer, each shape timed alone as a chain of
32 launches over 32 DIFFERENT weight buffers (so nothing is served from L2), graph-replayed with PDL.
Prints per-launch microseconds and achieved GB/s against canonical weight bytes."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ggufb200 import cabi  # noqa: E402

DEV = torch.device("cuda", 0)
L = cabi.lib()
BB = {12: 144, 14: 210, 8: 272}


def rand_weight(qt, rows, k):
    stride = L.ggb_repacked_row_stride(qt, k)
    return torch.randint(0, 256, (rows * stride + 16,), dtype=torch.uint8, device=DEV)  # valid packed fields; f16 scales random


def bench(name, segs, k, pro, epi, n_buf=32, reps=20, use_pdl=1):
    x = torch.randn(k, device=DEV) * 0.01
    nw = torch.ones(k, device=DEV)
    rows0 = segs[0][1]
    y = [torch.zeros(max(r, 1), device=DEV) for _, r in segs]
    res = torch.zeros(rows0, device=DEV)
    pos = torch.tensor([7], dtype=torch.int32, device=DEV)
    tab = torch.zeros(64 * 128, device=DEV)
    kc = torch.zeros(64 * 2048, dtype=torch.int16, device=DEV)
    pv = torch.zeros(1024, device=DEV)
    pi = torch.zeros(1024, dtype=torch.int32, device=DEV)
    bufs = [[rand_weight(qt, r, k) for qt, r in segs] for _ in range(n_buf)]
    args = []
    for b in bufs:
        a = cabi.make_gemv_args([(w.data_ptr(), qt, r, yy.data_ptr()) for w, (qt, r), yy in zip(b, segs, y)], k, x.data_ptr(),
                                prologue=pro, epilogue=epi, norm_w=nw.data_ptr(), eps=1e-5, use_pdl=use_pdl, residual=res.data_ptr(),
                                pos_dev=pos.data_ptr(), rope_tab=tab.data_ptr(), n_rot=128, head_dim=128, kcache=kc.data_ptr(),
                                vcache=kc.data_ptr(), part_val=pv.data_ptr(), part_idx=pi.data_ptr())
        args.append(a)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for a in args[:2]:
            cabi.check(L.ggb_gemv(C.byref(a), s.cuda_stream))
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for a in args:
                cabi.check(L.ggb_gemv(C.byref(a), torch.cuda.current_stream().cuda_stream))
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * n_buf)
    nbytes = sum(r * (k // 256) * BB[qt] for qt, r in segs)
    print(f"{name:10s} {us:7.2f} us/launch  {nbytes/us/1e3:7.0f} GB/s  ({nbytes/1e6:.1f} MB, ideal {nbytes/6553.3e3:.2f} us)", flush=True)
    return us


SHAPES = {
    "QKV": ([(12, 4096), (12, 1024), (14, 1024)], 4096, cabi.PRO_RMSNORM, cabi.EPI_ROPE_KV),
    "O": ([(12, 4096)], 4096, cabi.PRO_PLAIN, cabi.EPI_RESIDUAL),
    "GU": ([(12, 14336), (12, 14336)], 4096, cabi.PRO_RMSNORM, cabi.EPI_SWIGLU),
    "DOWN": ([(14, 4096)], 14336, cabi.PRO_PLAIN, cabi.EPI_RESIDUAL),
    "DOWN4": ([(12, 4096)], 14336, cabi.PRO_PLAIN, cabi.EPI_RESIDUAL),
}


def bench_seq(names, n_rep=8, reps=20, use_pdl=1):
    """An arbitrary sequence of launch shapes, repeated n_rep times over distinct buffers: consecutive launches may be
    DIFFERENT kernel instances with different shared-memory footprints -- what the decode step really does (minus the
    attention).  Prints us per repetition of the sequence."""
    x = {4096: torch.randn(4096, device=DEV) * 0.01, 14336: torch.randn(14336, device=DEV) * 0.01}
    nw = torch.ones(14336, device=DEV)
    y = [torch.zeros(14336, device=DEV) for _ in range(3)]
    res = torch.zeros(4096, device=DEV)
    pos = torch.tensor([7], dtype=torch.int32, device=DEV)
    tab = torch.zeros(64 * 128, device=DEV)
    kc = torch.zeros(64 * 2048, dtype=torch.int16, device=DEV)
    keep, args = [], []
    for _ in range(n_rep):
        for nm in names:
            segs, k, pro, epi = SHAPES[nm]
            ws = [rand_weight(qt, r, k) for qt, r in segs]
            keep.append(ws)
            args.append(cabi.make_gemv_args([(w.data_ptr(), qt, r, yy.data_ptr()) for w, (qt, r), yy in zip(ws, segs, y)], k, x[k].data_ptr(),
                                            prologue=pro, epilogue=epi, norm_w=nw.data_ptr(), eps=1e-5, use_pdl=use_pdl, residual=res.data_ptr(),
                                            pos_dev=pos.data_ptr(), rope_tab=tab.data_ptr(), n_rot=128, head_dim=128, kcache=kc.data_ptr(),
                                            vcache=kc.data_ptr()))
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for a in args[:len(names)]:
            cabi.check(L.ggb_gemv(C.byref(a), s.cuda_stream))
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for a in args:
                cabi.check(L.ggb_gemv(C.byref(a), torch.cuda.current_stream().cuda_stream))
        for _ in range(3):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * n_rep)
    print(f"sequence {'+'.join(names):24s} pdl={use_pdl}: {us:7.2f} us per repetition", flush=True)
    return us


if __name__ == "__main__":
    pdl = int(os.environ.get("PDL", "1"))
    t = 0
    t += bench("QKV", [(12, 4096), (12, 1024), (14, 1024)], 4096, cabi.PRO_RMSNORM, cabi.EPI_ROPE_KV, use_pdl=pdl)
    t += bench("O", [(12, 4096)], 4096, cabi.PRO_PLAIN, cabi.EPI_RESIDUAL, use_pdl=pdl)
    t += bench("GATEUP", [(12, 14336), (12, 14336)], 4096, cabi.PRO_RMSNORM, cabi.EPI_SWIGLU, use_pdl=pdl)
    t += bench("DOWN_q6", [(14, 4096)], 14336, cabi.PRO_PLAIN, cabi.EPI_RESIDUAL, use_pdl=pdl)
    d4 = bench("DOWN_q4", [(12, 4096)], 14336, cabi.PRO_PLAIN, cabi.EPI_RESIDUAL, use_pdl=pdl)
    bench("HEAD", [(14, 128256)], 4096, cabi.PRO_RMSNORM, cabi.EPI_ARGMAX, n_buf=4, use_pdl=pdl)
    print(f"layer GEMV sum (q6 down) {t:.1f} us -> 32 layers {t*32/1e3:.2f} ms")
    for seq in (["QKV"], ["O"], ["GU"], ["DOWN"], ["O", "GU"], ["GU", "DOWN"], ["DOWN", "QKV"], ["QKV", "O"], ["GU", "DOWN4"], ["QKV", "O", "GU", "DOWN"]):
        for p in (1, 0):
            bench_seq(seq, use_pdl=p)
