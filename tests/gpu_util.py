"""Helpers for the -m gpu tests: torch <-> C-ABI plumbing (device memory only; all compute is libggufb200)."""
import ctypes as C

import numpy as np
import torch

from ggufb200 import cabi

DEV = torch.device("cuda", 0)


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def to_dev(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def sync():
    torch.cuda.synchronize()


def gpu_dequant(qtype: int, raw: np.ndarray, n: int) -> np.ndarray:
    L = cabi.lib()
    w = to_dev(raw.reshape(-1)) if raw.size else torch.empty(0, dtype=torch.uint8, device=DEV)
    out = torch.empty(n, dtype=torch.float32, device=DEV)
    cabi.check(L.ggb_dequant(qtype, w.data_ptr(), out.data_ptr(), n, stream_ptr()), "ggb_dequant")
    sync()
    return out.cpu().numpy()


def gpu_repack(qtype: int, raw: np.ndarray, rows: int, k: int) -> torch.Tensor:
    L = cabi.lib()
    stride = L.ggb_repacked_row_stride(qtype, k)
    assert stride > 0
    src = to_dev(raw.reshape(-1))
    dst = torch.zeros(rows * stride + 16, dtype=torch.uint8, device=DEV)  # +16: bulk copies round the last tile up
    cabi.check(L.ggb_repack(qtype, src.data_ptr(), dst.data_ptr(), rows, k, stream_ptr()), "ggb_repack")
    sync()
    return dst


def gpu_gemv(segs, k, x, **kw):
    """segs: list of (repacked tensor, qtype, rows); returns list of per-segment outputs (numpy) for STORE."""
    L = cabi.lib()
    xd = to_dev(x.astype(np.float32))
    ys = [torch.zeros(max(r, 1), dtype=torch.float32, device=DEV) for _, _, r in segs]
    a = cabi.make_gemv_args([(w.data_ptr(), t, r, y.data_ptr()) for (w, t, r), y in zip(segs, ys)], k, xd.data_ptr(), **kw)
    cabi.check(L.ggb_gemv(C.byref(a), stream_ptr()), "ggb_gemv")
    sync()
    return [y.cpu().numpy()[:r] for y, (_, _, r) in zip(ys, segs)]
