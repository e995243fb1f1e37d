"""In-kernel timeline of a PDL chain of identical GEMV launches (needs a build with -DGGB_TIMELINE):
   GGB_NVCC_EXTRA=-DGGB_TIMELINE python llama-gguf-inference_b200/build.py --force"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ggufb200 import cabi  # noqa: E402
import tools.gemv_bench as GB  # noqa: E402

L = cabi.lib()
L.ggb_debug_timeline.argtypes = [C.c_void_p]
L.ggb_debug_timeline.restype = C.c_int


def run(name, segs, k, pro, epi):
    # 8 launches in a chain (eager, PDL); slots = launch index % 8
    GB.bench(name, segs, k, pro, epi, n_buf=8, reps=1)
    buf = np.zeros(8 * 1024 * 8, dtype=np.uint64)
    assert L.ggb_debug_timeline(buf.ctypes.data) == 0
    t = buf.reshape(8, 1024, 8)[:, :148, :6].astype(np.int64)
    # the last graph replay wrote slots in launch order; order slots by their mean entry time
    order = np.argsort(t[:, :, 0].mean(axis=1))
    t = t[order]
    t0 = t[0, :, 0].min()
    print(f"== {name}: per launch, ns relative to first entry: [entry, ring issued, dep wait done, prologue done, main done, exit] (median over CTAs; min..max of exit)")
    for i in range(8):
        med = np.median(t[i], axis=0) - t0
        print(f"   launch {i}: " + " ".join(f"{int(v):7d}" for v in med) + f"   exit range {int(t[i,:,5].min()-t0)}..{int(t[i,:,5].max()-t0)}  entry range {int(t[i,:,0].min()-t0)}..{int(t[i,:,0].max()-t0)}")


if __name__ == "__main__":
    run("O", [(12, 4096)], 4096, cabi.PRO_PLAIN, cabi.EPI_RESIDUAL)
    run("GATEUP", [(12, 14336), (12, 14336)], 4096, cabi.PRO_RMSNORM, cabi.EPI_SWIGLU)
    run("DOWN_q6", [(14, 4096)], 14336, cabi.PRO_PLAIN, cabi.EPI_RESIDUAL)
