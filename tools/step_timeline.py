"""In-kernel timeline of the LAST eight GEMV launches of a real decode step (graph replay, PDL): needs a build with
-DGGB_TIMELINE (GGB_NVCC_EXTRA=-DGGB_TIMELINE python llama-gguf-inference_b200/build.py --force; GGB_LIB_PATH=that .so).
Prints, per launch, the medians over the 148 CTAs of: entry, ring issued, dependency wait done, prologue done, main loop
done, exit -- relative to the first entry -- so the gaps BETWEEN launches of different shapes are visible in situ."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from ggufb200 import cabi  # noqa: E402
from ggufb200.model import Engine  # noqa: E402

L = cabi.lib()
L.ggb_debug_timeline.argtypes = [C.c_void_p]
L.ggb_debug_timeline.restype = C.c_int
path = bench.model_path("llama3-8b", "Q4_K_M", 0xB200)
eng = Engine(path, n_ctx=1024)
eng.warmup()
eng.reset()
eng.prefill([1] + list(range(300, 555)))
eng.decode(6)
torch.cuda.synchronize()
buf = np.zeros(8 * 1024 * 8, dtype=np.uint64)
assert L.ggb_debug_timeline(buf.ctypes.data) == 0
t = buf.reshape(8, 1024, 8)[:, :148, :6].astype(np.int64)
order = np.argsort(t[:, :, 0].mean(axis=1))
t = t[order]
t0 = t[0, :, 0].min()
names = ["O(30)", "GU(30)", "DOWN(30)", "QKV(31)", "O(31)", "GU(31)", "DOWN(31)", "HEAD"]
print("launch      entry  ring_issued  dep_done  prologue_done  main_done   exit(med)  exit(max) | dur(dep->exit max)  gap to next dep_done")
for i in range(8):
    med = (np.median(t[i], axis=0) - t0) / 1e3
    ex = (t[i, :, 5].max() - t0) / 1e3
    nxt = (np.median(t[i + 1, :, 2]) - t0) / 1e3 if i + 1 < 8 else float("nan")
    print(f"{names[i]:9s} " + " ".join(f"{v:10.2f}" for v in med) + f" {ex:10.2f} | {ex - med[2]:8.2f}   {nxt - ex:8.2f}"
          f"   entry min/max {(t[i, :, 0].min() - t0) / 1e3:.2f}/{(t[i, :, 0].max() - t0) / 1e3:.2f}  main_done max {(t[i, :, 4].max() - t0) / 1e3:.2f}")
try:
    L.ggb_debug_timeline_attn.argtypes = [C.c_void_p]
    L.ggb_debug_timeline_attn.restype = C.c_int
    ab = np.zeros(2 * 512 * 8, dtype=np.uint64)
    assert L.ggb_debug_timeline_attn(ab.ctypes.data) == 0
    a = ab[:512 * 8].reshape(512, 8)[:256, [0, 1, 5, 2, 6, 7, 3, 4]].astype(np.int64)
    print("attention of the last layer (256 CTAs), us relative to the same origin: [entry, dependency wait done, pass-1 loop done, max exchanged, "
          "pass-2 loop done, partials pushed, sums exchanged, exit]")
    print("   min   ", " ".join(f"{(a[:, i].min() - t0) / 1e3:9.2f}" for i in range(8)))
    print("   median", " ".join(f"{(np.median(a[:, i]) - t0) / 1e3:9.2f}" for i in range(8)))
    print("   max   ", " ".join(f"{(a[:, i].max() - t0) / 1e3:9.2f}" for i in range(8)))
    ck = ab[512 * 8:].reshape(512, 8)[:256, [0, 1, 5, 2, 6, 7, 3, 4]].astype(np.int64)
    d = np.diff(ck, axis=1)
    print("   the same intervals in SM clock cycles (clock64, per CTA): median / max")
    print("   median", " ".join(f"{np.median(d[:, i]):9.0f}" for i in range(7)))
    print("   max   ", " ".join(f"{d[:, i].max():9.0f}" for i in range(7)))
except AttributeError:
    pass
eng.close()
