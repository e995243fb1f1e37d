"""Load time of a GGUF file into HBM (BASELINE/SURVEY f4: scripts/start.sh:600-635 gives the backend ~33 s to answer
/health): the pipelined loader (pinned staging ring + reader threads + on-GPU re-ordering) against the per-tensor path.
usage: python tools/load_bench.py [--model llama3-8b] [--ftype Q4_K_M]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from ggufb200.model import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model", default="llama3-8b")
ap.add_argument("--ftype", default="Q4_K_M")
a = ap.parse_args()
path = bench.model_path(a.model, a.ftype, 0xB200)
size = os.path.getsize(path)
toks = {}
for mode in ("0", "1", "1", "0", "1", "1"):
    os.environ["GGB_LOAD_PIPELINE"] = mode
    torch.cuda.synchronize()
    t0 = time.time()
    eng = Engine(path, n_ctx=512)
    torch.cuda.synchronize()
    dt = time.time() - t0
    eng.warmup()
    toks[mode] = eng.generate([1, 300, 301, 302, 303], 8)
    print(f"{a.model} {a.ftype} {size / 1e9:.2f} GB  pipeline={mode}: Engine() {dt:6.2f} s  (weights {eng.load_seconds:6.2f} s, {size / 1e9 / eng.load_seconds:5.2f} GB/s)  tokens {toks[mode]}", flush=True)
    eng.close()
    del eng
    torch.cuda.empty_cache()
assert toks["0"] == toks["1"], "the two load paths give different models"
