"""Where does the tensor-core prefill leave the integer path?  Compares the final residual stream x and the logits of
the last prompt token between Engine.prefill through the GEMM path and through the exact kernels, for depth-trimmed
variants of a preset (diagnosis tool)."""
import os
import sys
import tempfile
from dataclasses import replace

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ggufb200 import synth  # noqa: E402
from ggufb200.model import Engine  # noqa: E402

d = tempfile.mkdtemp()
base = synth.PRESETS["medium"]
for damp, nl in [(0.0, 2), (0.003, 1), (0.003, 2), (0.003, base.n_layer), (None, 1), (None, 2)]:
    cfg = replace(base, n_layer=nl)
    if damp is not None:
        cfg = replace(synth.damped(cfg, max(damp, 1e-12)))
    path = os.path.join(d, f"m-{damp}-{nl}.gguf")
    synth.write_gguf(path, cfg, "Q4_K_M", seed=0xB200)
    eng = Engine(path, n_ctx=512)
    eng.warmup()
    rng = np.random.default_rng(5)
    prompt = [1] + [int(t) for t in rng.integers(300, 500, size=70)]
    res = {}
    for name, floor in (("exact", 10 ** 9), ("gemm", 16)):
        eng.gemm_prefill_min = floor
        eng.reset(); eng.prefill(prompt)
        eng.stream.synchronize()
        res[name] = (eng.slots[0].x.cpu().numpy().copy(), eng.last_logits().copy(), eng.slots[0].kc.float().cpu().numpy()[:, :len(prompt)].copy())
    xa, la, ka = res["exact"]; xb, lb, kb = res["gemm"]
    print(f"damp {damp} layers {nl}: x rel L2 {np.linalg.norm(xa - xb) / np.linalg.norm(xa):.3e}  logits rel L2 {np.linalg.norm(la - lb) / np.linalg.norm(la):.3e}  "
          f"K cache max rel {np.abs(ka - kb).max() / np.abs(ka).max():.3e}  |x| {np.linalg.norm(xa):.3e}", flush=True)
    eng.close()
