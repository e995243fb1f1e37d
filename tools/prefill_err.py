"""Error of the tensor-core prefill (fp16 dequant-GEMM + f16 flash attention) against the bit-exact integer path of the
same engine, on the synthetic test models: relative L2 / max-abs error of the logits after a 71-token and a 300-token
prompt.  Quoted in DESIGN.md and used to set the bounds in tests/test_gpu_engine.py."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ggufb200 import synth  # noqa: E402
from ggufb200.model import Engine  # noqa: E402

d = tempfile.mkdtemp()
for preset, ftype, damped in [("small", "Q4_K_M", False), ("medium", "Q4_K_M", False), ("medium", "Q8_0", False), ("medium", "Q6_K", False),
                              ("medium", "Q5_K_M", False), ("medium", "Q4_K_M", True)]:
    path = os.path.join(d, f"{preset}-{ftype}-{damped}.gguf")
    synth.write_gguf(path, synth.damped(preset) if damped else preset, ftype, seed=0xB200)
    eng = Engine(path, n_ctx=512)
    eng.warmup()
    for n in (70, 299):
        rng = np.random.default_rng(5)
        prompt = [1] + [int(t) for t in rng.integers(300, 500, size=n)]
        eng.gemm_prefill_min = 10 ** 9
        eng.reset(); eng.prefill(prompt)
        a = eng.last_logits().copy()
        eng.gemm_prefill_min = 16
        eng.reset(); eng.prefill(prompt)
        b = eng.last_logits().copy()
        print(f"{preset:7s} {ftype:7s} damped={int(damped)} prompt {n + 1:4d}: rel L2 {np.linalg.norm(a - b) / np.linalg.norm(a):.3e}  "
              f"max abs / max |logit| {np.abs(a - b).max() / np.abs(a).max():.3e}  argmax equal {int(a.argmax() == b.argmax())}", flush=True)
    eng.close()
