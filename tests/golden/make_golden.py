"""Generates the committed golden vectors under tests/golden/ (run from the repo root in the BUILD container).

Sources of truth:
  * gguf-py 0.19.0 (`gguf.quants`, the python package published from the llama.cpp tree; present in this image)
    for dequantisation of Q8_0/Q4_K/Q5_K/Q6_K/Q4_0/Q5_0 blocks and for Q8_0 quantisation -- these PIN the oracle;
  * the oracle itself for end-to-end regression vectors (greedy tokens + logits of the seeded `tiny`
    synthetic model) -- these pin nothing upstream, they only detect drift of the oracle and of the
    synthetic-model generator.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    import gguf
    from gguf import GGMLQuantizationType as T
    from conftest import rand_blocks
    from oracle import oracle as O

    rng = np.random.default_rng(20261018)
    for name, qt, gt in (("q8_0", O.Q8_0, T.Q8_0), ("q4_k", O.Q4_K, T.Q4_K), ("q5_k", O.Q5_K, T.Q5_K), ("q6_k", O.Q6_K, T.Q6_K),
                         ("q4_0", O.Q4_0, T.Q4_0), ("q5_0", O.Q5_0, T.Q5_0)):
        if os.path.exists(os.path.join(HERE, f"dequant_{name}.npz")):
            continue          # committed vectors stay as they are; new formats are added
        tame = rand_blocks(qt, 48, rng)
        wild = rand_blocks(qt, 16, rng, wild=True)
        raw = np.concatenate([tame, wild])
        with np.errstate(all="ignore"):
            ref = gguf.quants.dequantize(raw.reshape(-1), gt).reshape(-1).astype(np.float32)
        np.savez_compressed(os.path.join(HERE, f"dequant_{name}.npz"), raw=raw, out_bits=ref.view(np.uint32),
                            source=np.array(f"gguf-py {getattr(gguf, '__version__', '0.19.0')} gguf.quants.dequantize"))
    x = (rng.standard_normal(32 * 64) * np.exp(rng.uniform(-4, 4, 32 * 64))).astype(np.float32)
    x[:32] = 0
    q = gguf.quants.quantize(x.reshape(1, -1), T.Q8_0).reshape(-1)
    np.savez_compressed(os.path.join(HERE, "quantize_q8_0.npz"), x=x, packed=q)

    # oracle regression vector on the seeded tiny model
    from ggufb200 import synth
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        for ftype in ("Q4_K_M", "Q8_0"):
            p = os.path.join(d, "m.gguf")
            synth.write_gguf(p, "tiny", ftype, seed=0xB200)
            m = O.OracleLlama(p, n_ctx=64)
            toks, logits = m.greedy([1, 300, 301, 302], 24, return_logits=True)
            np.savez_compressed(os.path.join(HERE, f"oracle_tiny_{ftype}.npz"), prompt=np.array([1, 300, 301, 302]),
                                tokens=np.array(toks), first_logits=logits[0], last_logits=logits[-1])
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
