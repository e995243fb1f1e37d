"""Small profiling driver: a depth-trimmed model with the REAL per-launch shapes of a preset (same d, ff, heads,
vocab), a few eager decode steps.  Used under ncu (see profiles/README.md)."""
import argparse
import os
import sys
from dataclasses import replace

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama3-8b")
    ap.add_argument("--ftype", default="Q4_K_M")
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--graph", action="store_true")
    ap.add_argument("--no-pdl", action="store_true")
    ap.add_argument("--pos", type=int, default=256, help="start decoding at this position (KV filled with zeros)")
    a = ap.parse_args()
    import torch
    from ggufb200 import synth
    from ggufb200.model import Engine
    cfg = replace(synth.PRESETS[a.model], n_layer=a.layers)
    path = f"/dev/shm/prof-{a.model}-L{a.layers}-{a.ftype}.gguf"
    if not os.path.exists(path):
        synth.write_gguf(path, cfg, a.ftype, 0xB200)
    eng = Engine(path, n_ctx=1024, use_graph=a.graph, use_pdl=not a.no_pdl)
    eng.warmup()
    eng.reset()
    eng.prefill([1, 300, 301])
    with torch.cuda.stream(eng.stream):
        eng._set_tok_pos(300, a.pos)
    eng.stream.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(eng.stream)
    eng.decode(a.steps)
    ev1.record(eng.stream)
    torch.cuda.synchronize()
    print(f"{a.steps} steps, {a.layers} layers: {ev0.elapsed_time(ev1) / a.steps * 1e3:.1f} us/step", flush=True)


if __name__ == "__main__":
    main()
